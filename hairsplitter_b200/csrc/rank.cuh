// Allele ranking with the reference's tie-breaking, usable from device and host code.
//
// The reference ranks the codes of a pileup column by iterating a robin_hood::unordered_map
// <unsigned char,int> and std::sort-ing the (code,count) pairs by count (src/call_variants.cpp:477-501),
// and picks the alternative allele of a partition by iterating such a map and keeping the first
// strict maximum (:832-844). Both results depend on (a) the slot order of robin_hood's flat table
// (src/robin_hood.h: hash_int :749, keyToIdx :1349, insertKeyPrepareEmptySpot :2332, shiftUp :1377,
// insert_move :1451, try_increase_info :2385, increase_size :2414, rehashPowerOfTwo :2203) and
// (b) libstdc++'s introsort. This header re-implements the observable behaviour of both for the
// only case the hot path needs: <= 128 distinct unsigned-char keys inserted once each.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HS_HD __host__ __device__ __forceinline__
#else
#define HS_HD inline
#endif

#define HS_RH_MAXKEYS 128
// 128 keys -> at most 256 buckets (80 % load) + overflow buffer min(204,255) + 8 bytes padding
#define HS_RH_SLOTS (256 + 204 + 8)

struct HsRhTable {
    uint8_t info[HS_RH_SLOTS];
    uint8_t key[HS_RH_SLOTS];
    uint64_t mult;
    uint32_t mask, n, max_allowed, info_inc, info_shift;
};

HS_HD uint32_t hs_rh_max_allowed(uint32_t buckets) { return buckets * 80u / 100u; }
HS_HD uint32_t hs_rh_slots(uint32_t buckets) {
    uint32_t m = hs_rh_max_allowed(buckets);
    return buckets + (m < 255u ? m : 255u);
}
HS_HD void hs_rh_init_data(HsRhTable& t, uint32_t buckets) {
    t.n = 0;
    t.mask = buckets - 1;
    t.max_allowed = hs_rh_max_allowed(buckets);
    uint32_t nb = hs_rh_slots(buckets);
    for (uint32_t i = 0; i < nb + 8; i++) t.info[i] = 0;
    t.info[nb] = 1;  // sentinel
    t.info_inc = 32;
    t.info_shift = 0;
}
HS_HD void hs_rh_new(HsRhTable& t) {
    t.mult = 0xc4ceb9fe1a85ec53ull;
    t.mask = 0;
    t.n = 0;
    t.max_allowed = 0;
    t.info_inc = 32;
    t.info_shift = 0;
}
HS_HD void hs_rh_key_to_idx(const HsRhTable& t, uint8_t key, uint32_t& idx, uint32_t& info) {
    uint64_t h = (uint64_t)key;  // hash_int: x ^= x >> 33 is the identity for x < 256
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= t.mult;
    h ^= h >> 33;
    info = t.info_inc + (uint32_t)((h & 31u) >> t.info_shift);
    idx = (uint32_t)(h >> 5) & t.mask;
}
HS_HD void hs_rh_shift_up(HsRhTable& t, uint32_t start, uint32_t ins) {
    for (uint32_t i = start; i != ins; i--) {
        t.key[i] = t.key[i - 1];
        uint32_t v = (uint32_t)t.info[i - 1] + t.info_inc;
        t.info[i] = (uint8_t)v;
        if ((uint32_t)t.info[i] + t.info_inc > 0xFF) t.max_allowed = 0;
    }
}
HS_HD bool hs_rh_try_increase_info(HsRhTable& t) {
    if (t.info_inc <= 2) return false;
    t.info_inc >>= 1;
    t.info_shift++;
    uint32_t nb = hs_rh_slots(t.mask + 1);
    for (uint32_t i = 0; i < nb + 8; i++) t.info[i] = (uint8_t)((t.info[i] >> 1) & 0x7f);
    t.info[nb] = 1;
    t.max_allowed = hs_rh_max_allowed(t.mask + 1);
    return true;
}
// place a key that is known to be absent; `le` selects insert_move's "<=" skip (rehash) -- with
// distinct keys both loops of insertKeyPrepareEmptySpot amount to the same skip
HS_HD void hs_rh_place(HsRhTable& t, uint8_t key) {
    uint32_t idx, info;
    hs_rh_key_to_idx(t, key, idx, info);
    while (info <= t.info[idx]) {
        idx++;
        info += t.info_inc;
    }
    uint32_t ins = idx;
    if ((info & 0xFF) + t.info_inc > 0xFF || info + t.info_inc > 0xFF) t.max_allowed = 0;
    while (t.info[idx] != 0) idx++;
    if (idx != ins) hs_rh_shift_up(t, idx, ins);
    t.key[ins] = key;
    t.info[ins] = (uint8_t)info;
    t.n++;
}
HS_HD void hs_rh_rehash(HsRhTable& t, uint32_t buckets) {
    uint8_t old[HS_RH_MAXKEYS];
    uint32_t nb = hs_rh_slots(t.mask + 1), m = 0;
    for (uint32_t i = 0; i < nb; i++)
        if (t.info[i] != 0) old[m++] = t.key[i];
    hs_rh_init_data(t, buckets);
    for (uint32_t i = 0; i < m; i++) {
        if (t.max_allowed == 0) hs_rh_try_increase_info(t);  // insert_move :1454 (overflow would throw)
        hs_rh_place(t, old[i]);
    }
}
HS_HD void hs_rh_insert(HsRhTable& t, uint8_t key) {
    for (int attempt = 0; attempt < 256; attempt++) {
        if (t.n >= t.max_allowed) {  // also the unallocated state (mask == 0, max_allowed == 0)
            if (t.mask == 0) {
                hs_rh_init_data(t, 8);
            } else if (t.n < hs_rh_max_allowed(t.mask + 1) && hs_rh_try_increase_info(t)) {
                // info bytes widened, retry
            } else {
                t.mult += 0xc4ceb9fe1a85ec54ull;
                if (t.n * 2 < hs_rh_max_allowed(t.mask + 1)) hs_rh_rehash(t, t.mask + 1);
                else hs_rh_rehash(t, (t.mask + 1) * 2);
            }
            continue;
        }
        hs_rh_place(t, key);
        return;
    }
}
// keys in iteration (slot) order; returns their number
HS_HD int hs_rh_iterate(const HsRhTable& t, uint8_t* out) {
    if (t.mask == 0) return 0;
    uint32_t nb = hs_rh_slots(t.mask + 1);
    int k = 0;
    for (uint32_t i = 0; i < nb; i++)
        if (t.info[i] != 0) out[k++] = t.key[i];
    return k;
}

// ---- libstdc++ std::sort on packed (count << 8 | key) entries, comparator count descending ------
#define HS_KC_COMP(a, b) (((a) >> 8) > ((b) >> 8))

HS_HD void hs_kc_unguarded_linear_insert(uint32_t* v, int last) {
    uint32_t val = v[last];
    int next = last - 1;
    while (HS_KC_COMP(val, v[next])) {
        v[last] = v[next];
        last = next;
        --next;
    }
    v[last] = val;
}
HS_HD void hs_kc_insertion_sort(uint32_t* v, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (HS_KC_COMP(v[i], v[first])) {
            uint32_t val = v[i];
            for (int j = i; j > first; j--) v[j] = v[j - 1];
            v[first] = val;
        } else {
            hs_kc_unguarded_linear_insert(v, i);
        }
    }
}
HS_HD void hs_kc_adjust_heap(uint32_t* f, int hole, int len, uint32_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (HS_KC_COMP(f[child], f[child - 1])) child--;
        f[hole] = f[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        f[hole] = f[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && HS_KC_COMP(f[parent], value)) {
        f[hole] = f[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    f[hole] = value;
}
HS_HD void hs_kc_heap_sort(uint32_t* f, int len) {
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            uint32_t v = f[parent];
            hs_kc_adjust_heap(f, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int last = len; last > 1;) {
        --last;
        uint32_t v = f[last];
        f[last] = f[0];
        hs_kc_adjust_heap(f, 0, last, v);
    }
}
HS_HD void hs_kc_swap(uint32_t& a, uint32_t& b) {
    uint32_t t = a;
    a = b;
    b = t;
}
HS_HD void hs_kc_std_sort(uint32_t* v, int n) {
    if (n <= 1) return;
    if (n > 16) {
        int lg = 0;
        while ((1 << (lg + 1)) <= n) lg++;
        // __introsort_loop: recursion on the right part, iteration on the left -> explicit stack
        int stack_first[128], stack_last[128], stack_depth[128];  // <= one entry per partition step
        int sp = 0;
        stack_first[0] = 0;
        stack_last[0] = n;
        stack_depth[0] = 2 * lg;
        sp = 1;
        while (sp > 0) {
            sp--;
            int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
            // The recursive original fully processes [cut,last) BEFORE continuing with [first,cut); the
            // sub-ranges are disjoint, so the order in which they are processed does not change the result.
            while (last - first > 16) {
                if (depth == 0) {
                    hs_kc_heap_sort(v + first, last - first);
                    break;
                }
                --depth;
                int mid = first + (last - first) / 2;
                // __move_median_to_first(first, first+1, mid, last-1)
                int a = first + 1, b = mid, c = last - 1;
                if (HS_KC_COMP(v[a], v[b])) {
                    if (HS_KC_COMP(v[b], v[c])) hs_kc_swap(v[first], v[b]);
                    else if (HS_KC_COMP(v[a], v[c])) hs_kc_swap(v[first], v[c]);
                    else hs_kc_swap(v[first], v[a]);
                } else if (HS_KC_COMP(v[a], v[c])) hs_kc_swap(v[first], v[a]);
                else if (HS_KC_COMP(v[b], v[c])) hs_kc_swap(v[first], v[c]);
                else hs_kc_swap(v[first], v[b]);
                // __unguarded_partition(first+1, last, pivot = first)
                int lo = first + 1, hi = last;
                for (;;) {
                    while (HS_KC_COMP(v[lo], v[first])) ++lo;
                    --hi;
                    while (HS_KC_COMP(v[first], v[hi])) --hi;
                    if (!(lo < hi)) break;
                    hs_kc_swap(v[lo], v[hi]);
                    ++lo;
                }
                int cut = lo;
                stack_first[sp] = cut;
                stack_last[sp] = last;
                stack_depth[sp] = depth;
                sp++;
                last = cut;
            }
        }
        hs_kc_insertion_sort(v, 0, 16);
        for (int i = 16; i != n; ++i) hs_kc_unguarded_linear_insert(v, i);
    } else {
        hs_kc_insertion_sort(v, 0, n);
    }
}

// ---- column ranking: fast tie resolution + literal replay fallback ---------------------------------
// For n = m + 3 keys (m distinct codes, then the dummy keys 0,1,2; src/call_variants.cpp:477-494) the
// final robin_hood table has 8 / 16 / 32 buckets for n <= 6 / 12 / 25 and has been rehashed 0 / 1 / 2
// times (each rehash advances the hash multiplier, robin_hood.h:2446-2449). A flat robin-hood table
// without wrap-around keeps its entries sorted by home bucket, so two keys with DIFFERENT home buckets
// iterate in bucket order whatever the insertion history was. For n <= 16 libstdc++'s std::sort is a
// plain (stable) insertion sort, so equal counts keep the iteration order. Hence: ties between keys with
// pairwise distinct home buckets are decided by the precomputed bucket table below; everything else
// (shared bucket, n > 16 where introsort is unstable) replays the reference's map + sort literally.
struct alignas(16) HsRankLut {
    // the part hs_rank_fast reads (HS_RANK_LUT_FAST_BYTES; the column kernel stages only this much)
    uint16_t tie[4][160];  // (255 - home) << 8 | key: the tie-breaking half of hs_rank_fast's sort word
    uint8_t single[128];   // m == 1: second_base (a dummy key) for code 33 + i
    uint8_t empty[2];      // m == 0: ref_base, second_base
    uint8_t pad_[2];
    // hs_rank_hashbits
    uint8_t home[4][160];  // home bucket of key k in the table incarnation `level`
    uint8_t hbits[4][160]; // the 5 hash bits robin_hood keeps in the info byte (robin_hood.h:1349-1356)
};
#define HS_RANK_LUT_FAST_BYTES (4 * 160 * 2 + 128 + 4)

// 8 / 16 / 32 / 64 buckets hold up to 6 / 12 / 25 / 51 keys (80 % load); larger n never reaches the tables
HS_HD int hs_rank_level(int n) { return n <= 6 ? 0 : (n <= 12 ? 1 : (n <= 25 ? 2 : 3)); }

// literal replay: keys in first-seen order via acc.key(k), counts via acc.count(key)
template <class Acc>
HS_HD void hs_rank_literal(const Acc& acc, int m, int& k0, int& k1, unsigned& c0, unsigned& c1, unsigned& c2) {
    HsRhTable t;
    hs_rh_new(t);
    for (int k = 0; k < m; k++) hs_rh_insert(t, (uint8_t)acc.key(k));
    hs_rh_insert(t, 0);
    hs_rh_insert(t, 1);
    hs_rh_insert(t, 2);
    uint8_t it[HS_RH_MAXKEYS];
    uint32_t kc[HS_RH_MAXKEYS];
    const int n = hs_rh_iterate(t, it);
    for (int i = 0; i < n; i++) {
        const int key = it[i];
        const unsigned cnt = key >= 33 ? acc.count(key) : 0u;
        kc[i] = (cnt << 8) | (unsigned)key;
    }
    hs_kc_std_sort(kc, n);
    k0 = kc[0] & 0xff;
    k1 = kc[1] & 0xff;
    c0 = kc[0] >> 8;
    c1 = kc[1] >> 8;
    c2 = kc[2] >> 8;
}

// returns 0 when resolved on the fast path, 1 when the literal replay is needed (outputs then unset).
// Branch-free in the keys: every key becomes one 32-bit word  count << 16 | (255 - home bucket) << 8 | key
// (lut->tie), so that "larger word" means "earlier in the reference's sorted vector" whenever the two keys
// differ in count or in home bucket; the three largest words come out of a 5-instruction min/max network.
template <class Acc>
HS_HD int hs_rank_fast(const Acc& acc, int m, const HsRankLut* lut, int& k0, int& k1, unsigned& c0, unsigned& c1,
                       unsigned& c2) {
    if (m == 0) {
        k0 = lut->empty[0];
        k1 = lut->empty[1];
        c0 = c1 = c2 = 0;
        return 0;
    }
    if (m == 1) {
        k0 = acc.key(0);
        k1 = lut->single[k0 - 33];
        c0 = acc.count(k0);
        c1 = c2 = 0;
        return 0;
    }
    const int n = m + 3;
    const uint16_t* tie = lut->tie[hs_rank_level(n)];
    uint32_t t0 = 0, t1 = 0, t2 = 0;
    for (int k = 0; k < m; k++) {
        const int key = acc.key(k);
        uint32_t v = ((uint32_t)acc.count(key) << 16) | tie[key];
        uint32_t hi = t0 > v ? t0 : v;
        v = t0 > v ? v : t0;
        t0 = hi;
        hi = t1 > v ? t1 : v;
        v = t1 > v ? v : t1;
        t1 = hi;
        t2 = t2 > v ? t2 : v;
    }
    c0 = t0 >> 16;
    c1 = t1 >> 16;
    c2 = t2 >> 16;
    k0 = (int)(t0 & 0xffu);
    k1 = (int)(t1 & 0xffu);
    if (c0 > c1 && c1 > c2) return 0;  // both ranks unique
    if (n > 16) return 1;              // introsort is not stable
    // neighbours in the ranking that agree in count AND home bucket: their slot order depends on the history
    if ((t0 >> 8) == (t1 >> 8) || (t1 >> 8) == (t2 >> 8)) return 1;
    return 0;
}

// Second-level tie resolution for n <= 16: keys that share a home bucket. insertKeyPrepareEmptySpot skips
// occupants whose info byte is >= the new key's (robin_hood.h:2339-2343); the info byte is
// (distance+1)*32 + 5 hash bits, so inside one home bucket the entries end up ordered by those hash bits,
// descending, and only keys that agree in bucket AND hash bits fall back on the insertion history. The
// 5-bit comparison holds as long as the final table incarnation never widened its info bytes
// (try_increase_info, :2385), i.e. while no entry sits 6 or more slots from home; displacements only
// grow, so checking the final layout is enough. Returns 0 when resolved, 1 when the literal replay is
// still needed. Keys/counts are given in first-seen order, like for hs_rank_literal.
template <class Acc>
HS_HD int hs_rank_hashbits(const Acc& acc, int m, const HsRankLut* lut, int& k0, int& k1, unsigned& c0, unsigned& c1,
                           unsigned& c2) {
    const int n = m + 3;
    if (m < 2 || n > 16) return 1;
    const int level = hs_rank_level(n);
    const uint8_t* home = lut->home[level];
    const uint8_t* hbits = lut->hbits[level];
    // final layout: bucket occupancy as 4-bit counters (n <= 16 keys, up to 32 buckets)
    unsigned long long occ_lo = 0, occ_hi = 0;
    for (int k = -3; k < m; k++) {
        const int key = k < 0 ? k + 3 : acc.key(k);  // the dummy keys 0,1,2 (:477-494) sit in the table too
        const int h = home[key];
        if (h < 16) occ_lo += 1ull << (4 * h); else occ_hi += 1ull << (4 * (h - 16));
    }
    int next_free = 0, maxd = 0;
    const int nb = 8 << level;
    for (int b = 0; b < nb; b++) {
        const int cnt = (int)(((b < 16) ? (occ_lo >> (4 * b)) : (occ_hi >> (4 * (b - 16)))) & 15ull);
        if (next_free < b) next_free = b;
        next_free += cnt;
        if (cnt && next_free - 1 - b > maxd) maxd = next_free - 1 - b;
    }
    if (maxd >= 6) return 1;
    c0 = c1 = c2 = 0;
    k0 = k1 = 0;
    for (int k = 0; k < m; k++) {
        const unsigned cnt = acc.count(acc.key(k));
        if (cnt > c0) { c2 = c1; c1 = c0; c0 = cnt; }
        else if (cnt > c1) { c2 = c1; c1 = cnt; }
        else if (cnt > c2) { c2 = cnt; }
    }
    // iteration rank of a key: (home bucket, 31 - hash bits); the three smallest among count == c0 and the
    // two smallest among count == c1 decide
    int a0 = 0, a1 = 0, b0 = 0, ra0 = 4096, ra1 = 4096, ra2 = 4096, rb0 = 4096, rb1 = 4096, n0 = 0, n1 = 0;
    for (int k = 0; k < m; k++) {
        const int key = acc.key(k);
        const unsigned cnt = acc.count(key);
        const int rk = home[key] * 32 + 31 - hbits[key];
        if (cnt == c0) {
            n0++;
            if (rk < ra0) { ra2 = ra1; ra1 = ra0; a1 = a0; ra0 = rk; a0 = key; }
            else if (rk < ra1) { ra2 = ra1; ra1 = rk; a1 = key; }
            else if (rk < ra2) { ra2 = rk; }
        } else if (cnt == c1) {
            n1++;
            if (rk < rb0) { rb1 = rb0; rb0 = rk; b0 = key; }
            else if (rk < rb1) { rb1 = rk; }
        }
    }
    if (n0 >= 2) {  // c1 == c0 here: ranks 0 and 1 are the first two of the c0 set
        if (ra0 == ra1 || ra1 == ra2) return 1;
        k0 = a0;
        k1 = a1;
    } else {
        k0 = a0;
        if (n1 >= 2 && rb0 == rb1) return 1;
        k1 = b0;
    }
    return 0;
}

// Third-level resolution for any n <= 51 (the columns whose ties meet libstdc++'s unstable introsort, n > 16):
// the COMPLETE slot order of the final table incarnation is rebuilt from (home bucket, hash bits) -- under the
// conditions of hs_rank_hashbits, now required of every key: no two keys agree in bucket and hash bits, no
// entry 6 or more slots from home -- and the reference's std::sort is replayed on that order. This skips the
// replay of the map itself (three or four table incarnations), which is most of hs_rank_literal's work.
#define HS_RANK_SLOT_MAXN 51
template <class Acc>
HS_HD int hs_rank_slotorder(const Acc& acc, int m, const HsRankLut* lut, int& k0, int& k1, unsigned& c0, unsigned& c1,
                            unsigned& c2) {
    const int n = m + 3;
    if (m < 2 || n > HS_RANK_SLOT_MAXN) return 1;
    const int level = hs_rank_level(n);
    const uint8_t* home = lut->home[level];
    const uint8_t* hbits = lut->hbits[level];
    uint32_t e[HS_RANK_SLOT_MAXN];  // (home << 5 | 31 - hash bits) << 8 | key, insertion-sorted ascending
    for (int i = 0; i < n; i++) {
        const int key = i < 3 ? i : acc.key(i - 3);  // the dummy keys 0,1,2 (:492-494) sit in the table too
        const uint32_t v = ((uint32_t)(home[key] * 32 + 31 - hbits[key]) << 8) | (uint32_t)key;
        int j = i;
        while (j > 0 && e[j - 1] > v) {
            e[j] = e[j - 1];
            j--;
        }
        e[j] = v;
    }
    int next_free = 0;
    for (int i = 0; i < n; i++) {
        const int rk = (int)(e[i] >> 8), h = rk >> 5;
        if (i > 0 && rk == (int)(e[i - 1] >> 8)) return 1;  // same bucket, same hash bits: history decides
        const int slot = next_free > h ? next_free : h;
        if (slot - h >= 6) return 1;  // the info bytes may have been widened (try_increase_info)
        next_free = slot + 1;
    }
    for (int i = 0; i < n; i++) {
        const int key = (int)(e[i] & 0xffu);
        const unsigned cnt = key >= 33 ? acc.count(key) : 0u;
        e[i] = (cnt << 8) | (unsigned)key;
    }
    hs_kc_std_sort(e, n);
    k0 = e[0] & 0xff;
    k1 = e[1] & 0xff;
    c0 = e[0] >> 8;
    c1 = e[1] >> 8;
    c2 = e[2] >> 8;
    return 0;
}

// host only: builds the tables by replaying the reference behaviour
struct HsRankOneKey {
    int code;
    HS_HD int key(int) const { return code; }
    HS_HD unsigned count(int) const { return 1; }
};
inline void hs_build_rank_lut(HsRankLut& lut) {
    for (int level = 0; level < 4; level++) {
        HsRhTable t;
        hs_rh_new(t);
        t.mask = (8u << level) - 1;
        for (int l = 0; l < level; l++) t.mult += 0xc4ceb9fe1a85ec54ull;
        for (int key = 0; key < 160; key++) {
            uint32_t idx, info;
            hs_rh_key_to_idx(t, (uint8_t)key, idx, info);
            lut.home[level][key] = (uint8_t)idx;
            lut.hbits[level][key] = (uint8_t)(info - t.info_inc);
            lut.tie[level][key] = (uint16_t)(((255u - idx) << 8) | (unsigned)key);
        }
    }
    for (int i = 0; i < 128; i++) {
        int k0 = 0, k1 = 0;
        unsigned c0, c1, c2;
        HsRankOneKey acc{33 + i};
        hs_rank_literal(acc, i < 125 ? 1 : 0, k0, k1, c0, c1, c2);
        lut.single[i] = (uint8_t)k1;
        if (i == 127) {
            lut.empty[0] = (uint8_t)k0;
            lut.empty[1] = (uint8_t)k1;
        }
    }
}
