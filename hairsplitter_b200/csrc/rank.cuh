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
