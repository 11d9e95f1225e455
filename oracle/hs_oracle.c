/* TEST INFRASTRUCTURE ONLY -- see hs_oracle.h. Plain-C restatement of the reference's hot path.
 * Every function cites the reference lines it follows (paths relative to /root/reference). */
#include "hs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Pileup: generate_msa, src/call_variants.cpp:163-364.
 * ------------------------------------------------------------------------------------------- */
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

/* 3-mer code: '!' + 5*i(c-2) + i(c-1) + 25*i(c0), alphabet "ACGT-" (:62,238,287). */
static inline uint8_t three_mer(int c2, int c1, int c0) { return (uint8_t)(33 + 5 * c2 + c1 + 25 * c0); }

static inline int oriented_base(const uint8_t* rb, int64_t len, int strand, int64_t t) {
    /* strand==0: the reference aligns sequence_.reverse_complement() (:108-115); revcomp of the
     * 2-bit code is reversal + bitwise NOT (sequence.cpp:54-65) */
    if (t >= len) return 0; /* malformed CIGAR; the reference would read past the string (UB) */
    return strand ? rb[t] : 3 - rb[len - 1 - t];
}

int64_t hso_pileup(const uint8_t* contig, int32_t L, int32_t n_reads, const uint8_t* read_bases,
                   const int64_t* read_off, const uint32_t* cigar, const int64_t* cigar_off,
                   const int32_t* start, const uint8_t* strand, int64_t cell_capacity, int64_t* col_off,
                   uint32_t* read_idx, uint8_t* code, int64_t* stats, int32_t* read_end) {
    int64_t* cursor = (int64_t*)calloc((size_t)L + 1, sizeof(int64_t));
    int64_t total = 0, dist = 0, alen = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int64_t acc = 0;
            for (int32_t q = 0; q < L; q++) {
                col_off[q] = acc;
                acc += cursor[q];
                cursor[q] = col_off[q];
            }
            col_off[L] = acc;
            total = acc;
            if (cell_capacity < total) break; /* count-only call */
        }
        for (int32_t n = 0; n < n_reads; n++) {
            const uint8_t* rb = read_bases + read_off[n];
            int64_t rlen = read_off[n + 1] - read_off[n];
            int64_t q = start[n]; /* indexQuery (:189) */
            int64_t t = 0;        /* indexTarget (:190) */
            int c3 = 0, c2 = 1, c1 = 2; /* context 'A','C','G' (:212-214) */
            for (int64_t k = cigar_off[n]; k < cigar_off[n + 1]; k++) {
                int op = (int)(cigar[k] & 15);
                int64_t len = cigar[k] >> 4;
                for (int64_t l = 0; l < len; l++) {
                    if (q >= L) break; /* :217 -- nothing at all happens once the contig end is reached */
                    if (op == OP_M || op == OP_EQ || op == OP_X) { /* :226-268 */
                        int b = oriented_base(rb, rlen, strand[n], t);
                        c3 = c2; c2 = c1; c1 = b;
                        if (pass == 0) {
                            cursor[q]++;
                            if (b != contig[q]) dist++;
                            alen++;
                        } else {
                            read_idx[cursor[q]] = (uint32_t)n;
                            code[cursor[q]] = three_mer(c3, c2, c1);
                            cursor[q]++;
                        }
                        q++; t++;
                    } else if (op == OP_S || op == OP_H) { /* :269-273 */
                        t++;
                    } else if (op == OP_D) { /* :274-310 */
                        c3 = c2; c2 = c1; c1 = 4;
                        if (pass == 0) {
                            cursor[q]++;
                            dist++; alen++;
                        } else {
                            read_idx[cursor[q]] = (uint32_t)n;
                            code[cursor[q]] = three_mer(c3, c2, c1);
                            cursor[q]++;
                        }
                        q++;
                    } else if (op == OP_I) { /* :311-342 */
                        int b = oriented_base(rb, rlen, strand[n], t);
                        c3 = c2; c2 = c1; c1 = b;
                        t++;
                        if (pass == 0) { dist++; alen++; }
                    }
                    /* any other letter (N, P) matches no branch of the reference loop */
                }
            }
            if (pass == 0 && read_end) read_end[n] = (int32_t)q; /* :354 */
        }
    }
    if (stats) { stats[0] = dist; stats[1] = alen; }
    free(cursor);
    return total;
}

float hso_mean_distance(int64_t distance_sum, int64_t aligned_sum) {
    /* float totalDistance incremented by 1 saturates at 2^24 (:67,255); double length starts at 1 (:68) */
    float totalDistance = (float)(distance_sum > 16777216 ? 16777216 : distance_sum);
    double totalLength = 1.0 + (double)aligned_sum;
    return (float)(totalDistance / totalLength); /* :434, float/double -> double -> float return */
}

void hso_ref_codes(const uint8_t* contig, int32_t L, uint8_t* out) {
    int c3 = 0, c2 = 1, c1 = 2; /* :367-369 */
    for (int32_t i = 0; i < L; i++) {
        c3 = c2; c2 = c1; c1 = contig[i];
        out[i] = three_mer(c3, c2, c1);
    }
}

/* ---------------------------------------------------------------------------------------------
 * robin_hood flat table, behaviour restated from src/robin_hood.h:
 *   hash_int :749-760, keyToIdx :1349-1361, insertKeyPrepareEmptySpot :2332-2382, shiftUp :1377-1396,
 *   insert_move :1451-1494, try_increase_info :2385-2411, increase_size :2414-2443,
 *   rehashPowerOfTwo :2203-2237, initData :2304-2325, iteration = slot order (:1307-1330).
 * Keys are unsigned char, so at most 256 entries and tables up to 512 slots (+ overflow buffer).
 * ------------------------------------------------------------------------------------------- */
#define RH_MAXSLOTS (1024 + 256 + 8)
typedef struct {
    uint8_t info[RH_MAXSLOTS];
    uint8_t key[RH_MAXSLOTS];
    uint64_t mult;
    uint32_t mask; /* buckets - 1; 0 = unallocated */
    uint32_t n, max_allowed;
    uint32_t info_inc, info_shift;
} rh_t;

static uint32_t rh_max_allowed(uint32_t buckets) { return buckets * 80u / 100u; }
static uint32_t rh_with_buffer(uint32_t buckets) {
    uint32_t m = rh_max_allowed(buckets);
    return buckets + (m < 0xFFu ? m : 0xFFu);
}
static void rh_init_data(rh_t* t, uint32_t buckets) {
    t->n = 0;
    t->mask = buckets - 1;
    t->max_allowed = rh_max_allowed(buckets);
    uint32_t nb = rh_with_buffer(buckets);
    memset(t->info, 0, nb + 8);
    t->info[nb] = 1; /* sentinel */
    t->info_inc = 32;
    t->info_shift = 0;
}
static void rh_key_to_idx(const rh_t* t, uint8_t key, uint32_t* idx, uint32_t* info) {
    uint64_t h = (uint64_t)key;
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= t->mult;
    h ^= h >> 33;
    *info = t->info_inc + (uint32_t)((h & 31u) >> t->info_shift);
    *idx = (uint32_t)(h >> 5) & t->mask;
}
static void rh_shift_up(rh_t* t, uint32_t start, uint32_t ins) {
    for (uint32_t i = start; i != ins; i--) t->key[i] = t->key[i - 1];
    for (uint32_t i = start; i != ins; i--) {
        t->info[i] = (uint8_t)(t->info[i - 1] + t->info_inc);
        if ((uint32_t)t->info[i] + t->info_inc > 0xFF) t->max_allowed = 0;
    }
}
static int rh_try_increase_info(rh_t* t) {
    if (t->info_inc <= 2) return 0;
    t->info_inc >>= 1;
    t->info_shift++;
    uint32_t nb = rh_with_buffer(t->mask + 1);
    for (uint32_t i = 0; i < nb + 8; i++) t->info[i] = (t->info[i] >> 1) & 0x7f; /* 8 bytes at a time in the original */
    t->info[nb] = 1;
    t->max_allowed = rh_max_allowed(t->mask + 1);
    return 1;
}
static void rh_insert_move(rh_t* t, uint8_t key) {
    if (t->max_allowed == 0 && !rh_try_increase_info(t)) abort();
    uint32_t idx, info;
    rh_key_to_idx(t, key, &idx, &info);
    while (info <= t->info[idx]) { idx++; info += t->info_inc; }
    uint32_t ins = idx, ins_info = info & 0xFF;
    if (ins_info + t->info_inc > 0xFF) t->max_allowed = 0;
    while (t->info[idx] != 0) idx++;
    if (idx != ins) rh_shift_up(t, idx, ins);
    t->key[ins] = key;
    t->info[ins] = (uint8_t)ins_info;
    t->n++;
}
static void rh_rehash(rh_t* t, uint32_t buckets) {
    uint8_t oi[RH_MAXSLOTS], ok[RH_MAXSLOTS];
    uint32_t old_nb = rh_with_buffer(t->mask + 1);
    memcpy(oi, t->info, old_nb);
    memcpy(ok, t->key, old_nb);
    rh_init_data(t, buckets);
    for (uint32_t i = 0; i < old_nb; i++)
        if (oi[i] != 0) rh_insert_move(t, ok[i]);
}
static int rh_increase_size(rh_t* t) {
    if (t->mask == 0) { rh_init_data(t, 8); return 1; }
    uint32_t max_allowed = rh_max_allowed(t->mask + 1);
    if (t->n < max_allowed && rh_try_increase_info(t)) return 1;
    t->mult += 0xc4ceb9fe1a85ec54ULL;
    if (t->n * 2 < rh_max_allowed(t->mask + 1)) rh_rehash(t, t->mask + 1);
    else rh_rehash(t, (t->mask + 1) * 2);
    return 1;
}
static void rh_new(rh_t* t) {
    t->mult = 0xc4ceb9fe1a85ec53ULL;
    t->mask = 0; t->n = 0; t->max_allowed = 0; t->info_inc = 32; t->info_shift = 0;
    memset(t->info, 0, 16); /* mInfo aliases &mMask (== 0) while unallocated */
}
/* insert a key known to be absent (find()==end() then operator[]) */
static void rh_insert(rh_t* t, uint8_t key) {
    for (int attempt = 0; attempt < 256; attempt++) {
        uint32_t idx, info;
        rh_key_to_idx(t, key, &idx, &info);
        if (t->mask != 0) {
            while (info < t->info[idx]) { idx++; info += t->info_inc; }
            while (info == t->info[idx]) { /* key differs by precondition */ idx++; info += t->info_inc; }
        }
        if (t->n >= t->max_allowed) { rh_increase_size(t); continue; }
        uint32_t ins = idx, ins_info = info;
        if (ins_info + t->info_inc > 0xFF) t->max_allowed = 0;
        while (t->info[idx] != 0) idx++;
        if (idx != ins) rh_shift_up(t, idx, ins);
        t->key[ins] = key;
        t->info[ins] = (uint8_t)ins_info;
        t->n++;
        return;
    }
    abort();
}
static int rh_iterate(const rh_t* t, uint8_t* out) {
    if (t->mask == 0) return 0;
    uint32_t nb = rh_with_buffer(t->mask + 1);
    int k = 0;
    for (uint32_t i = 0; i < nb; i++)
        if (t->info[i] != 0) out[k++] = t->key[i];
    return k;
}

int hso_rh_order(const uint8_t* keys, int n, uint8_t* out) {
    rh_t t;
    rh_new(&t);
    for (int i = 0; i < n; i++) rh_insert(&t, keys[i]);
    return rh_iterate(&t, out);
}

/* ---------------------------------------------------------------------------------------------
 * std::sort of libstdc++ (bits/stl_algo.h: __introsort_loop, __unguarded_partition_pivot,
 * __move_median_to_first, __final_insertion_sort, __heap_select/__sort_heap), restated for
 * (key,count) pairs with comp(a,b) = a.count > b.count (src/call_variants.cpp:501).
 * ------------------------------------------------------------------------------------------- */
typedef struct { uint8_t k; int32_t c; } kc_t;
#define COMP(a, b) ((a).c > (b).c)
static void kc_swap(kc_t* a, kc_t* b) { kc_t t = *a; *a = *b; *b = t; }

static void adjust_heap(kc_t* first, long hole, long len, kc_t value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (COMP(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    /* __push_heap */
    long parent = (hole - 1) / 2;
    while (hole > top && COMP(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void heap_sort_all(kc_t* first, long len) { /* partial_sort(first,last,last) */
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            kc_t v = first[parent];
            adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    /* __heap_select's loop over [middle,last) is empty; then __sort_heap */
    for (long last = len; last > 1;) {
        --last;
        kc_t v = first[last];
        first[last] = first[0];
        adjust_heap(first, 0, last, v);
    }
}
static void move_median_to_first(kc_t* result, kc_t* a, kc_t* b, kc_t* c) {
    if (COMP(*a, *b)) {
        if (COMP(*b, *c)) kc_swap(result, b);
        else if (COMP(*a, *c)) kc_swap(result, c);
        else kc_swap(result, a);
    } else if (COMP(*a, *c)) kc_swap(result, a);
    else if (COMP(*b, *c)) kc_swap(result, c);
    else kc_swap(result, b);
}
static kc_t* unguarded_partition(kc_t* first, kc_t* last, kc_t* pivot) {
    for (;;) {
        while (COMP(*first, *pivot)) ++first;
        --last;
        while (COMP(*pivot, *last)) --last;
        if (!(first < last)) return first;
        kc_swap(first, last);
        ++first;
    }
}
static void introsort_loop(kc_t* first, kc_t* last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) { heap_sort_all(first, last - first); return; }
        --depth_limit;
        kc_t* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        kc_t* cut = unguarded_partition(first + 1, last, first);
        introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}
static void unguarded_linear_insert(kc_t* last) {
    kc_t val = *last;
    kc_t* next = last - 1;
    while (COMP(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void insertion_sort(kc_t* first, kc_t* last) {
    if (first == last) return;
    for (kc_t* i = first + 1; i != last; ++i) {
        if (COMP(*i, *first)) {
            kc_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(kc_t));
            *first = val;
        } else unguarded_linear_insert(i);
    }
}
static void std_sort(kc_t* first, kc_t* last) {
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((1L << (lg + 1)) <= n) lg++;
    introsort_loop(first, last, 2 * lg);
    if (last - first > 16) {
        insertion_sort(first, first + 16);
        for (kc_t* i = first + 16; i != last; ++i) unguarded_linear_insert(i);
    } else insertion_sort(first, last);
}

void hso_sort_desc(uint8_t* keys, int32_t* counts, int n) {
    kc_t v[256 + 8];
    for (int i = 0; i < n; i++) { v[i].k = keys[i]; v[i].c = counts[i]; }
    std_sort(v, v + n);
    for (int i = 0; i < n; i++) { keys[i] = v[i].k; counts[i] = v[i].c; }
}

/* ---------------------------------------------------------------------------------------------
 * Column ranking + SNP predicate: call_variants, src/call_variants.cpp:447-567.
 * ------------------------------------------------------------------------------------------- */
static int rank_column(const uint8_t* codes, int n, kc_t* sorted) {
    /* :477-494: histogram in a robin_hood map; keys enter in order of first occurrence going down
     * the column (loop counter is a `short`, :479 -- columns deeper than 32767 are not reachable),
     * then the dummy keys 0,1,2. A cell equal to ' ' is registered but not counted (:484). */
    int32_t cnt[256];
    uint8_t order[256 + 3];
    int m = 0;
    memset(cnt, 0xff, sizeof(cnt));
    for (int i = 0; i < n; i++) {
        uint8_t b = codes[i];
        if (cnt[b] < 0) { cnt[b] = 0; order[m++] = b; }
        if (b != ' ') cnt[b]++;
    }
    for (uint8_t d = 0; d < 3; d++)
        if (cnt[d] < 0) { cnt[d] = 0; order[m++] = d; }
    rh_t t;
    rh_new(&t);
    for (int i = 0; i < m; i++) rh_insert(&t, order[i]);
    uint8_t it[256 + 8];
    int k = rh_iterate(&t, it);
    for (int i = 0; i < k; i++) { sorted[i].k = it[i]; sorted[i].c = cnt[it[i]]; } /* :497-500 */
    std_sort(sorted, sorted + k);                                                     /* :501 */
    return k;
}

/* The counting loop of call_variants runs on a `short` index (:479): at n = 32767 the increment wraps to -32768, which
 * compared with the unsigned size ends the loop -- only the first 32768 cells of a column are ever counted (checked
 * against the compiled reference on a 40000-deep column, tests/test_oracle.py). */
#define HSO_CELLS_COUNTED 32768

void hso_column_rank(const uint8_t* codes, int n, int32_t* out) {
    kc_t s[256 + 8];
    if (n > HSO_CELLS_COUNTED) n = HSO_CELLS_COUNTED;
    rank_column(codes, n, s);
    out[0] = s[0].k; out[1] = s[1].k; out[2] = s[0].c; out[3] = s[1].c; out[4] = s[2].c;
}

int32_t hso_call_variants(const int64_t* col_off, const uint8_t* code, int32_t L, float mean_error,
                          float auto_threshold, uint8_t* ref_base, uint8_t* second_base,
                          int32_t* suspect_pos, uint8_t* suspect_is_auto, int32_t suspect_capacity,
                          int64_t* depth_sum) {
    int min_reads = 5;               /* :463 */
    if (mean_error < 0.015) min_reads = 3; /* :464-466 (float < double literal) */
    int64_t depth = 0;
    int posoflastsnp = -5;           /* :470 */
    int32_t ns = 0;
    kc_t s[256 + 8];
    for (int32_t pos = 0; pos < L; pos++) {
        const uint8_t* col = code + col_off[pos];
        int n = (int)(col_off[pos + 1] - col_off[pos]);
        if (n > HSO_CELLS_COUNTED) n = HSO_CELLS_COUNTED; /* :479, see above */
        rank_column(col, n, s);
        for (int i = 0; i < n; i++) depth += (col[i] != ' ');
        ref_base[pos] = s[0].k;      /* :503-507 */
        second_base[pos] = s[1].k;
        int k0 = s[0].k, k1 = s[1].k;
        if (s[1].c > min_reads && (s[1].c > s[2].c * 5 || min_reads == 2) && k0 % 5 != k1 % 5 &&
            ((k1 - '!') % 5 != 4 || (k1 / 5 % 5 != k0 % 5 && k1 / 25 % 5 != k0 % 5)) &&
            pos - posoflastsnp > 5) { /* :525-529 */
            if (ns < suspect_capacity) {
                suspect_pos[ns] = pos;
                suspect_is_auto[ns] = (float)s[1].c > auto_threshold * (float)s[0].c; /* :531 int > float*int */
            }
            ns++;
            posoflastsnp = pos;
        }
    }
    if (depth_sum) *depth_sum = depth;
    return ns;
}

/* ---------------------------------------------------------------------------------------------
 * distance(Partition&, Column&, char ref_base): src/call_variants.cpp:778-967.
 * ------------------------------------------------------------------------------------------- */
void hso_distance(int32_t np, const int32_t* p_idx, const int16_t* p_state, const int32_t* p_more,
                  const int32_t* p_less, int32_t nc, const uint32_t* c_idx, const uint8_t* c_code,
                  int32_t ref_base_in, int32_t* out) {
    uint8_t ref_base = (uint8_t)ref_base_in;
    int32_t cnt[256];
    uint8_t order[256 + 1];
    int m = 0;
    memset(cnt, 0xff, sizeof(cnt));
    int number_of_bases = 0;
    /* :798-815 merge-join; histogram of the column's codes over reads of the partition with state != -2 */
    int32_t n1 = 0;
    for (int32_t n2 = 0; n2 < nc; n2++) {
        while (n1 < np && (int64_t)p_idx[n1] < (int64_t)c_idx[n2]) n1++;
        if (n1 >= np) break;
        if ((int64_t)p_idx[n1] == (int64_t)c_idx[n2] && p_state[n1] != -2) {
            uint8_t b = c_code[n2];
            if (cnt[b] < 0) { cnt[b] = 0; order[m++] = b; }
            number_of_bases++;
            cnt[b]++;
        }
    }
    memset(out, 0, 10 * sizeof(int32_t));
    if (number_of_bases == 0) return; /* :817-828, augmented = false */

    /* :832-844: content2[ref_base] inserts the key when absent, then the map is iterated; the first
     * strictly greater count wins, so ties resolve by the table's iteration order */
    if (cnt[ref_base] < 0) { cnt[ref_base] = 0; order[m++] = ref_base; }
    rh_t t;
    rh_new(&t);
    for (int i = 0; i < m; i++) rh_insert(&t, order[i]);
    uint8_t it[256 + 8];
    int k = rh_iterate(&t, it);
    /* `ref_base != c.first` compares a (signed) char with an unsigned char (:838): for codes >= 128
     * the test is always true, so the reference code itself can become "secondFrequent" */
    int ref_as_char = (int)(signed char)ref_base;
    uint8_t second = ' ';
    int max2 = -1;
    for (int i = 0; i < k; i++) {
        if (ref_as_char != (int)it[i] && cnt[it[i]] > max2) { second = it[i]; max2 = cnt[it[i]]; }
    }
    /* :889-949 second merge-join: 2x2 table over states +1/-1 */
    int m00 = 0, m01 = 0, m10 = 0, m11 = 0, s00 = 0, s01 = 0, s10 = 0, s11 = 0;
    int32_t i1 = 0, i2 = 0;
    while (i1 < np && i2 < nc) {
        if ((int64_t)p_idx[i1] == (int64_t)c_idx[i2]) {
            int solid = (p_less[i1] <= 1 && p_more[i1] >= 3);
            if (c_code[i2] == ref_base) {
                if (p_state[i1] == 1) { m11++; s11 += solid; }
                else if (p_state[i1] == -1) { m01++; s01 += solid; }
            } else if (c_code[i2] == second) {
                if (p_state[i1] == 1) { m10++; s10 += solid; }
                else if (p_state[i1] == -1) { m00++; s00 += solid; }
            }
            i1++; i2++;
        } else if ((int64_t)c_idx[i2] > (int64_t)p_idx[i1]) i1++;
        else i2++;
    }
    out[0] = m00; out[1] = m01; out[2] = m10; out[3] = m11;
    out[4] = s00; out[5] = s01; out[6] = s10; out[7] = s11;
    out[8] = second; out[9] = 1;
}

float hso_chi_square(int32_t n00, int32_t n01, int32_t n10, int32_t n11) {
    /* :1135-1163 verbatim arithmetic: float margins, float expected counts, pow() in double */
    int n = n00 + n01 + n10 + n11;
    if (n == 0) return 0;
    float pmax1 = (float)(n10 + n11) / n;
    float pmax2 = (float)(n01 + n11) / n;
    if (pmax1 * (1 - pmax1) == 0 && pmax2 * (1 - pmax2) == 0) return -1;
    if (pmax1 * pmax2 * (1 - pmax1) * (1 - pmax2) == 0) return 0;
    float res;
    res = (float)(pow((n00 - (1 - pmax1) * (1 - pmax2) * n), 2) / ((1 - pmax1) * (1 - pmax2) * n) +
                  pow((n01 - (1 - pmax1) * pmax2 * n), 2) / ((1 - pmax1) * pmax2 * n) +
                  pow((n10 - pmax1 * (1 - pmax2) * n), 2) / (pmax1 * (1 - pmax2) * n) +
                  pow((n11 - pmax1 * pmax2 * n), 2) / (pmax1 * pmax2 * n));
    return res;
}

int hso_rescue_prefilter(int32_t ref_base, int32_t second_base) {
    /* :751-752, raw unsigned char arithmetic */
    return ref_base % 5 != second_base % 5 &&
           ((second_base - '!') % 5 != 4 ||
            (second_base / 5 % 5 != ref_base % 5 && second_base / 25 % 5 != ref_base % 5));
}

/* Loops 3 and 4 of keep_only_robust_variants, src/call_variants.cpp:718-764, given the final
 * partitions (loops 1-2 are sequential host logic and are not restated here). */
int32_t hso_robust_filter(int32_t L, const int64_t* col_off, const uint32_t* read_idx, const uint8_t* code,
                          const uint8_t* ref_base, const uint8_t* second_base, int32_t n_parts,
                          const int64_t* part_off, const int32_t* p_idx, const int16_t* p_state,
                          const int32_t* p_more, const int32_t* p_less, int32_t n_suspects,
                          const int32_t* suspect_pos, int32_t* kept) {
    if (n_parts == 0) return 0; /* :640-642 */
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_suspects + 1));
    int32_t ntmp = 0, nk = 0, d[10];
    for (int32_t i = 0; i < n_suspects; i++) { /* :721-738 */
        int32_t pos = suspect_pos[i];
        int32_t nc = (int32_t)(col_off[pos + 1] - col_off[pos]);
        for (int32_t p = 0; p < n_parts; p++) {
            int64_t a = part_off[p];
            hso_distance((int32_t)(part_off[p + 1] - a), p_idx + a, p_state + a, p_more + a, p_less + a, nc,
                         read_idx + col_off[pos], code + col_off[pos], ref_base[pos], d);
            float chisqu = hso_chi_square(d[0], d[1], d[2], d[3]);
            if (d[0] + d[1] + d[2] + d[3] > 0.5 * nc && chisqu > 15) { tmp[ntmp++] = pos; break; }
        }
    }
    int32_t idx = 0;
    for (int32_t pos = 0; pos < L; pos++) { /* :745-764 */
        if (idx < ntmp && tmp[idx] == pos) { kept[nk++] = pos; idx++; }
        else if (hso_rescue_prefilter(ref_base[pos], second_base[pos])) {
            int32_t nc = (int32_t)(col_off[pos + 1] - col_off[pos]);
            for (int32_t p = 0; p < n_parts; p++) {
                int64_t a = part_off[p];
                hso_distance((int32_t)(part_off[p + 1] - a), p_idx + a, p_state + a, p_more + a, p_less + a, nc,
                             read_idx + col_off[pos], code + col_off[pos], ref_base[pos], d);
                if (hso_chi_square(d[0], d[1], d[2], d[3]) > 20.0 && d[2] + d[0] > 4 && d[1] + d[3] > 4) {
                    kept[nk++] = pos;
                    break;
                }
            }
        }
    }
    free(tmp);
    return nk;
}

/* ---------------------------------------------------------------------------------------------
 * Read x read counts: list_similarities_and_differences_between_reads3, src/separate_reads.cpp:374-433.
 * similarity = 3*A*At + R*Rt, difference = A*Rt + R*At, diagonals zeroed; A[r,s] = content==second_base,
 * R[r,s] = content==ref_base (ref tested first, :384-394).
 * ------------------------------------------------------------------------------------------- */
void hso_read_pair_counts(int32_t n_reads, int32_t n_snps, const int64_t* snp_off, const uint32_t* read_idx,
                          const uint8_t* code, const uint8_t* ref_base, const uint8_t* second_base,
                          int32_t* sim, int32_t* diff) {
    memset(sim, 0, (size_t)n_reads * n_reads * sizeof(int32_t));
    memset(diff, 0, (size_t)n_reads * n_reads * sizeof(int32_t));
    for (int32_t s = 0; s < n_snps; s++) {
        for (int64_t a = snp_off[s]; a < snp_off[s + 1]; a++) {
            int ka = code[a] == ref_base[s] ? 1 : (code[a] == second_base[s] ? 2 : 0);
            if (!ka) continue;
            for (int64_t b = snp_off[s]; b < snp_off[s + 1]; b++) {
                int kb = code[b] == ref_base[s] ? 1 : (code[b] == second_base[s] ? 2 : 0);
                if (!kb || read_idx[a] == read_idx[b]) continue;
                size_t o = (size_t)read_idx[a] * n_reads + read_idx[b];
                if (ka == kb) sim[o] += (ka == 2) ? 3 : 1;
                else diff[o] += 1;
            }
        }
    }
}

/* ---- read clipping: the CIGAR walk of modify_GFA (src/create_new_contigs.cpp:392-447) -------------------------------
 * One character of the expanded CIGAR at a time, as the reference does it. Only 'M', 'D' and 'I' move a cursor
 * (:426-435); 'S' / 'H' advance the read cursor until the start has been found and end the walk afterwards
 * (:407-415); every other letter ('=', 'X', 'N', 'P') is looked at by the two position tests and moves nothing. */
int32_t hso_clip_read(const uint32_t* ops, int64_t n_ops, int32_t pos_2_1, int32_t left_to_polish, int32_t right_to_polish,
                      int32_t* out) {
    static const char letters[] = "MIDNSHP=X";
    int posOnRead = 0, posOnCIGAR = 0;
    int posOnInterval = pos_2_1;
    int posOnReadStart = -1, posOnReadEnd = -1, posOnCIGARStart = -1, posOnCIGAREnd = -1;
    int done = 0;
    for (int64_t k = 0; k < n_ops && !done; k++) {
        const char c = letters[ops[k] & 15u];
        const uint32_t n = ops[k] >> 4;
        for (uint32_t i = 0; i < n; i++) {
            posOnCIGAR++;
            if (c == 'S' || c == 'H') {
                if (posOnReadStart != -1) {
                    posOnReadEnd = posOnRead;
                    posOnCIGAREnd = posOnCIGAR - 1;
                    done = 1;
                    break;
                }
                posOnRead++;
                continue;
            }
            if (posOnReadStart == -1 && posOnInterval >= left_to_polish) {
                posOnReadStart = posOnRead;
                posOnCIGARStart = posOnCIGAR - 1;
            }
            if (posOnReadEnd == -1 && posOnInterval == right_to_polish) {
                posOnReadEnd = posOnRead;
                posOnCIGAREnd = posOnCIGAR - 1;
                done = 1;
                break;
            }
            if (c == 'M') {
                posOnRead++;
                posOnInterval++;
            } else if (c == 'D') {
                posOnInterval++;
            } else if (c == 'I') {
                posOnRead++;
            }
        }
    }
    if (posOnReadEnd == -1) {
        posOnReadEnd = posOnRead;
        posOnCIGAREnd = posOnCIGAR;
    }
    out[0] = posOnReadStart;
    out[1] = posOnReadEnd;
    out[2] = posOnCIGARStart;
    out[3] = posOnCIGAREnd;
    if (posOnReadStart > posOnReadEnd || posOnReadStart == -1) return -2;
    return 0;
}
