// Partition x column contingency tables: distance(Partition&, Column&, char) + computeChiSquare
// (reference src/call_variants.cpp:778-967,1135-1163) and the two filtering loops of
// keep_only_robust_variants that call them for every column (:721-764, 85 % of the reference's
// HS_call_variants run time).
//
// The partition is a sparse vector over reads with state +1 / -1 / 0 / -2(masked). On the device the
// partitions of a contig are one dense byte matrix pstate[p][n] (0 = read absent or masked, 1 = +1,
// 2 = -1, 3 = state 0, bit 2 = "solid": less <= 1 && more >= 3), so the merge-joins of the reference
// become direct lookups. For a (column, partition) pair:
//   pass 1  histogram of the column's codes over the partition's reads -> alternative allele = most
//           frequent code different from ref_base, ties broken by robin_hood iteration order (rank.cuh);
//   pass 2  the 2x2 table n11/n01/n10/n00 (+ solid variants) over reads with state +-1.
#include <cstdlib>
#include <cstring>
#include <vector>

#include <atomic>

#include <chrono>

#include "common.cuh"
#include "rank.cuh"

#define CT_ROWS 256  // rows (reads) staged per batch
#define CT_NPA 32    // active partitions staged per chunk
#define CT_SSTRIDE (CT_ROWS + 4)  // row stride of the staged states: the 32 partitions of a chunk fall in different banks

// computeChiSquare (:1135-1163) with the reference's exact mix of float and double arithmetic
// (x86-64 SSE2, no FMA contraction): float margins and expected counts, squares and quotients in
// double, sum in double, result rounded to float.
__device__ __forceinline__ float hs_chi_square(int n00, int n01, int n10, int n11) {
    const int n = n00 + n01 + n10 + n11;
    if (n == 0) return 0.f;
    const float fn = (float)n;
    const float p1 = __fdiv_rn((float)(n10 + n11), fn);
    const float p2 = __fdiv_rn((float)(n01 + n11), fn);
    const float q1 = __fsub_rn(1.f, p1), q2 = __fsub_rn(1.f, p2);
    if (__fmul_rn(p1, q1) == 0.f && __fmul_rn(p2, q2) == 0.f) return -1.f;
    if (__fmul_rn(__fmul_rn(__fmul_rn(p1, p2), q1), q2) == 0.f) return 0.f;
    const float e00 = __fmul_rn(__fmul_rn(q1, q2), fn);
    const float e01 = __fmul_rn(__fmul_rn(q1, p2), fn);
    const float e10 = __fmul_rn(__fmul_rn(p1, q2), fn);
    const float e11 = __fmul_rn(__fmul_rn(p1, p2), fn);
    const double d00 = (double)__fsub_rn((float)n00, e00), d01 = (double)__fsub_rn((float)n01, e01);
    const double d10 = (double)__fsub_rn((float)n10, e10), d11 = (double)__fsub_rn((float)n11, e11);
    double s = __ddiv_rn(__dmul_rn(d00, d00), (double)e00);
    s = __dadd_rn(s, __ddiv_rn(__dmul_rn(d01, d01), (double)e01));
    s = __dadd_rn(s, __ddiv_rn(__dmul_rn(d10, d10), (double)e10));
    s = __dadd_rn(s, __ddiv_rn(__dmul_rn(d11, d11), (double)e11));
    return __double2float_rn(s);
}

// secondFrequent of :832-844. `order` holds the codes (minus 33) in order of first appearance among the
// partition's reads, hist their counts. The reference compares `char ref_base != unsigned char key`
// (:838), which is always true for codes >= 128, so there the reference code itself competes.
__device__ int hs_select_alt(const uint8_t* s_order, const uint16_t* s_hist, int stride, int tid, int m, int ref,
                             const HsRankLut* __restrict__ lut) {
    const bool ref_excluded = ref < 128;
    int max2 = -1, alt = ' ', ties = 0;
    bool ref_seen = false;
    for (int k = 0; k < m; k++) {
        const int key = s_order[k * stride + tid] + HS_CODE0;
        if (key == ref) ref_seen = true;
        if (ref_excluded && key == ref) continue;
        const int cnt = s_hist[(key - HS_CODE0) * stride + tid];
        if (cnt > max2) { max2 = cnt; alt = key; ties = 1; }
        else if (cnt == max2) ties++;
    }
    if (ties <= 1) return alt;
    // Several codes share the maximum: the first one in the map's iteration order wins. A flat robin-hood table
    // keeps its entries sorted by home bucket (rank.cuh), so the tied key with the smallest home bucket in the
    // table's final incarnation wins whenever no other tied key shares that bucket and the layout never forced
    // an early growth (no entry 6 or more slots from home, as in hs_rank_hashbits). Otherwise: literal replay.
    const int n = m + (ref_seen ? 0 : 1);  // content2[ref_base] creates the entry (:833)
    if (lut && n <= 51) {
        const int level = hs_rank_level(n);
        const uint8_t* __restrict__ home = lut->home[level];
        unsigned long long occ[4] = {0, 0, 0, 0};  // 4-bit occupancy counters of up to 64 buckets
        int best = 1 << 30, nbest = 0;
        for (int k = -1; k < m; k++) {
            if (k < 0 && ref_seen) continue;
            const int key = k < 0 ? ref : s_order[k * stride + tid] + HS_CODE0;
            const int h = __ldg(home + key);
            occ[h >> 4] += 1ull << (4 * (h & 15));
            if (k < 0 || (ref_excluded && key == ref)) continue;
            if ((int)s_hist[(key - HS_CODE0) * stride + tid] != max2) continue;
            if (h < (best >> 8)) { best = (h << 8) | key; nbest = 1; }
            else if (h == (best >> 8)) nbest++;
        }
        int next_free = 0, maxd = 0;
        const int nb = 8 << level;
        for (int b = 0; b < nb; b++) {
            const int cnt = (int)((occ[b >> 4] >> (4 * (b & 15))) & 15ull);
            if (next_free < b) next_free = b;
            next_free += cnt;
            if (cnt && next_free - 1 - b > maxd) maxd = next_free - 1 - b;
        }
        if (nbest == 1 && maxd < 6) return best & 0xff;
    }
    HsRhTable t;
    hs_rh_new(t);
    for (int k = 0; k < m; k++) hs_rh_insert(t, (uint8_t)(s_order[k * stride + tid] + HS_CODE0));
    if (!ref_seen) hs_rh_insert(t, (uint8_t)ref);
    uint8_t it[HS_RH_MAXKEYS];
    const int nk = hs_rh_iterate(t, it);
    for (int i = 0; i < nk; i++) {
        const int key = it[i];
        if (ref_excluded && key == ref) continue;
        if (key < HS_CODE0) continue;
        if ((int)s_hist[(key - HS_CODE0) * stride + tid] == max2) return key;
    }
    return alt;
}

// ---- kernel A: full tables for listed columns x all partitions -----------------------------------
struct TablesArgs {
    int n_cols;
    const int32_t* pos;
    int64_t tile0, read0, g0;
    int n_reads, n_parts;
    const uint8_t* pstate;  // [n_parts][n_reads]
    const int64_t* tile_off;
    const int32_t* tile_reads;
    const int32_t* read_start;
    const int32_t* read_end;
    const int64_t* row_base;
    const uint8_t* codes;
    const uint8_t* k0;
    const HsRankLut* lut;
    hsgpu_distance* out;
};

__global__ void __launch_bounds__(128) partition_tables_kernel(TablesArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint16_t* s_hist = reinterpret_cast<uint16_t*>(smem);                       // [125][128]
    uint8_t* s_order = smem + HS_NCODES * 128 * 2;                              // [125][128]
    int32_t* s_n = reinterpret_cast<int32_t*>(smem + HS_NCODES * 128 * 3);      // [CT_ROWS]
    uint8_t* s_code = smem + HS_NCODES * 128 * 3 + CT_ROWS * 4;                 // [CT_ROWS]
    __shared__ int s_cnt;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = a.pos[blockIdx.x];
    const int ref = a.k0[a.g0 + q];
    const int64_t tile = a.tile0 + q / HS_TILE;
    const int64_t l0 = a.tile_off[tile], l1 = a.tile_off[tile + 1];
    for (int i = tid; i < HS_NCODES * 128 * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(s_hist)[i] = 0;

    for (int pb = 0; pb < a.n_parts; pb += 128) {
        const int p = pb + tid;
        const bool valid = p < a.n_parts;
        const uint8_t* __restrict__ ps = a.pstate + (int64_t)(valid ? p : 0) * a.n_reads;
        int m = 0, nb = 0, alt = ' ';
        int n00 = 0, n01 = 0, n10 = 0, n11 = 0, s00 = 0, s01 = 0, s10 = 0, s11 = 0;
        for (int pass = 0; pass < 2; pass++) {
            for (int64_t lb = l0; lb < l1; lb += CT_ROWS) {
                // gather the cells of this column among reads lb..lb+CT_ROWS of the tile list (ordered)
                __syncthreads();
                if (tid == 0) s_cnt = 0;
                __syncthreads();
                for (int64_t sb = lb; sb < min(l1, lb + (int64_t)CT_ROWS); sb += 128) {
                    // 4 warps compact in order: warp w handles entries sb+32w..; sequential over warps via s_cnt
                    for (int w = 0; w < 4; w++) {
                        if (wid == w) {
                            const int64_t l = sb + 32 * w + lane;
                            bool hit = false;
                            int32_t r = 0;
                            if (l < l1 && l < lb + CT_ROWS) {
                                r = a.tile_reads[l];
                                hit = a.read_start[r] <= q && q < a.read_end[r];
                            }
                            const unsigned mk = __ballot_sync(0xffffffffu, hit);
                            const int base = s_cnt;
                            if (hit) {
                                const int o = base + __popc(mk & ((1u << lane) - 1u));
                                s_n[o] = (int32_t)(r - a.read0);
                                s_code[o] = a.codes[a.row_base[r] + q];
                            }
                            __syncwarp();
                            if (lane == 0) s_cnt = base + __popc(mk);
                        }
                        __syncthreads();
                    }
                }
                const int ncell = s_cnt;
                if (valid) {
                    if (pass == 0) {
                        for (int i = 0; i < ncell; i++) {
                            if (ps[s_n[i]] & 3) {
                                const int idx = s_code[i] - HS_CODE0;
                                const unsigned c = s_hist[idx * 128 + tid];
                                if (c == 0) s_order[(m++) * 128 + tid] = (uint8_t)idx;
                                s_hist[idx * 128 + tid] = (uint16_t)(c + 1);
                                nb++;
                            }
                        }
                    } else if (nb > 0) {
                        for (int i = 0; i < ncell; i++) {
                            const int st = ps[s_n[i]];
                            const int sg = st & 3;
                            if (sg == 1 || sg == 2) {
                                const int code = s_code[i];
                                const int solid = (st >> 2) & 1;
                                if (code == ref) {
                                    if (sg == 1) { n11++; s11 += solid; } else { n01++; s01 += solid; }
                                } else if (code == alt) {
                                    if (sg == 1) { n10++; s10 += solid; } else { n00++; s00 += solid; }
                                }
                            }
                        }
                    }
                }
            }
            if (pass == 0 && valid && nb > 0) alt = hs_select_alt(s_order, s_hist, 128, tid, m, ref, a.lut);
        }
        if (valid) {
            hsgpu_distance d;
            d.n00 = n00; d.n01 = n01; d.n10 = n10; d.n11 = n11;
            d.solid00 = s00; d.solid01 = s01; d.solid10 = s10; d.solid11 = s11;
            d.second_base = nb > 0 ? (uint8_t)alt : 0;
            d.augmented = nb > 0 ? 1 : 0;
            d.pad[0] = d.pad[1] = 0;
            d.chi_square = hs_chi_square(n00, n01, n10, n11);
            a.out[(int64_t)blockIdx.x * a.n_parts + p] = d;
            for (int k = 0; k < m; k++) s_hist[s_order[k * 128 + tid] * 128 + tid] = 0;
        }
    }
}

// ---- kernel B: loops 3 and 4 of keep_only_robust_variants for every contig of a pileup -------------
// A column can only be kept if it is in snps_in (loop 3, :721-738) or passes the rescue pre-filter with at
// least 5 carriers of its second allele (loop 4, :745-764; see filter_active_kernel) -- about 3 % of the
// columns at 60x / 10 % error. Those ACTIVE columns are compacted first; then one WARP takes one active column:
// its cells (read, code) are gathered once into shared memory, the partitions that hold one of its reads come
// from one 128-bit presence row per read, and for every such partition the warp builds the reference's 2x2
// table with ballots (the column's own majority code) and a small shared-memory histogram (everything else).
// One launch covers all contigs of the batch; the work is proportional to the active columns, not to the tiles.
#define RF_WARPS 8
#define RF_CAP 128  // cells of a column staged per warp; deeper columns (amplicons) are re-gathered per partition

// Partition states on the device: per read one row of 2-bit states, 128 partitions (8 words) per block: 0 = the read
// is not in the partition (or masked), 1 = +1, 2 = -1, 3 = 0. A column's cells load their row of the current block
// once (32 bytes per read) into shared memory; which partitions hold one of the column's reads, and every state the
// table passes need, come from there.
struct FilterDesc {     // per contig
    int64_t row_off;    // word offset of its [n_reads][pwords] state rows
    int32_t n_parts, pwords;  // pwords = 8 * ceil(n_parts / 128)
};

struct FilterArgs {
    int n_contigs;
    unsigned in_flag;  // the flag bit that marks snps_in: HS_FLAG_INLIST (caller's list) or HS_FLAG_SUSPECT (the pileup's own)
    const FilterDesc* desc;
    const uint32_t* rows;    // 2-bit partition states, all contigs
    const int64_t* col_base;
    const int64_t* tile_base;
    const int64_t* contig_read_off;
    const int64_t* tile_off;
    const int32_t* tile_reads;
    const int32_t* read_start;
    const int32_t* read_end;
    const int64_t* row_base;
    const uint8_t* codes;
    const uint8_t* k0;
    const uint8_t* k1;
    const uint8_t* flags;
    const uint32_t* depth;
    const uint32_t* counts;  // c0,c1,c2 per column (column ranking)
    const HsRankLut* lut;
    int64_t g_begin, g_end;  // global column range handled by this call
    uint32_t* active;        // compacted global ids of the active columns
    unsigned int* counters;  // [0] active columns of ordinary depth (front of `active`), [1] their work cursor, [2] kept
                             // columns (list reservation), [3] deep active columns (back of `active`), [5] their work
                             // cursor, [8..9] / [10..11] 64-bit: cells of the active columns / states read
    uint32_t* kept;          // bitmap over the columns of the batch (bit g & 31 of word g >> 5), zeroed by the caller
    uint32_t* overflow;      // columns robust_filter_lanes_kernel hands to robust_filter_kernel ([6] their number, [7] cursor)
    int from_overflow;       // robust_filter_kernel<false>: 1 = work off the overflow list instead of the front of `active`
};

__device__ __forceinline__ int rf_contig_of(const int64_t* __restrict__ col_base, int n_contigs, int64_t g) {
    int lo = 0, hi = n_contigs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(col_base + mid) <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Loop 4 keeps a column only if n10 + n00 > 4 (:756): reads that carry the alternative allele. No code but the
// column's own most frequent one (ref_base = k0) has more than c1 carriers, and when the reference's
// char/unsigned char quirk makes ref_base its own alternative the two counters stay 0 -- so a column with c1 <= 4
// can only be kept by loop 3, i.e. if it is in snps_in.
__global__ void __launch_bounds__(256) filter_active_kernel(FilterArgs a) {
    __shared__ unsigned s_cnt[2][8], s_base[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // four columns per thread: one 4-byte load of the flags (the flag array starts 256-byte aligned)
    const int64_t g4 = (a.g_begin & ~(int64_t)3) + 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    uint32_t f4 = 0;
    if (g4 < a.g_end) f4 = __ldg(reinterpret_cast<const uint32_t*>(a.flags + g4));
    unsigned front = 0, back = 0;  // bit k: column g4 + k goes to the front / the back of the list
    if (f4 & (0x01010101u * (a.in_flag | HS_FLAG_ACTIVE))) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t g = g4 + k;
            const unsigned f = (f4 >> (8 * k)) & 0xffu;
            if (g >= a.g_begin && g < a.g_end && (f & (a.in_flag | HS_FLAG_ACTIVE))) {  // ACTIVE: rescue column with c1 > 4 (write_column)
                const int c = rf_contig_of(a.col_base, a.n_contigs, g);
                if (a.desc[c].n_parts > 0) {  // :640-642: no partition, nothing kept
                    // columns of tiles with more than RF_CAP reads go to the back of the list (robust_filter_kernel<true>)
                    const int64_t tile = a.tile_base[c] + (g - a.col_base[c]) / HS_TILE;
                    if (a.tile_off[tile + 1] - a.tile_off[tile] > RF_CAP) back |= 1u << k;
                    else front |= 1u << k;
                }
            }
        }
    }
    // one reservation per CTA and list (the list order is free)
    const unsigned nf = __popc(front), nk = __popc(back);
    const unsigned inc_f = hs_warp_incl_scan((int)nf, lane), inc_b = hs_warp_incl_scan((int)nk, lane);
    if (lane == 31) {
        s_cnt[0][wid] = inc_f;
        s_cnt[1][wid] = inc_b;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned tot = 0;
        for (int w = 0; w < 8; w++) {
            const unsigned v = s_cnt[threadIdx.x][w];
            s_cnt[threadIdx.x][w] = tot;
            tot += v;
        }
        s_base[threadIdx.x] = tot ? atomicAdd(a.counters + 3 * threadIdx.x, tot) : 0u;
    }
    __syncthreads();
    unsigned of = s_base[0] + s_cnt[0][wid] + inc_f - nf, ob = s_base[1] + s_cnt[1][wid] + inc_b - nk;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (front & (1u << k)) a.active[of++] = (uint32_t)(g4 + k);
        if (back & (1u << k)) a.active[(a.g_end - a.g_begin) - 1 - (ob++)] = (uint32_t)(g4 + k);
    }
}

// alternative allele of a (column, partition) pair among tied codes: the reference iterates its robin_hood map
// and keeps the first strict maximum (:832-844). `order` = the codes (minus 33) in order of first appearance among
// the partition's reads, hist their counts (both in this warp's shared memory); see hs_select_alt.
__device__ __noinline__ int rf_select_alt_tied(const uint8_t* order, const uint32_t* hist, int m, int ref, int max2,
                                                const HsRankLut* __restrict__ lut) {
    const bool ref_excluded = ref < 128;
    bool ref_seen = false;
    for (int k = 0; k < m; k++) ref_seen |= (order[k] + HS_CODE0 == ref);
    const int n = m + (ref_seen ? 0 : 1);  // content2[ref_base] creates the entry (:833)
    if (lut && n <= 51) {
        const int level = hs_rank_level(n);
        const uint8_t* __restrict__ home = lut->home[level];
        unsigned long long occ[4] = {0, 0, 0, 0};  // 4-bit occupancy counters of up to 64 buckets
        int best = 1 << 30, nbest = 0;
        for (int k = -1; k < m; k++) {
            if (k < 0 && ref_seen) continue;
            const int key = k < 0 ? ref : order[k] + HS_CODE0;
            const int h = __ldg(home + key);
            occ[h >> 4] += 1ull << (4 * (h & 15));
            if (k < 0 || (ref_excluded && key == ref)) continue;
            if ((int)hist[key - HS_CODE0] != max2) continue;
            if (h < (best >> 8)) { best = (h << 8) | key; nbest = 1; }
            else if (h == (best >> 8)) nbest++;
        }
        int next_free = 0, maxd = 0;
        const int nb = 8 << level;
        for (int b = 0; b < nb; b++) {
            const int cnt = (int)((occ[b >> 4] >> (4 * (b & 15))) & 15ull);
            if (next_free < b) next_free = b;
            next_free += cnt;
            if (cnt && next_free - 1 - b > maxd) maxd = next_free - 1 - b;
        }
        if (nbest == 1 && maxd < 6) return best & 0xff;
    }
    HsRhTable t;
    hs_rh_new(t);
    for (int k = 0; k < m; k++) hs_rh_insert(t, (uint8_t)(order[k] + HS_CODE0));
    if (!ref_seen) hs_rh_insert(t, (uint8_t)ref);
    uint8_t it[HS_RH_MAXKEYS];
    const int nk = hs_rh_iterate(t, it);
    for (int i = 0; i < nk; i++) {
        const int key = it[i];
        if (ref_excluded && key == ref) continue;
        if (key < HS_CODE0) continue;
        if ((int)hist[key - HS_CODE0] == max2) return key;
    }
    return ' ';
}

// The lane's cell of chunk `lb` of a column's tile list (code 0 = the read does not cover the column)
struct RfCell {
    int code;
    int32_t n;  // read index inside the contig
};
__device__ __forceinline__ RfCell rf_gather_cell(const FilterArgs& a, int q, int64_t lb, int64_t l1, int64_t read0, int lane) {
    RfCell c;
    c.code = 0;
    c.n = 0;
    const int64_t l = lb + lane;
    if (l < l1) {
        const int32_t r = __ldg(a.tile_reads + l);
        if (__ldg(a.read_start + r) <= q && q < __ldg(a.read_end + r)) {
            c.code = __ldg(a.codes + __ldg(a.row_base + r) + q);
            c.n = (int32_t)(r - read0);
        }
    }
    return c;
}

__device__ __noinline__ float rf_chi_square(int n00, int n01, int n10, int n11) { return hs_chi_square(n00, n01, n10, n11); }

// Several codes share the maximum of a (column, partition) pair: the reference's map iteration order decides
// (:832-844). Lane 0 rebuilds the order of first appearance among the partition's reads -- from the staged cells and
// state rows (s_code, s_row: word kw of every row, shift sh), or from the pileup for a column too deep to stage.
__device__ __noinline__ int rf_tied_alt(const FilterArgs& a, const uint32_t* rows_c, int pwords, int p, int q, int64_t l0, int64_t l1,
                                         int64_t read0, const uint8_t* s_code, const uint32_t* s_row, int ncell_staged,
                                         uint32_t* s_hist, int ref, int nref, int max2) {
    uint8_t order[HS_NCODES + 1];
    int mo = 0;
    auto see = [&](int code) {
        const int idx = code - HS_CODE0;
        bool seen = false;
        for (int k = 0; k < mo; k++) seen |= order[k] == idx;
        if (!seen) order[mo++] = (uint8_t)idx;
    };
    const int kw = (p & 127) >> 4, sh = 2 * (p & 15);
    if (ncell_staged >= 0) {
        for (int i = 0; i < ncell_staged; i++)
            if ((s_row[i * 9 + kw] >> sh) & 3u) see(s_code[i]);
    } else {
        for (int64_t l = l0; l < l1; l++) {
            const int32_t r = a.tile_reads[l];
            if (!(a.read_start[r] <= q && q < a.read_end[r])) continue;
            const uint32_t w = rows_c[(int64_t)(r - read0) * pwords + (p >> 4)];
            if (((w >> sh) & 3u) == 0) continue;
            see(a.codes[a.row_base[r] + q]);
        }
    }
    if (nref > 0) s_hist[ref - HS_CODE0] = (uint32_t)nref;
    const int alt = rf_select_alt_tied(order, s_hist, mo, ref, max2, a.lut);
    if (nref > 0) s_hist[ref - HS_CODE0] = 0;
    return alt;
}

// DEEP = false: columns whose tile has at most RF_CAP reads (every ordinary depth): the cells' codes sit in four
// registers per lane, their state rows in shared memory, and the chunk loops are unrolled. DEEP = true: amplicon-deep
// columns, gathered again from the pileup for every partition. The two kinds come as the front and the back of the
// active list (filter_active_kernel).
template <bool DEEP, int MINB>
__global__ void __launch_bounds__(32 * RF_WARPS, MINB) robust_filter_kernel(FilterArgs a) {
    __shared__ uint8_t s_code_all[RF_WARPS][DEEP ? 4 : RF_CAP];
    __shared__ __align__(16) uint32_t s_row_all[RF_WARPS][DEEP ? 4 : RF_CAP * 9];  // 8 state words per cell, rows padded to 9 words
    __shared__ uint32_t s_hist_all[RF_WARPS][HS_NCODES + 3];
    __shared__ uint8_t s_touched_all[RF_WARPS][HS_NCODES + 3];
    __shared__ int s_m_all[RF_WARPS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* const s_code = s_code_all[wid];
    uint32_t* const s_row = s_row_all[wid];
    uint32_t* const s_hist = s_hist_all[wid];
    uint8_t* const s_touched = s_touched_all[wid];
    int* const s_m = &s_m_all[wid];
    const unsigned lt = (1u << lane) - 1u;
    for (int i = lane; i < HS_NCODES + 3; i += 32) s_hist[i] = 0;
    if (lane == 0) *s_m = 0;
    __syncwarp();
    const unsigned n_items = DEEP ? a.counters[3] : (a.from_overflow ? a.counters[6] : a.counters[0]);
    const int64_t n_cols_all = a.g_end - a.g_begin;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(a.counters + (DEEP ? 5 : (a.from_overflow ? 7 : 1)), 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int64_t g = DEEP ? a.active[n_cols_all - 1 - item] : (a.from_overflow ? a.overflow[item] : a.active[item]);
        const int c = rf_contig_of(a.col_base, a.n_contigs, g);
        const FilterDesc d = a.desc[c];
        const int q = (int)(g - a.col_base[c]);
        const int64_t read0 = a.contig_read_off[c];
        const int64_t tile = a.tile_base[c] + q / HS_TILE;
        const int64_t l0 = a.tile_off[tile], l1 = a.tile_off[tile + 1];
        const int ref = a.k0[g];
        const unsigned f = a.flags[g];
        const bool inlist = (f & a.in_flag) != 0;
        const uint32_t* __restrict__ rows_c = a.rows + d.row_off;
        // A column outside snps_in is only kept through loop 4, which needs n10 + n00 > 4 (:756): at least 5 of a
        // partition's reads on one code other than ref_base. When the column's third most frequent code has at most 4
        // carriers in all (c2 <= 4), only its second code k1 can get there, and it then is the strict maximum of the
        // partition: no histogram, one pass of ballots per partition. (Codes >= 128 never count as ref_base in the
        // reference's comparison, :838: those columns take the general path.)
        const bool single = !inlist && ref < 128 && a.counts[3 * g + 2] <= 4u;
        const int k1 = a.k1[g];
        // ---- the column's cells, in ascending read order: compacted, slot 32 * ch + lane (not DEEP) ----
        constexpr int NCH = DEEP ? 1 : RF_CAP / 32;
        int code4[NCH];
        int32_t n4[NCH];
        int ncell = 0;
        if (!DEEP) {
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) { code4[ch] = 0; n4[ch] = 0; }
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                if (l0 + 32 * ch < l1) {
                    const RfCell cl = rf_gather_cell(a, q, l0 + 32 * ch, l1, read0, lane);
                    const unsigned mk = __ballot_sync(0xffffffffu, cl.code != 0);
                    const int o = ncell + __popc(mk & lt);
                    // compaction through shuffles would need a variable source lane per slot: two staging bytes instead
                    if (cl.code != 0) {
                        s_code[o] = (uint8_t)cl.code;
                        s_row[o * 9 + 8] = (uint32_t)cl.n;  // the pad word of the row carries the read index until staged
                    }
                    ncell += __popc(mk);
                }
            }
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                const int slot = 32 * ch + lane;
                if (slot < ncell) {
                    code4[ch] = s_code[slot];
                    n4[ch] = (int32_t)s_row[slot * 9 + 8];
                }
            }
        }
        const int nchunk = DEEP ? (int)((l1 - l0 + 31) >> 5) : (ncell + 31) >> 5;
        // ballots of the two codes the single-candidate path looks at (one per chunk, constant over the partitions)
        unsigned refm[NCH], altm[NCH];
        if (!DEEP) {
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                refm[ch] = __ballot_sync(0xffffffffu, code4[ch] == ref);
                altm[ch] = __ballot_sync(0xffffffffu, code4[ch] != 0 && code4[ch] == k1);
            }
        }
        bool keep = false;
        int n_visited = 0;  // partitions whose states were read
        for (int pb = 0; pb < d.n_parts && !keep; pb += 128) {
            // the cells' state rows of this block of 128 partitions; which partitions hold one of the column's reads
            uint32_t nz[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (!DEEP) {
#pragma unroll
                for (int ch = 0; ch < NCH; ch++) {
                    if (code4[ch] != 0) {
                        const int slot = 32 * ch + lane;
                        const uint4* src = reinterpret_cast<const uint4*>(rows_c + (int64_t)n4[ch] * d.pwords + (pb >> 4));
                        const uint4 x = __ldg(src), y = __ldg(src + 1);
                        const uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            nz[k] |= w[k];
                            s_row[slot * 9 + k] = w[k];
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int ch = 0; ch < nchunk; ch++) {
                    const RfCell cl = rf_gather_cell(a, q, l0 + 32 * (int64_t)ch, l1, read0, lane);
                    if (ch == 0) ncell = 0;
                    ncell += __popc(__ballot_sync(0xffffffffu, cl.code != 0));
                    if (cl.code != 0) {
                        const uint4* src = reinterpret_cast<const uint4*>(rows_c + (int64_t)cl.n * d.pwords + (pb >> 4));
                        const uint4 x = __ldg(src), y = __ldg(src + 1);
                        nz[0] |= x.x; nz[1] |= x.y; nz[2] |= x.z; nz[3] |= x.w;
                        nz[4] |= y.x; nz[5] |= y.y; nz[6] |= y.z; nz[7] |= y.w;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                nz[k] = __reduce_or_sync(0xffffffffu, nz[k]);
                nz[k] = (nz[k] | (nz[k] >> 1)) & 0x55555555u;  // bit 2j = partition 16k + j holds a read of the column
            }
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < 8 && !keep; k++) {
                uint32_t bits = nz[k];
                if (!bits) continue;
                // word k of the cells' state rows: the states of partitions 16k .. 16k+15 of this block
                uint32_t wr[NCH];
                if (!DEEP) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) wr[ch] = code4[ch] != 0 ? s_row[(32 * ch + lane) * 9 + k] : 0u;
                }
#pragma unroll 1
                while (bits && !keep) {
                    const int pl = 16 * k + ((__ffs(bits) - 1) >> 1);
                    const int sh = 2 * (pl & 15);
                    bits &= bits - 1;
                    n_visited++;
                    // the lane's cell of chunk ch and its state in this partition
#define RF_EACH_CHUNK(...)                                                                           \
    if (!DEEP) {                                                                                      \
        _Pragma("unroll") for (int ch = 0; ch < NCH; ch++) {                                          \
            if (32 * ch < ncell) {                                                                    \
                const int code = code4[ch];                                                           \
                const int sg = (int)((wr[ch] >> sh) & 3u);                                            \
                __VA_ARGS__                                                                           \
            }                                                                                         \
        }                                                                                             \
    } else {                                                                                          \
        _Pragma("unroll 1") for (int ch = 0; ch < nchunk; ch++) {                                     \
            const RfCell cl_ = rf_gather_cell(a, q, l0 + 32 * (int64_t)ch, l1, read0, lane);           \
            const int code = cl_.code;                                                                \
            const int sg = code == 0 ? 0 : (int)((__ldg(rows_c + (int64_t)cl_.n * d.pwords + ((pb + pl) >> 4)) >> sh) & 3u); \
            __VA_ARGS__                                                                               \
        }                                                                                             \
    }
                    int n00 = 0, n01 = 0, n10 = 0, n11 = 0;
                    bool table = false;
                    if (single) {
                        int cnt1 = 0;
                        if (!DEEP) {
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++) {
                                if (32 * ch < ncell) {
                                    const int sg = (int)((wr[ch] >> sh) & 3u);
                                    const unsigned b1 = __ballot_sync(0xffffffffu, sg == 1), b2 = __ballot_sync(0xffffffffu, sg == 2);
                                    const unsigned b3 = __ballot_sync(0xffffffffu, sg == 3);
                                    n11 += __popc(b1 & refm[ch]);
                                    n01 += __popc(b2 & refm[ch]);
                                    n10 += __popc(b1 & altm[ch]);
                                    n00 += __popc(b2 & altm[ch]);
                                    cnt1 += __popc((b1 | b2 | b3) & altm[ch]);
                                }
                            }
                        } else {
                            RF_EACH_CHUNK({
                                const unsigned b1 = __ballot_sync(0xffffffffu, sg == 1), b2 = __ballot_sync(0xffffffffu, sg == 2);
                                const unsigned b3 = __ballot_sync(0xffffffffu, sg == 3);
                                const unsigned bref = __ballot_sync(0xffffffffu, code == ref);
                                const unsigned balt = __ballot_sync(0xffffffffu, code != 0 && code == k1);
                                n11 += __popc(b1 & bref);
                                n01 += __popc(b2 & bref);
                                n10 += __popc(b1 & balt);
                                n00 += __popc(b2 & balt);
                                cnt1 += __popc((b1 | b2 | b3) & balt);
                            })
                        }
                        table = cnt1 > 4;
                    } else {
                        // ---- pass 1: the partition's reads on this column. The column's own majority code is counted
                        // with ballots (its rows with state +1 / -1 are n11 / n01, :893-949), every other code goes
                        // through the warp's histogram ----
                        int nb = 0, nref = 0;
                        RF_EACH_CHUNK({
                            const bool act = sg != 0;
                            const bool isref = act && code == ref;
                            nb += __popc(__ballot_sync(0xffffffffu, act));
                            nref += __popc(__ballot_sync(0xffffffffu, isref));
                            n11 += __popc(__ballot_sync(0xffffffffu, isref && sg == 1));
                            n01 += __popc(__ballot_sync(0xffffffffu, isref && sg == 2));
                            if (act && !isref) {
                                if (atomicAdd(&s_hist[code - HS_CODE0], 1u) == 0u) s_touched[atomicAdd(s_m, 1)] = (uint8_t)(code - HS_CODE0);
                            }
                        })
                        __syncwarp();
                        const int m = *s_m;  // distinct codes other than ref_base
                        int maxc = -1;
                        for (int k2 = lane; k2 < m; k2 += 32) maxc = max(maxc, (int)s_hist[s_touched[k2]]);
                        maxc = __reduce_max_sync(0xffffffffu, maxc);
                        if (nb > 0 && (inlist || ref >= 128 || maxc > 4)) {
                            // secondFrequent (:832-844): the most frequent code other than ref_base (ref_base itself
                            // competes when it is >= 128)
                            int max2 = maxc, alt = ' ';
                            const bool ref_competes = ref >= 128 && nref > 0;
                            if (ref_competes && nref > max2) max2 = nref;
                            int ties = 0, cand = ' ';
                            for (int k0 = 0; k0 < m; k0 += 32) {
                                const int k2 = k0 + lane;
                                const bool hitk = k2 < m && (int)s_hist[s_touched[k2]] == max2;
                                const unsigned mk = __ballot_sync(0xffffffffu, hitk);
                                ties += __popc(mk);
                                if (mk) cand = __shfl_sync(0xffffffffu, hitk ? (int)s_touched[k2] + HS_CODE0 : 0, __ffs(mk) - 1);
                            }
                            if (ref_competes && nref == max2) { ties++; cand = ref; }
                            if (max2 >= 0) {
                                if (ties == 1) {
                                    alt = cand;
                                } else {
                                    int alt0 = ' ';
                                    if (lane == 0)
                                        alt0 = rf_tied_alt(a, rows_c, d.pwords, pb + pl, q, l0, l1, read0, s_code, s_row, DEEP ? -1 : ncell,
                                                           s_hist, ref, nref, max2);
                                    alt = __shfl_sync(0xffffffffu, alt0, 0);
                                }
                            }
                            // ---- pass 2: n10 / n00 ----
                            if (alt != ref && alt != ' ') {
                                RF_EACH_CHUNK({
                                    const int s2 = code == alt ? sg : 0;
                                    n10 += __popc(__ballot_sync(0xffffffffu, s2 == 1));
                                    n00 += __popc(__ballot_sync(0xffffffffu, s2 == 2));
                                })
                            }
                            table = true;
                        }
                        // reset the histogram
                        for (int k2 = lane; k2 < m; k2 += 32) s_hist[s_touched[k2]] = 0;
                        if (lane == 0) *s_m = 0;
                        __syncwarp();
                    }
#undef RF_EACH_CHUNK
                    if (table) {
                        // loop 3 (:721-738): columns of snps_in; loop 4 (:745-764): rescue of every other column (also
                        // of suspects that failed loop 3). The chi-square is only evaluated where an integer
                        // condition leaves the decision open.
                        const bool c3 = inlist && (double)(n00 + n01 + n10 + n11) > __dmul_rn(0.5, (double)a.depth[g]);
                        const bool c4 = (f & HS_FLAG_RESCUE) && n10 + n00 > 4 && n01 + n11 > 4;
                        if (c3 || c4) {
                            const float chi = rf_chi_square(n00, n01, n10, n11);
                            if ((c3 && chi > 15.f) || (c4 && (double)chi > 20.0)) keep = true;
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (keep) atomicOr(a.kept + (g >> 5), 1u << (g & 31));
            atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 8), (unsigned long long)ncell);
            atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 10), (unsigned long long)ncell * (unsigned)n_visited);
        }
    }
}

// ---- loops 3+4, one LANE per partition (robust_filter_lanes_kernel) -------------------------------------------------
// The kernel above takes the partitions of a column one after the other with the column's reads on the lanes; most
// of its time went into the pairs where several codes tie for the alternative allele (a partition whose reads carry
// ref_base plus a few different singletons is the common case of a suspect column), which one lane settled alone.
// Here a warp still owns a column, but the lanes are the PARTITIONS that hold one of the column's reads: the column's
// cells are grouped by code once (distinct codes in order of first appearance, at most RL_MAXC), every lane runs over
// the cells of a code and counts its partition's states in one packed register (a byte per state), and the
// per-partition decisions -- most frequent other code, ties by the bucket-table rule, 2x2 table, chi-square -- are
// taken by all lanes at once. Columns with more distinct codes than RL_MAXC (24) go to an overflow list that the kernel
// above works off afterwards.
#define RL_WARPS 4
#ifndef RL_MAXC
#define RL_MAXC 24  // measured: 16 sends too many columns to the overflow kernel (config 3: +16 ms), 32 costs a CTA per SM
#endif
#ifndef RL_MINB
#define RL_MINB 6
#endif

// the alternative allele among tied codes, for this lane's partition: rf_select_alt_tied's bucket-table rule on the
// column's code list and the lane's counts (s_acc column); returns the index of the code in the list, -1 for none,
// -2 when the rule cannot decide (the literal replay needs the order of first appearance: rl_tied_slow)
__device__ __noinline__ int rl_tied_fast(const uint8_t* s_clist, const uint32_t* s_acc_lane, int M, int jref, int ref, int nref,
                                         int max2, int n_keys, const HsRankLut* __restrict__ lut) {
    if (!lut || n_keys > 51) return -2;
    const bool ref_excluded = ref < 128;
    const int level = hs_rank_level(n_keys);
    const uint8_t* __restrict__ home = lut->home[level];
    unsigned long long occ0 = 0, occ1 = 0, occ2 = 0, occ3 = 0;  // 4-bit occupancy counters of up to 64 buckets
    int best = 1 << 30, nbest = 0;
    auto put = [&](int h) {
        const unsigned long long one = 1ull << (4 * (h & 15));
        const int w = h >> 4;
        if (w == 0) occ0 += one; else if (w == 1) occ1 += one; else if (w == 2) occ2 += one; else occ3 += one;
    };
    if (jref < 0) put(__ldg(home + ref));  // content2[ref_base] creates the entry (:833)
    for (int j = 0; j < M; j++) {
        const uint32_t acc = s_acc_lane[j * 32];
        const int any = (int)((acc >> 8) & 255u) + (int)((acc >> 16) & 255u) + (int)(acc >> 24);
        const bool is_ref = j == jref;
        if (!is_ref && any == 0) continue;  // not a key of this partition's map
        const int h = __ldg(home + s_clist[j]);
        put(h);
        if (is_ref && (nref == 0 || ref_excluded)) continue;
        if ((is_ref ? nref : any) != max2) continue;
        if (h < (best >> 8)) { best = (h << 8) | j; nbest = 1; }
        else if (h == (best >> 8)) nbest++;
    }
    int next_free = 0, maxd = 0;
    const int nb = 8 << level;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        if (16 * w < nb) {
            const unsigned long long o = w == 0 ? occ0 : (w == 1 ? occ1 : (w == 2 ? occ2 : occ3));
#pragma unroll
            for (int bb = 0; bb < 16; bb++) {
                const int b = 16 * w + bb;
                if (b < nb) {
                    const int cnt = (int)((o >> (4 * bb)) & 15ull);
                    if (next_free < b) next_free = b;
                    next_free += cnt;
                    if (cnt && next_free - 1 - b > maxd) maxd = next_free - 1 - b;
                }
            }
        }
    }
    if (nbest == 1 && maxd < 6) return best & 0xff;
    return -2;
}

// the literal replay for this lane's partition: order of first appearance among the partition's reads from the staged
// cells (slot order = read order), counts from the lane's s_acc column; returns the code or ' '
__device__ __noinline__ int rl_tied_slow(const uint8_t* s_code, const uint32_t* s_row, int ncell, int kw, int sh,
                                         const uint8_t* s_clist, const uint32_t* s_acc_lane, int M, int ref, int nref, int max2) {
    uint8_t order[HS_NCODES + 1];
    uint32_t hist[HS_NCODES + 3];
    for (int i = 0; i < HS_NCODES + 3; i++) hist[i] = 0;
    int mo = 0;
    for (int i = 0; i < ncell; i++) {
        if (((s_row[i * 9 + kw] >> sh) & 3u) == 0) continue;
        const int idx = s_code[i] - HS_CODE0;
        bool seen = false;
        for (int k = 0; k < mo; k++) seen |= order[k] == idx;
        if (!seen) order[mo++] = (uint8_t)idx;
    }
    for (int j = 0; j < M; j++) {
        const uint32_t acc = s_acc_lane[j * 32];
        hist[s_clist[j] - HS_CODE0] = ((acc >> 8) & 255u) + ((acc >> 16) & 255u) + (acc >> 24);
    }
    (void)nref;  // the count of ref_base is part of the histogram already
    return rf_select_alt_tied(order, hist, mo, ref, max2, nullptr);
}

__global__ void __launch_bounds__(32 * RL_WARPS, RL_MINB) robust_filter_lanes_kernel(FilterArgs a) {
    __shared__ uint8_t s_code_all[RL_WARPS][RF_CAP];
    __shared__ uint16_t s_sorted_all[RL_WARPS][RF_CAP];  // word offsets of the cells' state rows, grouped by code
    __shared__ uint32_t s_map_all[RL_WARPS][32];         // byte per code: its index in the column's list
    __shared__ uint32_t s_cnt_all[RL_WARPS][RL_MAXC + 4];
    __shared__ uint8_t s_clist_all[RL_WARPS][RL_MAXC];
    __shared__ uint8_t s_cstart_all[RL_WARPS][RL_MAXC + 4];
    __shared__ uint8_t s_plist_all[RL_WARPS][128];
    __shared__ __align__(16) uint32_t s_row_all[RL_WARPS][RF_CAP * 9];  // 8 state words per cell, rows padded to 9 words
    __shared__ uint32_t s_acc_all[RL_WARPS][RL_MAXC * 32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* const s_code = s_code_all[wid];
    uint16_t* const s_sorted = s_sorted_all[wid];
    uint32_t* const s_map = s_map_all[wid];
    uint32_t* const s_cnt32 = s_cnt_all[wid];
    uint8_t* const s_clist = s_clist_all[wid];
    uint8_t* const s_cstart = s_cstart_all[wid];
    uint8_t* const s_plist = s_plist_all[wid];
    uint32_t* const s_row = s_row_all[wid];
    uint32_t* const s_acc = s_acc_all[wid];
    const unsigned lt = (1u << lane) - 1u;
    const unsigned n_items = a.counters[0];
    constexpr int NCH = RF_CAP / 32;
    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(a.counters + 1, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int64_t g = a.active[item];
        const int c = rf_contig_of(a.col_base, a.n_contigs, g);
        const FilterDesc d = a.desc[c];
        const int q = (int)(g - a.col_base[c]);
        const int64_t read0 = a.contig_read_off[c];
        const int64_t tile = a.tile_base[c] + q / HS_TILE;
        const int64_t l0 = a.tile_off[tile], l1 = a.tile_off[tile + 1];
        const int ref = a.k0[g];
        const unsigned f = a.flags[g];
        const bool inlist = (f & a.in_flag) != 0;
        const uint32_t* __restrict__ rows_c = a.rows + d.row_off;
        __syncwarp();
        // ---- the column's cells, in ascending read order: compacted, slot 32 * ch + lane ----
        int code4[NCH];
        int32_t n4[NCH];
        int ncell = 0;
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) { code4[ch] = 0; n4[ch] = 0; }
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            if (l0 + 32 * ch < l1) {
                const RfCell cl = rf_gather_cell(a, q, l0 + 32 * ch, l1, read0, lane);
                const unsigned mk = __ballot_sync(0xffffffffu, cl.code != 0);
                const int o = ncell + __popc(mk & lt);
                if (cl.code != 0) {
                    s_code[o] = (uint8_t)cl.code;
                    s_row[o * 9 + 8] = (uint32_t)cl.n;  // the pad word of the row carries the read index until staged
                }
                ncell += __popc(mk);
            }
        }
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            const int slot = 32 * ch + lane;
            if (slot < ncell) {
                code4[ch] = s_code[slot];
                n4[ch] = (int32_t)s_row[slot * 9 + 8];
            }
        }
        // ---- the column's distinct codes (any order: the tie rules do not depend on it) and the cells grouped by
        // code: s_map[code] = index in the list, claimed chunk by chunk by the first lane of every group of equal
        // codes; then a counting sort of the cell slots by that index ----
        s_map[lane] = 0xffffffffu;  // 128 bytes: codes - HS_CODE0 < 128
        for (int i = lane; i < RL_MAXC + 4; i += 32) s_cnt32[i] = 0u;
        __syncwarp();
        int M = 0, jref = -1;
        int cidx4[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            cidx4[ch] = 0;
            if (32 * ch < ncell) {
                const bool valid = code4[ch] != 0;
                const unsigned peers = __match_any_sync(0xffffffffu, code4[ch]);
                uint8_t* const slot_of = reinterpret_cast<uint8_t*>(s_map) + (valid ? code4[ch] - HS_CODE0 : 0);
                const bool isnew = valid && (__ffs(peers) - 1 == lane) && *slot_of == 0xff;
                const unsigned nm = __ballot_sync(0xffffffffu, isnew);
                if (isnew) {
                    const int j = M + __popc(nm & lt);
                    *slot_of = (uint8_t)min(j, 0xfe);
                    if (j < RL_MAXC) s_clist[j] = (uint8_t)code4[ch];
                }
                M += __popc(nm);
                __syncwarp();
            }
        }
        if (M > RL_MAXC) {  // more distinct codes than the lanes keep counts for: the general kernel takes the column
            if (lane == 0) a.overflow[atomicAdd(a.counters + 6, 1u)] = (uint32_t)g;
            continue;
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
            if (code4[ch] != 0) {
                cidx4[ch] = reinterpret_cast<const uint8_t*>(s_map)[code4[ch] - HS_CODE0];
                atomicAdd(&s_cnt32[cidx4[ch]], 1u);
            }
        }
        __syncwarp();
        {
            const int cnt = lane < M ? (int)s_cnt32[lane] : 0;
            const int incl = hs_warp_incl_scan(cnt, lane);
            if (lane < M) {
                s_cstart[lane] = (uint8_t)(incl - cnt);
                s_cnt32[lane] = (uint32_t)(incl - cnt);  // becomes the fill cursor of the code
                if (s_clist[lane] == ref) jref = lane;
            }
            if (lane == M - 1 || (M == 0 && lane == 0)) s_cstart[M] = (uint8_t)(M ? incl : 0);
            jref = __reduce_max_sync(0xffffffffu, jref);
        }
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < NCH; ch++)
            if (code4[ch] != 0) s_sorted[atomicAdd(&s_cnt32[cidx4[ch]], 1u)] = (uint16_t)((32 * ch + lane) * 9);
        __syncwarp();
        bool keep = false;
        int n_visited = 0;
        for (int pb = 0; pb < d.n_parts && !keep; pb += 128) {
            // the cells' state rows of this block of 128 partitions; which partitions hold one of the column's reads
            uint32_t nz[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                if (code4[ch] != 0) {
                    const int slot = 32 * ch + lane;
                    const uint4* src = reinterpret_cast<const uint4*>(rows_c + (int64_t)n4[ch] * d.pwords + (pb >> 4));
                    const uint4 x = __ldg(src), y = __ldg(src + 1);
                    const uint32_t w[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        nz[k] |= w[k];
                        s_row[slot * 9 + k] = w[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                nz[k] = __reduce_or_sync(0xffffffffu, nz[k]);
                nz[k] = (nz[k] | (nz[k] >> 1)) & 0x55555555u;  // bit 2j = partition 16k + j holds a read of the column
            }
            // the list of those partitions, ascending
            int n_present = 0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t wz = lane < 16 ? nz[2 * r] : nz[2 * r + 1];
                const bool pres = (wz >> (2 * (lane & 15))) & 1u;
                const unsigned mk = __ballot_sync(0xffffffffu, pres);
                if (pres) s_plist[n_present + __popc(mk & lt)] = (uint8_t)(32 * r + lane);
                n_present += __popc(mk);
            }
            __syncwarp();
            n_visited += n_present;
            for (int g0 = 0; g0 < n_present && !keep; g0 += 32) {
                const bool has = g0 + lane < n_present;
                const int pl = has ? s_plist[g0 + lane] : 0;
                const int kw = pl >> 4, sh = 2 * (pl & 15);
                // counts of this lane's partition per code: byte 1 = state +1, byte 2 = state -1, byte 3 = state 0
                int nb = 0, nref = 0, n11 = 0, n01 = 0, maxc = -1, ndist = 0, ties = 0, jbest = -1;
#pragma unroll 1
                for (int j = 0; j < M; j++) {
                    const int x1 = s_cstart[j + 1];
                    uint32_t acc = 0;
#pragma unroll 4
                    for (int x = s_cstart[j]; x < x1; x++) {
                        const uint32_t t = (s_row[(int)s_sorted[x] + kw] >> sh) & 3u;
                        acc += 1u << (8 * t);
                    }
                    s_acc[j * 32 + lane] = acc;
                    const int b1 = (int)((acc >> 8) & 255u), b2 = (int)((acc >> 16) & 255u);
                    const int any = b1 + b2 + (int)(acc >> 24);
                    nb += any;
                    if (j == jref) {
                        nref = any;
                        n11 = b1;
                        n01 = b2;
                    } else if (any > 0) {
                        ndist++;
                        if (any > maxc) { maxc = any; ties = 1; jbest = j; }
                        else if (any == maxc) ties++;
                    }
                }
                bool keep_l = false;
                if (has && nb > 0 && (inlist || ref >= 128 || maxc > 4)) {
                    // secondFrequent (:832-844): the most frequent code other than ref_base (ref_base itself competes
                    // when it is >= 128)
                    int max2 = maxc, jalt = jbest;
                    if (ref >= 128 && nref > 0) {
                        if (nref > max2) { max2 = nref; ties = 1; jalt = jref; }
                        else if (nref == max2) ties++;
                    }
                    if (max2 < 0) jalt = -1;
                    if (max2 >= 0 && ties > 1) {
                        jalt = rl_tied_fast(s_clist, s_acc + lane, M, jref, ref, nref, max2, ndist + 1, a.lut);
                        if (jalt == -2) {
                            const int alt = rl_tied_slow(s_code, s_row, ncell, kw, sh, s_clist, s_acc + lane, M, ref, nref, max2);
                            jalt = -1;
                            for (int j = 0; j < M; j++) if (s_clist[j] == alt) jalt = j;
                        }
                    }
                    int n10 = 0, n00 = 0;
                    if (jalt >= 0 && jalt != jref) {
                        const uint32_t acc = s_acc[jalt * 32 + lane];
                        n10 = (int)((acc >> 8) & 255u);
                        n00 = (int)((acc >> 16) & 255u);
                    }
                    // loop 3 (:721-738): columns of snps_in; loop 4 (:745-764): rescue of every other column (also of
                    // suspects that failed loop 3)
                    const bool c3 = inlist && (double)(n00 + n01 + n10 + n11) > __dmul_rn(0.5, (double)a.depth[g]);
                    const bool c4 = (f & HS_FLAG_RESCUE) && n10 + n00 > 4 && n01 + n11 > 4;
                    if (c3 || c4) {
                        const float chi = rf_chi_square(n00, n01, n10, n11);
                        keep_l = (c3 && chi > 15.f) || (c4 && (double)chi > 20.0);
                    }
                }
                keep = __any_sync(0xffffffffu, keep_l);
            }
        }
        if (lane == 0) {
            if (keep) atomicOr(a.kept + (g >> 5), 1u << (g & 31));
            atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 8), (unsigned long long)ncell);
            atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 10), (unsigned long long)ncell * (unsigned)n_visited);
        }
    }
}

// ascending compaction of the kept columns, one CTA of 1024 threads per contig over the kept bitmap: a first sweep
// counts, one atomic reserves the contig's slice of the packed list (hdr[2c] = start, hdr[2c+1] = count), a second
// sweep writes the positions in order. A round covers 32 columns per thread.
__device__ __forceinline__ unsigned rf_kept_bits(const uint32_t* __restrict__ kept, int64_t g0, int L, int q0) {
    if (q0 >= L) return 0u;
    const int64_t g = g0 + q0;
    const int sh = (int)(g & 31);
    const uint32_t lo = __ldg(kept + (g >> 5));
    const uint32_t hi = sh ? __ldg(kept + (g >> 5) + 1) : 0u;  // the bitmap has a spare word at its end
    unsigned bits = __funnelshift_r(lo, hi, sh);
    if (L - q0 < 32) bits &= (1u << (L - q0)) - 1u;
    return bits;
}
__global__ void __launch_bounds__(1024) kept_scan_kernel(int contig0, const int64_t* __restrict__ col_base,
                                                         const int32_t* __restrict__ contig_len,
                                                         const uint32_t* __restrict__ kept, unsigned int* __restrict__ total,
                                                         int64_t capacity, int32_t* __restrict__ out, int64_t* __restrict__ hdr) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    __shared__ unsigned s_start;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c = contig0 + blockIdx.x;
    const int L = contig_len[c];
    const int64_t g0 = col_base[c];
    int cnt_all = 0;
    for (int r0 = 0; r0 < L; r0 += 1024 * 32) cnt_all += __popc(rf_kept_bits(kept, g0, L, r0 + 32 * tid));
    cnt_all = hs_warp_incl_scan(cnt_all, lane);
    if (lane == 31) s_warp[wid] = cnt_all;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 32; w++) t += s_warp[w];
        s_start = atomicAdd(total, (unsigned)t);
        hdr[2 * blockIdx.x] = s_start;
        hdr[2 * blockIdx.x + 1] = t;
    }
    __syncthreads();
    int64_t n = s_start;  // list position of the next kept column (same value in every thread)
    for (int r0 = 0; r0 < L; r0 += 1024 * 32) {
        const int q0 = r0 + 32 * tid;
        unsigned bits = rf_kept_bits(kept, g0, L, q0);
        const int cnt = __popc(bits);
        const int incl = hs_warp_incl_scan(cnt, lane);
        __syncthreads();
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int wi = hs_warp_incl_scan(s_warp[lane], lane);
            s_warp[lane] = wi - s_warp[lane];
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        int64_t i = n + s_warp[wid] + incl - cnt;
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            if (i < capacity) out[i] = q0 + j;
            i++;
        }
        n += s_total;
    }
}

__global__ void set_inlist_kernel(int n, const int32_t* __restrict__ pos, int64_t g0, uint8_t* __restrict__ flags,
                                  int set) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (set) flags[g0 + pos[i]] |= HS_FLAG_INLIST;
    else flags[g0 + pos[i]] &= (uint8_t)~HS_FLAG_INLIST;
}

// dense partition-state matrix on the host: 0 absent/masked, 1 = +1, 2 = -1, 3 = 0, |4 solid
// out_t (optional): the transpose [n_reads][npad], rows padded to npad partitions
static int build_pstate(hsgpu_ctx* ctx, const hsgpu_partitions* parts, int64_t n_reads, std::vector<uint8_t>& out,
                        std::vector<uint8_t>* out_t = nullptr, int npad = 0) {
    out.assign((size_t)parts->n_parts * (size_t)n_reads, 0);
    if (out_t) out_t->assign((size_t)std::max<int64_t>(n_reads, 1) * (size_t)npad, 0);
    for (int p = 0; p < parts->n_parts; p++) {
        for (int64_t i = parts->part_off[p]; i < parts->part_off[p + 1]; i++) {
            const int32_t n = parts->read_idx[i];
            if (n < 0 || n >= n_reads) HS_FAIL(ctx, HSGPU_ERR_ARG, "partition read index out of range");
            uint8_t v = 0;
            switch (parts->state[i]) {
                case 1: v = 1; break;
                case -1: v = 2; break;
                case 0: v = 3; break;
                default: v = 0; break;  // -2: masked
            }
            if (v && parts->less && parts->more && parts->less[i] <= 1 && parts->more[i] >= 3) v |= 4;
            out[(size_t)p * n_reads + n] = v;
            if (out_t) (*out_t)[(size_t)n * npad + p] = v;
        }
    }
    return HSGPU_OK;
}

static const int kTablesSmem = HS_NCODES * 128 * 3 + CT_ROWS * 5;

extern "C" {

int hsgpu_partition_tables(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_cols,
                           const int32_t* pos, hsgpu_distance* out) {
    if (!p || !parts || contig < 0 || contig >= p->n_contigs || n_cols < 0) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_partition_tables: call hsgpu_column_rank first");
    if (n_cols == 0 || parts->n_parts == 0) return HSGPU_OK;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t R = p->h_contig_read_off[contig + 1] - p->h_contig_read_off[contig];
    const int64_t L = p->h_contig_len[contig];
    for (int i = 0; i < n_cols; i++)
        if (pos[i] < 0 || pos[i] >= L) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_partition_tables: position out of range");
    std::vector<uint8_t> pst;
    int rc = build_pstate(ctx, parts, R, pst);
    if (rc) return rc;
    uint8_t* d_pst = nullptr;
    int32_t* d_pos = nullptr;
    hsgpu_distance* d_out = nullptr;
    const int64_t n_out = (int64_t)n_cols * parts->n_parts;
    HsTemps temps(ctx);
    temps.own(d_pst, d_pos, d_out);
    HS_CUDA(ctx, hs_alloc(ctx, &d_pst, (int64_t)pst.size()));
    HS_CUDA(ctx, hs_alloc(ctx, &d_pos, n_cols));
    HS_CUDA(ctx, hs_alloc(ctx, &d_out, n_out));
    HS_CUDA(ctx, hs_h2d(ctx, d_pst, pst.data(), (int64_t)pst.size()));
    HS_CUDA(ctx, hs_h2d(ctx, d_pos, pos, n_cols));
    static std::atomic<bool> attr[64];  // per device: function attributes belong to the device's context
    if (!attr[ctx->device & 63]) {
        HS_CUDA(ctx, cudaFuncSetAttribute(partition_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTablesSmem));
        attr[ctx->device & 63] = true;
    }
    TablesArgs a;
    a.n_cols = n_cols;
    a.pos = d_pos;
    a.tile0 = p->h_tile_base[contig];
    a.read0 = p->h_contig_read_off[contig];
    a.g0 = p->h_col_base[contig];
    a.n_reads = (int)R;
    a.n_parts = parts->n_parts;
    a.pstate = d_pst;
    a.tile_off = p->d_tile_off;
    a.tile_reads = p->d_tile_reads;
    a.read_start = p->d_read_start;
    a.read_end = p->d_read_end;
    a.row_base = p->d_row_base;
    a.codes = p->d_codes;
    a.k0 = p->d_k0;
    a.lut = (const HsRankLut*)ctx->d_rank_lut;
    a.out = d_out;
    HS_KERNEL(ctx, "partition_tables_kernel", partition_tables_kernel<<<n_cols, 128, kTablesSmem, ctx->stream>>>(a));
    HS_CUDA(ctx, hs_d2h(ctx, out, d_out, n_out));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    hs_free(ctx, d_pst);
    hs_free(ctx, d_pos);
    hs_free(ctx, d_out);
    return HSGPU_OK;
}

// the partitions as robust_filter_kernel reads them: per read one row of 2-bit states, 16 partitions per word,
// pwords words (1 = +1, 2 = -1, 3 = 0, 0 = absent or masked). The rows are scattered on the device from the
// partitions' entry lists: one thread per partition walks its entries in order (a read listed twice: the last entry
// counts), fields of different partitions in one word meet through atomics.
struct FilterPart {
    int32_t contig, local;  // the contig of a partition, its index inside the contig
};

__global__ void __launch_bounds__(128) filter_rows_kernel(int64_t n_parts_all, const FilterPart* __restrict__ meta,
                                                          const int64_t* __restrict__ ent_off,
                                                          const int32_t* __restrict__ read_idx,
                                                          const uint8_t* __restrict__ state2, const FilterDesc* __restrict__ desc,
                                                          uint32_t* __restrict__ rows) {
    const int64_t gp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= n_parts_all) return;
    const FilterPart pm = meta[gp];
    const FilterDesc d = desc[pm.contig];
    uint32_t* const base = rows + d.row_off + (pm.local >> 4);
    const int sh = 2 * (pm.local & 15);
    const int64_t e1 = ent_off[gp + 1];
    for (int64_t i = ent_off[gp]; i < e1; i++) {
        uint32_t* const w = base + (int64_t)read_idx[i] * d.pwords;
        atomicAnd(w, ~(3u << sh));
        atomicOr(w, (uint32_t)state2[i] << sh);
    }
}

static void filter_free(hsgpu_pileup* p) {
    hs_free(p->ctx, p->d_filter_block);
    p->have_parts = false;
}

// uploads the partitions of contigs [c0, c0 + n) (parts[i] belongs to contig c0 + i; every other contig of the
// pileup gets none) and keeps them with the pileup
static int filter_set(hsgpu_pileup* p, int c0, int n, const hsgpu_partitions* parts) {
    hsgpu_ctx* ctx = p->ctx;
    const int nc = p->n_contigs;
    std::vector<FilterDesc> desc((size_t)nc);
    int64_t row_words = 0;
    for (int c = 0; c < nc; c++) {
        FilterDesc& d = desc[c];
        memset(&d, 0, sizeof(d));
        const hsgpu_partitions* q = (c >= c0 && c < c0 + n) ? &parts[c - c0] : nullptr;
        if (!q || q->n_parts <= 0) continue;
        const int64_t R = p->h_contig_read_off[c + 1] - p->h_contig_read_off[c];
        d.n_parts = q->n_parts;
        d.pwords = ((q->n_parts + 127) / 128) * 8;
        d.row_off = row_words;
        row_words += R * d.pwords;
    }
    // staged for one upload: descriptors, per-partition (contig, index), entry offsets, read indices, 2-bit states
    int64_t n_parts_all = 0, n_ent = 0;
    for (int c = c0; c < c0 + n; c++) {
        if (desc[c].n_parts <= 0) continue;
        const hsgpu_partitions& q = parts[c - c0];
        n_parts_all += q.n_parts;
        n_ent += q.part_off[q.n_parts] - q.part_off[0];
    }
    static const bool timing = getenv("HSGPU_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    HsCarve hv;  // used for its offsets only: the pieces are cut out of the pinned staging block and of one device block
    FilterDesc* o_desc;
    FilterPart* o_meta;
    int64_t* o_off;
    int32_t* o_idx;
    uint8_t* o_st;
    hv.add(&o_desc, nc);
    hv.add(&o_meta, n_parts_all);
    hv.add(&o_off, n_parts_all + 1);
    hv.add(&o_idx, n_ent);
    hv.add(&o_st, n_ent);
    const size_t in_bytes = hv.total();
    uint8_t* h = reinterpret_cast<uint8_t*>(hs_host_stage(ctx, in_bytes));
    if (!h) HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu: pinned staging allocation failed");
    hv.place(h);
    memcpy(o_desc, desc.data(), sizeof(FilterDesc) * (size_t)nc);
    int64_t gp = 0, ge = 0;
    for (int c = c0; c < c0 + n; c++) {
        if (desc[c].n_parts <= 0) continue;
        const hsgpu_partitions& q = parts[c - c0];
        const int64_t R = p->h_contig_read_off[c + 1] - p->h_contig_read_off[c];
        const int64_t e0 = q.part_off[0];
        for (int k = 0; k < q.n_parts; k++) {
            o_meta[gp + k].contig = c;
            o_meta[gp + k].local = k;
            o_off[gp + k] = ge + (q.part_off[k] - e0);
        }
        const int64_t ne = q.part_off[q.n_parts] - e0;
        const int32_t* ri = q.read_idx + e0;
        const int16_t* st = q.state + e0;
        int32_t bad = 0;
        for (int64_t i = 0; i < ne; i++) {
            const int32_t r = ri[i];
            bad |= (r < 0) | (r >= R);
            o_idx[ge + i] = r;
            const int v = st[i];
            o_st[ge + i] = (uint8_t)(v == 1 ? 1 : v == -1 ? 2 : v == 0 ? 3 : 0);  // -2: masked
        }
        if (bad) HS_FAIL(ctx, HSGPU_ERR_ARG, "partition read index out of range");
        gp += q.n_parts;
        ge += ne;
    }
    o_off[gp] = ge;
    const auto t_staged = std::chrono::steady_clock::now();
    filter_free(p);
    const size_t in_pad = (in_bytes + 255) & ~(size_t)255;
    const size_t total = in_pad + (size_t)row_words * 4 + 256;
    HS_CUDA(ctx, hs_malloc_async(ctx, &p->d_filter_block, total));
    uint8_t* const dev = reinterpret_cast<uint8_t*>(p->d_filter_block);
    HS_CUDA(ctx, cudaMemcpyAsync(dev, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(dev + in_pad, 0, (size_t)row_words * 4 + 256, ctx->stream));
    p->d_fdesc = dev + ((uint8_t*)o_desc - h);
    p->d_frows = reinterpret_cast<uint32_t*>(dev + in_pad);
    if (n_parts_all > 0)
        HS_KERNEL(ctx, "filter_rows_kernel",
                  filter_rows_kernel<<<(unsigned)((n_parts_all + 127) / 128), 128, 0, ctx->stream>>>(
                      n_parts_all, reinterpret_cast<const FilterPart*>(dev + ((uint8_t*)o_meta - h)),
                      reinterpret_cast<const int64_t*>(dev + ((uint8_t*)o_off - h)),
                      reinterpret_cast<const int32_t*>(dev + ((uint8_t*)o_idx - h)), dev + ((uint8_t*)o_st - h),
                      reinterpret_cast<const FilterDesc*>(p->d_fdesc), p->d_frows));
    const auto t_queued = std::chrono::steady_clock::now();
    HS_CUDA(ctx, hs_stream_sync(ctx));  // the staging area is reused by the next call
    if (timing) {
        const auto t_end = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) {
            return std::chrono::duration<double, std::milli>(y - x).count();
        };
        fprintf(stderr, "[hsgpu timing] partitions_set: %lld partitions, %lld entries: stage %.3f ms, queue %.3f ms, sync %.3f ms\n",
                (long long)n_parts_all, (long long)n_ent, ms(t_begin, t_staged), ms(t_staged, t_queued), ms(t_queued, t_end));
    }
    p->have_parts = true;
    return HSGPU_OK;
}

// loops 3+4 over contigs [c0, c0 + n): kept positions packed per contig, off[i] .. off[i+1] for contig c0 + i
static int filter_run(hsgpu_pileup* p, int c0, int n, unsigned in_flag, int64_t capacity, int32_t* kept, int64_t* off) {
    hsgpu_ctx* ctx = p->ctx;
    const int64_t g_begin = p->h_col_base[c0], g_end = p->h_col_base[c0 + n];
    const int64_t ncols = g_end - g_begin;
    for (int i = 0; i <= n; i++) off[i] = 0;
    if (ncols <= 0) return HSGPU_OK;
    if (!p->d_filter_work) {  // active list, kept flags, kept list, counters: allocated once per pileup
        HsCarve cv;
        cv.add(&p->d_factive, p->n_cols);
        cv.add(&p->d_fkept, p->n_cols / 32 + 2);
        cv.add(&p->d_fkept_list, p->n_cols);
        cv.add(&p->d_fcounters, 12);
        cv.add(&p->d_foverflow, p->n_cols);
        cv.add(&p->d_fhdr, 2 * (int64_t)p->n_contigs + 2);
        HS_CUDA(ctx, cv.alloc(ctx, &p->d_filter_work));
    }
    FilterArgs a;
    a.n_contigs = p->n_contigs;
    a.in_flag = in_flag;
    a.desc = reinterpret_cast<const FilterDesc*>(p->d_fdesc);
    a.rows = p->d_frows;
    a.col_base = p->d_col_base;
    a.tile_base = p->d_tile_base;
    a.contig_read_off = p->d_contig_read_off;
    a.tile_off = p->d_tile_off;
    a.tile_reads = p->d_tile_reads;
    a.read_start = p->d_read_start;
    a.read_end = p->d_read_end;
    a.row_base = p->d_row_base;
    a.codes = p->d_codes;
    a.k0 = p->d_k0;
    a.k1 = p->d_k1;
    a.flags = p->d_flags;
    a.depth = p->d_depth;
    a.counts = p->d_counts;
    a.lut = (const HsRankLut*)ctx->d_rank_lut;
    a.g_begin = g_begin;
    a.g_end = g_end;
    a.active = p->d_factive;
    a.counters = p->d_fcounters;
    a.kept = p->d_fkept;
    HS_CUDA(ctx, cudaMemsetAsync(p->d_fcounters, 0, 12 * sizeof(unsigned int), ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(p->d_fkept + (g_begin >> 5), 0, sizeof(uint32_t) * (size_t)((g_end >> 5) - (g_begin >> 5) + 2), ctx->stream));
    HS_KERNEL(ctx, "filter_active_kernel", filter_active_kernel<<<(unsigned)((ncols + 3 + 4 * 256 - 1) / (4 * 256) + 1), 256, 0, ctx->stream>>>(a));
    // persistent warps pull active columns from a counter (their cost varies with depth and partition count)
    // HSGPU_FILTER_OCC=3: the build with 3 CTAs per SM (more registers, no spills) instead of 4, for A/B measurements
    // HSGPU_FILTER_LANES=0: the lane-per-read kernel for every column, for A/B measurements
    static const bool by_lanes = !getenv("HSGPU_FILTER_LANES") || atoi(getenv("HSGPU_FILTER_LANES")) != 0;
    static const int lanes_ctas = getenv("HSGPU_FILTER_CTAS") ? std::max(1, atoi(getenv("HSGPU_FILTER_CTAS"))) : RL_MINB;
    a.overflow = p->d_foverflow;
    a.from_overflow = 0;
    if (by_lanes) {
        HS_KERNEL(ctx, "robust_filter_lanes_kernel",
                  robust_filter_lanes_kernel<<<ctx->sm_count * lanes_ctas, 32 * RL_WARPS, 0, ctx->stream>>>(a));
        a.from_overflow = 1;  // columns with more distinct codes than the lanes keep counts for (rare)
        HS_KERNEL(ctx, "robust_filter_kernel<overflow>", robust_filter_kernel<false, 4><<<ctx->sm_count, 32 * RF_WARPS, 0, ctx->stream>>>(a));
        a.from_overflow = 0;
    } else {
        HS_KERNEL(ctx, "robust_filter_kernel", robust_filter_kernel<false, 4><<<ctx->sm_count * 4, 32 * RF_WARPS, 0, ctx->stream>>>(a));
    }
    // amplicon-deep columns (tiles with more than RF_CAP reads): only when the batch has such a tile
    {
        const int rc_max = hs_resolve_max_tile_reads(p);
        if (rc_max) return rc_max;
    }
    if (p->max_tile_reads > RF_CAP)
        HS_KERNEL(ctx, "robust_filter_kernel<deep>", robust_filter_kernel<true, 4><<<ctx->sm_count * 4, 32 * RF_WARPS, 0, ctx->stream>>>(a));
    HS_KERNEL(ctx, "kept_scan_kernel", kept_scan_kernel<<<n, 1024, 0, ctx->stream>>>(c0, p->d_col_base, p->d_contig_len, p->d_fkept,
                                                                                     p->d_fcounters + 2, p->n_cols, p->d_fkept_list, p->d_fhdr));
    int64_t* h_hdr = reinterpret_cast<int64_t*>(hs_host_stage(ctx, sizeof(int64_t) * (size_t)(2 * n)));
    if (!h_hdr) HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu: pinned staging allocation failed");
    HS_CUDA(ctx, cudaMemcpyAsync(h_hdr, p->d_fhdr, sizeof(int64_t) * (size_t)(2 * n), cudaMemcpyDeviceToHost, ctx->stream));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    std::vector<int64_t> start((size_t)n), count((size_t)n);
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        start[i] = h_hdr[2 * i];
        count[i] = h_hdr[2 * i + 1];
        off[i + 1] = off[i] + count[i];
        total += count[i];
    }
    if (total > capacity && kept) HS_FAIL(ctx, HSGPU_ERR_CAPACITY, "hsgpu_robust_filter: capacity too small");
    if (total > 0 && kept) {
        int32_t* h_list = reinterpret_cast<int32_t*>(hs_host_stage(ctx, sizeof(int32_t) * (size_t)total));
        if (!h_list) HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu: pinned staging allocation failed");
        HS_CUDA(ctx, cudaMemcpyAsync(h_list, p->d_fkept_list, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        for (int i = 0; i < n; i++)  // the contigs reserved their slices in the order their CTAs got there
            memcpy(kept + off[i], h_list + start[i], sizeof(int32_t) * (size_t)count[i]);
    }
    return HSGPU_OK;
}

int hsgpu_pileup_info(hsgpu_pileup* p, int64_t* info) {
    if (!p || !info) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < 8; i++) info[i] = 0;
    info[0] = p->n_cigar;
    info[1] = p->built ? p->codes_bytes : 0;
    info[2] = p->built ? p->tile_entries : 0;
    info[3] = p->built ? p->n_irregular : 0;
    if (p->d_filter_work) {
        unsigned int c[12];
        HS_CUDA(ctx, cudaMemcpyAsync(c, p->d_fcounters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        info[4] = (int64_t)c[0] + c[3];
        info[5] = (int64_t)c[8] | ((int64_t)c[9] << 32);
        info[6] = (int64_t)c[10] | ((int64_t)c[11] << 32);
        info[7] = c[2];
    }
    return HSGPU_OK;
}

int hsgpu_partitions_set(hsgpu_pileup* p, const hsgpu_partitions* parts) {
    if (!p || !parts) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_partitions_set: call hsgpu_pileup_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    return filter_set(p, 0, p->n_contigs, parts);
}

int hsgpu_robust_filter_all(hsgpu_pileup* p, int64_t capacity, int32_t* kept, int64_t* off) {
    if (!p || !off) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_robust_filter_all: call hsgpu_column_rank first");
    if (!p->have_parts) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_robust_filter_all: call hsgpu_partitions_set first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    return filter_run(p, 0, p->n_contigs, HS_FLAG_SUSPECT, capacity, kept, off);
}

int hsgpu_robust_filter(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_suspects,
                        const int32_t* suspect_pos, int32_t kept_capacity, int32_t* kept, int32_t* n_kept) {
    if (!p || !parts || contig < 0 || contig >= p->n_contigs || n_suspects < 0 || !n_kept) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_robust_filter: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    *n_kept = 0;
    const int64_t L = p->h_contig_len[contig];
    if (parts->n_parts == 0 || L == 0) return HSGPU_OK;  // :640-642: no partition, nothing is kept
    for (int i = 0; i < n_suspects; i++)
        if (suspect_pos[i] < 0 || suspect_pos[i] >= L) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_robust_filter: position out of range");
    int rc = filter_set(p, contig, 1, parts);
    if (rc) return rc;
    int32_t* d_pos = nullptr;
    HsTemps temps(ctx);
    temps.own(d_pos);
    HS_CUDA(ctx, hs_alloc(ctx, &d_pos, n_suspects));
    HS_CUDA(ctx, hs_h2d(ctx, d_pos, suspect_pos, n_suspects));
    const int64_t g0 = p->h_col_base[contig];
    if (n_suspects > 0)
        HS_KERNEL(ctx, "set_inlist_kernel", set_inlist_kernel<<<(n_suspects + 255) / 256, 256, 0, ctx->stream>>>(n_suspects, d_pos, g0, p->d_flags, 1));
    int64_t off[2] = {0, 0};
    std::vector<int32_t> tmp((size_t)std::max<int64_t>(L, 1));
    rc = filter_run(p, contig, 1, HS_FLAG_INLIST, L, tmp.data(), off);
    if (n_suspects > 0)
        HS_KERNEL(ctx, "set_inlist_kernel", set_inlist_kernel<<<(n_suspects + 255) / 256, 256, 0, ctx->stream>>>(n_suspects, d_pos, g0, p->d_flags, 0));
    HS_CUDA(ctx, hs_stream_sync(ctx));  // suspect_pos is caller memory
    hs_free(ctx, d_pos);
    filter_free(p);  // these partitions were the caller's for this call only
    if (rc) return rc;
    *n_kept = (int32_t)off[1];
    if (off[1] > kept_capacity) HS_FAIL(ctx, HSGPU_ERR_CAPACITY, "hsgpu_robust_filter: kept_capacity too small");
    if (kept) memcpy(kept, tmp.data(), sizeof(int32_t) * (size_t)off[1]);
    return HSGPU_OK;
}

}  // extern "C"
