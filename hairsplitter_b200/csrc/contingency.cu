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
#include <vector>

#include <atomic>

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

// ---- kernel B: loops 3 and 4 of keep_only_robust_variants over every column of a contig -----------
// One CTA per 128-column tile. Work items are (active column, active partition) pairs: a column is
// active when it is a suspect (loop 3) or passes the rescue pre-filter (loop 4); a partition is
// active for the tile when at least one of the tile's reads is present in it.
#define HS_FLAG_INLIST 16

struct FilterArgs {
    int contig;
    int64_t tile0, read0, g0;
    int L, n_reads, n_parts;
    const uint8_t* pstate;
    const int64_t* tile_off;
    const int32_t* tile_reads;
    const int32_t* read_start;
    const int32_t* read_end;
    const int64_t* row_base;
    const uint8_t* codes;
    const uint8_t* k0;
    const uint8_t* flags;
    const uint32_t* depth;
    const uint32_t* counts;   // c0,c1,c2 per column (column ranking)
    const uint8_t* pstate_t;  // [n_reads][npad]: the transpose of pstate, rows padded to 16 partitions
    int npad;
    const HsRankLut* lut;
    uint8_t* kept;  // [L]
};

__device__ __forceinline__ void ct_stage_rows(uint4* s_tile, int32_t* s_reads, const FilterArgs& a, int64_t list_off,
                                              int b0, int nrows, int q0, int tid) {
    for (int v = tid; v < nrows * (HS_TILE / 16); v += 128) {
        const int row = v >> 3, part = v & 7;
        const int32_t r = __ldg(a.tile_reads + list_off + b0 + row);
        const int s = __ldg(a.read_start + r) & ~(HS_ALIGN - 1);
        const int e = (__ldg(a.read_end + r) + HS_ALIGN - 1) & ~(HS_ALIGN - 1);
        const int qv = q0 + 16 * part;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (qv >= s && qv < e) val = __ldg(reinterpret_cast<const uint4*>(a.codes + __ldg(a.row_base + r) + qv));
        s_tile[v] = val;
    }
    for (int i = tid; i < nrows; i += 128) s_reads[i] = (int32_t)(a.tile_reads[list_off + b0 + i] - a.read0);
}

__global__ void __launch_bounds__(128) robust_filter_kernel(FilterArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint4* s_tile = reinterpret_cast<uint4*>(smem);                                          // CT_ROWS*128
    uint16_t* s_hist = reinterpret_cast<uint16_t*>(smem + CT_ROWS * HS_TILE);                // [125][128]
    uint8_t* s_order = smem + CT_ROWS * HS_TILE + HS_NCODES * 128 * 2;                       // [125][128]
    uint8_t* s_state = smem + CT_ROWS * HS_TILE + HS_NCODES * 128 * 3;                       // [CT_NPA][CT_SSTRIDE]
    int32_t* s_reads = reinterpret_cast<int32_t*>(s_state + CT_NPA * CT_SSTRIDE);            // [CT_ROWS]
    __shared__ int s_cols[HS_TILE];
    __shared__ int s_parts[CT_NPA];
    __shared__ int s_ncols, s_nparts;
    __shared__ unsigned s_keep[HS_TILE];  // per column: 1 = kept
    __shared__ unsigned s_pmask[4];       // partitions pb..pb+127 that hold one of the tile's reads

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q0 = blockIdx.x * HS_TILE;
    const int64_t tile = a.tile0 + blockIdx.x;
    const int64_t list_off = a.tile_off[tile];
    const int nlist = (int)(a.tile_off[tile + 1] - list_off);
    const unsigned char* s_bytes = reinterpret_cast<const unsigned char*>(s_tile);

    // active columns, in order
    if (tid == 0) s_ncols = 0;
    s_keep[tid] = 0;
    __syncthreads();
    {
        const int q = q0 + tid;
        const unsigned f = (q < a.L) ? a.flags[a.g0 + q] : 0u;
        // Loop 4 keeps a column only if n10 + n00 > 4 (:756): reads that carry the alternative allele. No code but
        // the column's own most frequent one (ref_base = k0) has more than c1 carriers, and when the reference's
        // char/unsigned char quirk makes ref_base its own alternative the two counters stay 0 -- so a column with
        // c1 <= 4 can only be kept by loop 3, i.e. if it is a suspect.
        bool act = (f & HS_FLAG_INLIST) != 0;
        if (!act && (f & HS_FLAG_RESCUE)) act = a.counts[3 * (a.g0 + q) + 1] > 4u;
        for (int w = 0; w < 4; w++) {
            if (wid == w) {
                const unsigned mk = __ballot_sync(0xffffffffu, act);
                const int base = s_ncols;
                if (act) s_cols[base + __popc(mk & ((1u << lane) - 1u))] = tid;
                __syncwarp();
                if (lane == 0) s_ncols = base + __popc(mk);
            }
            __syncthreads();
        }
    }
    const int ncols = s_ncols;
    if (ncols == 0 || nlist == 0 || a.n_parts == 0) {
        if (q0 + tid < a.L) a.kept[q0 + tid] = 0;
        return;
    }
    for (int i = tid; i < HS_NCODES * 128 * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(s_hist)[i] = 0;
    const bool single = nlist <= CT_ROWS;
    if (single) ct_stage_rows(s_tile, s_reads, a, list_off, 0, nlist, q0, tid);
    __syncthreads();

    for (int pb = 0; pb < a.n_parts; pb += 128) {
        // which of partitions pb..pb+127 have one of the tile's reads? One 16-byte load covers 16 partitions of a
        // read (pstate_t), every thread takes (read, vector) pairs, the answer is a 128-bit mask in shared memory
        __syncthreads();
        if (tid < 4) s_pmask[tid] = 0;
        __syncthreads();
        {
            const int nvec = min(8, (a.npad - pb) >> 4);
            for (int v = tid; v < nlist * nvec; v += 128) {
                const int i = v / nvec, j = v - i * nvec;
                const int32_t n = single ? s_reads[i] : (int32_t)(a.tile_reads[list_off + i] - a.read0);
                const uint4 x = __ldg(reinterpret_cast<const uint4*>(a.pstate_t + (int64_t)n * a.npad + pb + 16 * j));
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
                unsigned bits = 0;
#pragma unroll
                for (int b = 0; b < 16; b++) bits |= (((w[b >> 2] >> (8 * (b & 3))) & 3u) ? 1u : 0u) << b;
                if (bits) atomicOr(&s_pmask[j >> 1], bits << (16 * (j & 1)));
            }
        }
        __syncthreads();
        const int p = pb + tid;
        const bool present = p < a.n_parts && ((s_pmask[tid >> 5] >> (tid & 31)) & 1u);
        // chunks of up to CT_NPA active partitions
        int consumed = 0;  // number of active partitions (in thread order) already processed
        for (;;) {
            __syncthreads();
            if (tid == 0) s_nparts = 0;
            __syncthreads();
            // ordered compaction of the next CT_NPA active partitions
            for (int w = 0; w < 4; w++) {
                if (wid == w) {
                    const unsigned mk = __ballot_sync(0xffffffffu, present);
                    const int base = s_nparts;  // counts ALL active partitions seen so far in this sweep
                    const int my = base + __popc(mk & ((1u << lane) - 1u));
                    if (present && my >= consumed && my < consumed + CT_NPA) s_parts[my - consumed] = p;
                    __syncwarp();
                    if (lane == 0) s_nparts = base + __popc(mk);
                }
                __syncthreads();
            }
            const int total_active = s_nparts;
            const int npa = min(CT_NPA, total_active - consumed);
            if (npa <= 0) break;

            const int nitems = ncols * npa;
            for (int ib = 0; ib < nitems; ib += 128) {
                const int item = ib + tid;
                const bool valid = item < nitems;
                const int col = valid ? s_cols[item / npa] : 0;
                const int kk = valid ? item % npa : 0;
                const int q = q0 + col;
                const int ref = valid ? a.k0[a.g0 + q] : 0;
                const unsigned f = valid ? a.flags[a.g0 + q] : 0u;
                int m = 0, nb = 0, alt = ' ', nref = 0;
                int n00 = 0, n01 = 0, n10 = 0, n11 = 0;
                bool second = false;  // does this item need its 2x2 table?
                for (int pass = 0; pass < 2; pass++) {
                    for (int b0 = 0; b0 < nlist; b0 += CT_ROWS) {
                        const int nrows = min(CT_ROWS, nlist - b0);
                        if (!single) {
                            __syncthreads();
                            ct_stage_rows(s_tile, s_reads, a, list_off, b0, nrows, q0, tid);
                            __syncthreads();
                        }
                        if (!single || (ib == 0 && pass == 0)) {
                            // states of the chunk's partitions for the staged rows
                            if (single) __syncthreads();
                            for (int v = tid; v < npa * nrows; v += 128) {
                                const int k2 = v / nrows, row = v - k2 * nrows;
                                s_state[k2 * CT_SSTRIDE + row] = a.pstate[(int64_t)s_parts[k2] * a.n_reads + s_reads[row]];
                            }
                            __syncthreads();
                        }
                        if (valid) {
                            const uint8_t* st_row = s_state + kk * CT_SSTRIDE;
                            if (pass == 0) {
                                // histogram of the codes among the partition's reads, first-seen order kept. The
                                // column's own majority code (most of the rows) is counted in a register; the
                                // rows it holds with state +1 / -1 are n11 / n01 of the table (:893-949).
                                for (int row = 0; row < nrows; row++) {
                                    const int code = s_bytes[row * HS_TILE + col];
                                    const int sg = st_row[row] & 3;
                                    if (code && sg) {
                                        const int idx = code - HS_CODE0;
                                        if (code == ref) {
                                            if (nref == 0) s_order[(m++) * 128 + tid] = (uint8_t)idx;
                                            nref++;
                                            n11 += sg == 1;
                                            n01 += sg == 2;
                                        } else {
                                            const unsigned c = s_hist[idx * 128 + tid];
                                            if (c == 0) s_order[(m++) * 128 + tid] = (uint8_t)idx;
                                            s_hist[idx * 128 + tid] = (uint16_t)(c + 1);
                                        }
                                        nb++;
                                    }
                                }
                            } else if (second) {
                                for (int row = 0; row < nrows; row++) {
                                    const int code = s_bytes[row * HS_TILE + col];
                                    const int sg = st_row[row] & 3;
                                    if (code == alt && code != ref) {
                                        n10 += sg == 1;
                                        n00 += sg == 2;
                                    }
                                }
                            }
                        }
                    }
                    if (pass == 0 && valid && nb > 0) {
                        if (nref) s_hist[(ref - HS_CODE0) * 128 + tid] = (uint16_t)nref;
                        // A column that is not a suspect is only kept through loop 4, which needs n10 + n00 > 4
                        // (:756): at least 5 of the partition's reads on one code other than ref_base.
                        // (codes >= 128 never count as ref_base in the reference's comparison, :838: no shortcut there)
                        second = (f & HS_FLAG_INLIST) != 0 || ref >= 128;
                        if (!second) {
                            for (int k = 0; k < m && !second; k++) {
                                const int idx = s_order[k * 128 + tid];
                                second = idx + HS_CODE0 != ref && s_hist[idx * 128 + tid] > 4;
                            }
                        }
                        if (second) alt = hs_select_alt(s_order, s_hist, 128, tid, m, ref, a.lut);
                    }
                }
                if (valid) {
                    for (int k = 0; k < m; k++) s_hist[s_order[k * 128 + tid] * 128 + tid] = 0;
                    if (nb > 0 && second) {
                        const float chi = hs_chi_square(n00, n01, n10, n11);
                        bool keep = false;
                        // loop 3 (:721-738): suspects
                        if ((f & HS_FLAG_INLIST) &&
                            (double)(n00 + n01 + n10 + n11) > __dmul_rn(0.5, (double)a.depth[a.g0 + q]) && chi > 15.f)
                            keep = true;
                        // loop 4 (:745-764): rescue of every other column (also suspects that failed loop 3)
                        if ((f & HS_FLAG_RESCUE) && (double)chi > 20.0 && n10 + n00 > 4 && n01 + n11 > 4) keep = true;
                        if (keep) s_keep[col] = 1;
                    }
                }
            }
            consumed += npa;
            if (consumed >= total_active) break;
        }
    }
    __syncthreads();
    if (q0 + tid < a.L) a.kept[q0 + tid] = (uint8_t)s_keep[tid];
}

// ascending compaction of the kept columns: one CTA of 1024 threads; a round covers 16 columns per thread
// (one 16-byte load of the flags), block-scans the per-thread counts and writes the positions in order
__global__ void __launch_bounds__(1024) kept_scan_kernel(int L, const uint8_t* __restrict__ kept, int capacity,
                                                         int32_t* __restrict__ out, int32_t* __restrict__ n_out) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int n = 0;  // kept columns before this round (same value in every thread)
    for (int r0 = 0; r0 < L; r0 += 1024 * 16) {
        const int q0 = r0 + 16 * tid;
        uint32_t w[4] = {0, 0, 0, 0};
        if (q0 + 16 <= L) {  // d_kept is its own cudaMallocAsync block: 16-byte aligned
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(kept + q0));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            for (int j = 0; q0 + j < L && j < 16; j++) w[j >> 2] |= (kept[q0 + j] ? 1u : 0u) << (8 * (j & 3));
        }
        unsigned bits = 0;  // bit j = column q0 + j is kept
#pragma unroll
        for (int j = 0; j < 16; j++) bits |= (((w[j >> 2] >> (8 * (j & 3))) & 0xffu) ? 1u : 0u) << j;
        const int cnt = __popc(bits);
        const int incl = hs_warp_incl_scan(cnt, lane);
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int wi = hs_warp_incl_scan(s_warp[lane], lane);
            s_warp[lane] = wi - s_warp[lane];
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        int i = n + s_warp[wid] + incl - cnt;
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            if (i < capacity) out[i] = q0 + j;
            i++;
        }
        n += s_total;
        __syncthreads();
    }
    if (tid == 0) *n_out = n;
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
static const int kFilterSmem = CT_ROWS * HS_TILE + HS_NCODES * 128 * 3 + CT_NPA * CT_SSTRIDE + CT_ROWS * 4;

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
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    hs_free(ctx, d_pst);
    hs_free(ctx, d_pos);
    hs_free(ctx, d_out);
    return HSGPU_OK;
}

int hsgpu_robust_filter(hsgpu_pileup* p, int32_t contig, const hsgpu_partitions* parts, int32_t n_suspects,
                        const int32_t* suspect_pos, int32_t kept_capacity, int32_t* kept, int32_t* n_kept) {
    if (!p || !parts || contig < 0 || contig >= p->n_contigs || n_suspects < 0 || !n_kept) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_robust_filter: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    *n_kept = 0;
    const int64_t R = p->h_contig_read_off[contig + 1] - p->h_contig_read_off[contig];
    const int64_t L = p->h_contig_len[contig];
    if (parts->n_parts == 0 || L == 0) return HSGPU_OK;  // :640-642: no partition, nothing is kept
    for (int i = 0; i < n_suspects; i++)
        if (suspect_pos[i] < 0 || suspect_pos[i] >= L) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_robust_filter: position out of range");
    // with the transpose, rows padded to 16 partitions (the kernel's presence scan reads 16 partitions of a read at once)
    const int npad = (parts->n_parts + 15) & ~15;
    std::vector<uint8_t> pst, pst_t;
    int rc = build_pstate(ctx, parts, R, pst, &pst_t, npad);
    if (rc) return rc;
    uint8_t* d_pst = nullptr;
    uint8_t* d_pst_t = nullptr;
    int32_t* d_pos = nullptr;
    uint8_t* d_kept = nullptr;
    int32_t* d_list = nullptr;
    HS_CUDA(ctx, hs_alloc(ctx, &d_pst, (int64_t)pst.size()));
    HS_CUDA(ctx, hs_alloc(ctx, &d_pos, n_suspects));
    HS_CUDA(ctx, hs_alloc(ctx, &d_kept, L));
    HS_CUDA(ctx, hs_alloc(ctx, &d_list, (int64_t)kept_capacity + 1));
    HS_CUDA(ctx, hs_h2d(ctx, d_pst, pst.data(), (int64_t)pst.size()));
    HS_CUDA(ctx, hs_alloc(ctx, &d_pst_t, (int64_t)pst_t.size()));
    HS_CUDA(ctx, hs_h2d(ctx, d_pst_t, pst_t.data(), (int64_t)pst_t.size()));
    HS_CUDA(ctx, hs_h2d(ctx, d_pos, suspect_pos, n_suspects));
    const int64_t g0 = p->h_col_base[contig];
    if (n_suspects > 0) {
        HS_KERNEL(ctx, "set_inlist_kernel", set_inlist_kernel<<<(n_suspects + 255) / 256, 256, 0, ctx->stream>>>(n_suspects, d_pos, g0, p->d_flags, 1));
    }
    static std::atomic<bool> attr[64];  // per device: function attributes belong to the device's context
    if (!attr[ctx->device & 63]) {
        HS_CUDA(ctx, cudaFuncSetAttribute(robust_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFilterSmem));
        attr[ctx->device & 63] = true;
    }
    FilterArgs a;
    a.contig = contig;
    a.tile0 = p->h_tile_base[contig];
    a.read0 = p->h_contig_read_off[contig];
    a.g0 = g0;
    a.L = (int)L;
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
    a.flags = p->d_flags;
    a.depth = p->d_depth;
    a.counts = p->d_counts;
    a.pstate_t = d_pst_t;
    a.npad = npad;
    a.lut = (const HsRankLut*)ctx->d_rank_lut;
    a.kept = d_kept;
    const unsigned ntile = (unsigned)((L + HS_TILE - 1) / HS_TILE);
    HS_KERNEL(ctx, "robust_filter_kernel", robust_filter_kernel<<<ntile, 128, kFilterSmem, ctx->stream>>>(a));
    HS_KERNEL(ctx, "kept_scan_kernel", kept_scan_kernel<<<1, 1024, 0, ctx->stream>>>((int)L, d_kept, kept_capacity, d_list, d_list + kept_capacity));
    if (n_suspects > 0) {
        HS_KERNEL(ctx, "set_inlist_kernel", set_inlist_kernel<<<(n_suspects + 255) / 256, 256, 0, ctx->stream>>>(n_suspects, d_pos, g0, p->d_flags, 0));
    }
    HS_CUDA(ctx, hs_d2h(ctx, n_kept, d_list + kept_capacity, 1));
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int rc2 = HSGPU_OK;
    if (*n_kept > kept_capacity) {
        hs_set_error(ctx, "hsgpu_robust_filter: kept_capacity too small");
        rc2 = HSGPU_ERR_CAPACITY;
    } else if (kept) {
        HS_CUDA(ctx, hs_d2h(ctx, kept, d_list, *n_kept));
        HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    hs_free(ctx, d_pst);
    hs_free(ctx, d_pst_t);
    hs_free(ctx, d_pos);
    hs_free(ctx, d_kept);
    hs_free(ctx, d_list);
    return rc2;
}

}  // extern "C"
