// Per-column allele counting and ranking: the body of call_variants (reference
// src/call_variants.cpp:447-567), plus export of the pileup in the reference's column layout.
//
// One CTA per 128-column tile, one thread per column. The tile's rows (reads overlapping it, in
// ascending neighbour order) are staged through shared memory 64 rows at a time with aligned 16-byte
// loads; each thread walks down its column updating a private 125-bin histogram that lives in shared
// memory ([code][column] layout, so the common case "all reads agree" is conflict-free) and
// recording the order in which codes first appear -- that order is the insertion order of the
// reference's robin_hood map and decides ties (rank.cuh).
#include <algorithm>

#include <atomic>

#include "common.cuh"
#include "rank.cuh"

#define COL_ROWS 64  // rows staged per batch (export_kernel)

struct ColumnArgs {
    int64_t n_tiles;
    const int32_t* tile_contig;
    const int64_t* tile_base;
    const int64_t* col_base;
    const int32_t* contig_len;
    const int64_t* tile_off;
    const int32_t* tile_reads;
    const int32_t* read_start;
    const int32_t* read_end;
    const int64_t* row_base;
    const uint8_t* codes;
    const int32_t* min_reads;
    float auto_threshold;
    uint8_t* k0;
    uint8_t* k1;
    uint8_t* flags;
    uint32_t* counts;
    uint32_t* depth;
    unsigned long long* depth_sum;
    int32_t* error_flag;
    const HsRankLut* lut;
    // deferred tie resolution: columns the bucket table cannot decide leave their (code, count) list, in
    // first-seen order, in an arena; the few that do not fit are listed for a re-read of the column
    uint32_t* arena;
    unsigned int arena_words;
    unsigned int* counters;    // [0] arena words used, [1] arena items, [2] re-read items
    uint32_t* item_off;        // arena offset of every item
    int32_t* reread;           // global column ids
};

// stage rows [b0, b0+nrows) of the tile into s_tile[row][128]
__device__ __forceinline__ void stage_rows(uint4* s_tile, const int32_t* __restrict__ tile_reads, int64_t list_off,
                                           int b0, int nrows, int q0, const int32_t* __restrict__ read_start,
                                           const int32_t* __restrict__ read_end, const int64_t* __restrict__ row_base,
                                           const uint8_t* __restrict__ codes, int tid, int nthreads) {
    for (int v = tid; v < nrows * (HS_TILE / 16); v += nthreads) {
        const int row = v >> 3, part = v & 7;
        const int32_t r = __ldg(tile_reads + list_off + b0 + row);
        const int s = __ldg(read_start + r) & ~(HS_ALIGN - 1);
        const int e = (__ldg(read_end + r) + HS_ALIGN - 1) & ~(HS_ALIGN - 1);
        const int qv = q0 + 16 * part;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (qv >= s && qv < e) val = __ldg(reinterpret_cast<const uint4*>(codes + __ldg(row_base + r) + qv));
        s_tile[v] = val;
    }
}

// the suspect predicate of :525-528 without the spacing rule, and the rescue pre-filter of :751-752
__device__ __forceinline__ bool central_base_differs(int k0, int k1) {
    return (k0 % 5 != k1 % 5) && (((k1 - '!') % 5 != 4) || ((k1 / 5 % 5 != k0 % 5) && (k1 / 25 % 5 != k0 % 5)));
}

__device__ __forceinline__ void write_column(const ColumnArgs& a, int64_t g, int k0, int k1, unsigned c0, unsigned c1,
                                             unsigned c2, int mr) {
    a.k0[g] = (uint8_t)k0;
    a.k1[g] = (uint8_t)k1;
    a.counts[3 * g + 0] = c0;
    a.counts[3 * g + 1] = c1;
    a.counts[3 * g + 2] = c2;
    const bool cbd = central_base_differs(k0, k1);
    unsigned f = 0;
    if (cbd) f |= HS_FLAG_RESCUE | (c1 > 4u ? HS_FLAG_ACTIVE : 0u);
    if ((int)c1 > mr && ((int)c1 > (int)c2 * 5 || mr == 2) && cbd) {
        f |= HS_FLAG_CANDIDATE;
        if ((float)(int)c1 > __fmul_rn(a.auto_threshold, (float)(int)c0)) f |= HS_FLAG_AUTO;  // :531
    }
    a.flags[g] = (uint8_t)f;
}

// ---- the column histogram kernel -----------------------------------------------------------------------
// One CTA per 128-column tile, one thread per column. Row metadata (row pointer, which of the eight
// 16-byte vectors of the tile the read covers) is resolved once per 128 list entries; the rows then stream
// through a double-buffered shared-memory stage with cp.async (16 rows per batch, one 16-byte copy per
// thread, zero-fill where the read does not reach), so the loads of batch b+1 overlap the histogram
// updates of batch b. Histogram: u16 [126][128] in shared memory, bin 0 = "no cell here" so the update is
// branch-free; the first CR_ORD distinct codes of a column are kept in first-seen order.
#define CR_ROWS 16
#define CR_ORD 32
#define CR_BINS (HS_NCODES + 1)
#define CR_META 128

template <typename H>
struct SmemAcc {
    const uint8_t* order;   // [CR_ORD][128] codes in first-seen order
    const H* hist;          // [CR_BINS][128], bin = code - 32
    int col;
    __device__ __forceinline__ int key(int k) const { return order[k * HS_TILE + col]; }
    __device__ __forceinline__ unsigned count(int key) const { return hist[(key - 32) * HS_TILE + col]; }
};

// Histogram bins are bytes for tiles covered by at most CR_U8_MAX reads (every tile at ordinary sequencing
// depths: the histogram is what bounds the CTAs per SM) and 16-bit for deeper tiles (amplicons); the two
// instantiations are launched over the same grid and each CTA leaves at once when the tile is not its kind.
#define CR_U8_MAX 240
template <typename H>
constexpr int column_smem() {
    return 2 * CR_ROWS * HS_TILE + CR_BINS * HS_TILE * (int)sizeof(H) + (CR_ORD + 1) * HS_TILE + CR_META * 8 + CR_META;
}

__device__ __forceinline__ void cr_cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

template <typename H>
__global__ void __launch_bounds__(HS_TILE) column_rank_kernel(ColumnArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t list_off0 = a.tile_off[blockIdx.x];
    if (((int)(a.tile_off[blockIdx.x + 1] - list_off0) <= CR_U8_MAX) != (sizeof(H) == 1)) return;
    unsigned char* s_rows = smem;                                                        // 2 x 16 x 128 B
    H* s_hist = reinterpret_cast<H*>(smem + 2 * CR_ROWS * HS_TILE);                      // [126][128]
    uint8_t* s_order = smem + 2 * CR_ROWS * HS_TILE + CR_BINS * HS_TILE * (int)sizeof(H);  // [32][128] u8
    const uint8_t** s_ptr = reinterpret_cast<const uint8_t**>(s_order + (CR_ORD + 1) * HS_TILE);  // [128]; s_order has a spare row
    uint8_t* s_vmask = reinterpret_cast<uint8_t*>(s_ptr + CR_META);                      // [128]
    __shared__ unsigned long long s_depth;
    __shared__ __align__(16) uint32_t s_lut_words[HS_RANK_LUT_FAST_BYTES / 4];  // the part of HsRankLut hs_rank_fast reads
    __shared__ int s_scan[4];
    __shared__ unsigned int s_base[2];

    const int tid = threadIdx.x;
    // the column this thread owns: lane l of warp w takes column 4l + w, so that the 32 lanes of a warp sit in
    // 32 different shared-memory banks whatever bins they touch (byte histogram rows are 128 B = 32 banks)
    const int col = 4 * (tid & 31) + (tid >> 5);
    const int64_t tile = blockIdx.x;
    const int c = a.tile_contig[tile];
    const int q0 = (int)(tile - a.tile_base[c]) * HS_TILE;
    const int L = a.contig_len[c];
    const int64_t list_off = a.tile_off[tile];
    const int nlist = (int)(a.tile_off[tile + 1] - list_off);
    if (tid == 0) {
        s_depth = 0;
        if (nlist > 65000) atomicExch(a.error_flag, 1);  // u16 histogram bins; the reference's own loop is a `short` (:479)
    }
    {
        uint4* h4 = reinterpret_cast<uint4*>(s_hist);
        for (int i = tid; i < CR_BINS * HS_TILE * (int)sizeof(H) / 16; i += HS_TILE) h4[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < HS_RANK_LUT_FAST_BYTES / 4; i += HS_TILE)
            s_lut_words[i] = __ldg(reinterpret_cast<const uint32_t*>(a.lut) + i);
    }
    __syncthreads();
    s_hist[col] = 1;  // bin 0 ("no cell") never looks like a first sighting; set by the column's own thread (a tile
                      // without reads has no barrier between here and the read of this bin below)
    int m = 0, rows_done = 0;
    H* const hcol = s_hist + col;
    const int crow = tid >> 3, cpart = tid & 7;  // this thread's copy slot in a batch: row, 16-byte part
    for (int sb = 0; sb < nlist; sb += CR_META) {
        const int nmeta = min(CR_META, nlist - sb);
        __syncthreads();  // the previous super-batch's copies have all been issued and consumed
        if (tid < nmeta) {
            const int32_t r = __ldg(a.tile_reads + list_off + sb + tid);
            const int s = __ldg(a.read_start + r) & ~(HS_ALIGN - 1);
            const int e = (__ldg(a.read_end + r) + HS_ALIGN - 1) & ~(HS_ALIGN - 1);
            unsigned vm = 0;
#pragma unroll
            for (int p = 0; p < 8; p++) {
                const int qv = q0 + 16 * p;
                if (qv >= s && qv < e) vm |= 1u << p;
            }
            s_ptr[tid] = a.codes + __ldg(a.row_base + r) + q0;
            s_vmask[tid] = (uint8_t)vm;
        }
        __syncthreads();
        const int nb = (nmeta + CR_ROWS - 1) / CR_ROWS;
        rows_done += nb * CR_ROWS;  // the rows past the list's end are zero-filled: they count as "no cell"
        // issue batch 0
        {
            const int row = crow;
            const bool ok = row < nmeta && ((s_vmask[row] >> cpart) & 1);
            cr_cp_async16(s_rows + (crow * HS_TILE + 16 * cpart), ok ? (const void*)(s_ptr[row] + 16 * cpart) : (const void*)a.codes,
                          ok ? 16 : 0);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        for (int b = 0; b < nb; b++) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();  // batch b has landed for everyone; everyone is done with batch b-1
            if (b + 1 < nb) {
                const int row = (b + 1) * CR_ROWS + crow;
                const bool ok = row < nmeta && ((s_vmask[row] >> cpart) & 1);
                cr_cp_async16(s_rows + (((b + 1) & 1) * CR_ROWS * HS_TILE + crow * HS_TILE + 16 * cpart),
                              ok ? (const void*)(s_ptr[row] + 16 * cpart) : (const void*)a.codes, ok ? 16 : 0);
                asm volatile("cp.async.commit_group;\n" ::: "memory");
            }
            const unsigned char* colp = s_rows + (b & 1) * CR_ROWS * HS_TILE + col;
            int code[CR_ROWS];
#pragma unroll
            for (int row = 0; row < CR_ROWS; row++) code[row] = colp[row * HS_TILE];
#pragma unroll
            for (int row = 0; row < CR_ROWS; row++) {
                const int bin = max(code[row], 32) - 32;
                const unsigned int cnt = hcol[bin * HS_TILE];
                // first-seen order without a branch (a first sighting somewhere in the warp is the normal case, and the
                // divergent path cost every row its issue slots): the code always goes to slot m, which only becomes
                // valid when m moves on -- later rows overwrite the open slot until the next first sighting fills it.
                // Slot CR_ORD is the spare row for columns that have run out of slots (they are re-read anyway).
                s_order[min(m, CR_ORD) * HS_TILE + col] = (uint8_t)code[row];
                m += cnt == 0;
                hcol[bin * HS_TILE] = (H)(cnt + 1);
            }
        }
    }
    // ---- ranking. Ties are decided by the bucket table (rank.cuh); the columns it cannot decide are
    // deferred, so that no CTA waits for a straggling replay.
    const HsRankLut* const s_lut = reinterpret_cast<const HsRankLut*>(s_lut_words);
    const int q = q0 + col;
    const int mr = a.min_reads[c];
    const int64_t gbase = a.col_base[c];
    const unsigned int depth = (unsigned int)(rows_done + 1) - hcol[0];
    int defer = 0;  // 1: arena item, 2: re-read
    if (q < L) {
        if (m > CR_ORD) {
            defer = 2;
        } else {
            SmemAcc<H> acc{s_order, s_hist, col};
            int k0, k1;
            unsigned c0, c1, c2;
            if (hs_rank_fast(acc, m, s_lut, k0, k1, c0, c1, c2) != 0) defer = 1;
            else write_column(a, gbase + q, k0, k1, c0, c1, c2, mr);
        }
        a.depth[gbase + q] = depth;
    }
    {
        // arena space for the deferred columns of this tile: block-wide exclusive scan of the word counts
        const int lane = tid & 31, wid = tid >> 5;
        const int need = defer == 1 ? 2 + m : 0;
        const int incl = hs_warp_incl_scan(need, lane);
        const unsigned dm = __ballot_sync(0xffffffffu, defer == 1);
        if (lane == 31) s_scan[wid] = incl;
        __syncthreads();
        int woff = 0, total = 0, iwoff = 0, itotal = 0;
        for (int w = 0; w < 4; w++) {
            const int t = s_scan[w];
            if (w < wid) woff += t;
            total += t;
        }
        __syncthreads();
        if (lane == 0) s_scan[wid] = __popc(dm);
        __syncthreads();
        for (int w = 0; w < 4; w++) {
            const int t = s_scan[w];
            if (w < wid) iwoff += t;
            itotal += t;
        }
        if (tid == 0 && total > 0) {
            // counters[0] (words) and counters[1] (items) are one aligned 64-bit word: a single atomic reserves both
            // (every tile of the batch comes through here; the words stay far below 2^32, see arena_words)
            const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(a.counters),
                                                     (unsigned long long)(unsigned)total | ((unsigned long long)(unsigned)itotal << 32));
            s_base[0] = (unsigned)old;
            s_base[1] = (unsigned)(old >> 32);
        }
        __syncthreads();
        if (defer == 1) {
            const unsigned int off = s_base[0] + woff + incl - need;
            if (s_base[0] + (unsigned)total <= a.arena_words) {
                a.arena[off] = (uint32_t)(gbase + q);
                a.arena[off + 1] = (uint32_t)m;
                for (int k = 0; k < m; k++) {
                    const int key = s_order[k * HS_TILE + col];
                    a.arena[off + 2 + k] = ((uint32_t)hcol[(key - 32) * HS_TILE] << 8) | (uint32_t)key;
                }
                a.item_off[s_base[1] + iwoff + __popc(dm & ((1u << lane) - 1u))] = off;
            } else {
                a.item_off[s_base[1] + iwoff + __popc(dm & ((1u << lane) - 1u))] = 0xffffffffu;  // arena full
                defer = 2;
            }
        }
        if (defer == 2) a.reread[atomicAdd(a.counters + 2, 1u)] = (int32_t)(gbase + q);
    }
    // depthOfCoverage numerator (:486,565)
    unsigned long long d = (q < L) ? depth : 0;
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((tid & 31) == 0 && d) atomicAdd(&s_depth, d);
    __syncthreads();
    if (tid == 0 && s_depth) atomicAdd(a.depth_sum + c, s_depth);
}

// Deferred tie resolution, one thread per arena item: the hash-bit table (rank.cuh) decides most of them;
// what is left (about 1 % of the columns at 10 % error, 60x) is compacted into a list for the literal replay
// of the reference's map + sort, so that the replay runs with full warps.
struct ItemAcc {
    const uint8_t* order;
    const uint16_t* cnt;
    __device__ __forceinline__ int key(int k) const { return order[k]; }
    __device__ __forceinline__ unsigned count(int key) const { return cnt[key - HS_CODE0]; }
};

__device__ __forceinline__ int contig_of_column(const int64_t* __restrict__ col_base, int n_contigs, int64_t g) {
    int lo = 0, hi = n_contigs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (col_base[mid] <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// LITERAL = false: items 0 .. counters[1], hash-bit table, leftovers appended to lit_items (counters[3])
// LITERAL = true: the items listed in lit_items, literal replay
template <bool LITERAL>
__global__ void __launch_bounds__(128) column_rank_deferred_kernel(ColumnArgs a, int n_contigs, uint32_t* lit_items) {
    __shared__ HsRankLut s_lut;
    if (!LITERAL) {
        for (int i = threadIdx.x; i < (int)(sizeof(HsRankLut) / 4); i += blockDim.x)
            reinterpret_cast<uint32_t*>(&s_lut)[i] = reinterpret_cast<const uint32_t*>(a.lut)[i];
        __syncthreads();
    }
    const unsigned n = LITERAL ? a.counters[3] : a.counters[1];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned off = LITERAL ? lit_items[i] : a.item_off[i];
        if (off == 0xffffffffu) continue;  // went to the re-read list
        const int64_t g = a.arena[off];
        const int m = (int)a.arena[off + 1];
        uint16_t cnt[HS_NCODES];
        uint8_t order[CR_ORD];
        for (int k = 0; k < m; k++) {
            const uint32_t e = a.arena[off + 2 + k];
            order[k] = (uint8_t)(e & 0xff);
            cnt[(e & 0xff) - HS_CODE0] = (uint16_t)(e >> 8);
        }
        ItemAcc acc{order, cnt};
        int k0, k1;
        unsigned c0, c1, c2;
        if (LITERAL) {
            hs_rank_literal(acc, m, k0, k1, c0, c1, c2);
        } else if (hs_rank_hashbits(acc, m, &s_lut, k0, k1, c0, c1, c2) &&
                   hs_rank_slotorder(acc, m, &s_lut, k0, k1, c0, c1, c2)) {
            lit_items[atomicAdd(a.counters + 3, 1u)] = off;
            continue;
        }
        write_column(a, g, k0, k1, c0, c1, c2, a.min_reads[contig_of_column(a.col_base, n_contigs, g)]);
    }
}

// Deferred literal replay: one thread per listed column. The column is re-read from the pileup (a few
// dozen byte loads) in ascending read order, then the reference's map + sort are replayed (rank.cuh).
struct LiteralArgs {
    const int32_t* work;
    const unsigned int* n_work;  // re-read list (columns with more than CR_ORD codes, or arena overflow)
    int n_contigs;
    const int64_t* col_base;
    const int64_t* tile_base;
    const int64_t* tile_off;
    const int32_t* tile_reads;
    const int32_t* read_start;
    const int32_t* read_end;
    const int64_t* row_base;
    const uint8_t* codes;
    ColumnArgs col;
};

struct LocalAcc {
    const uint8_t* order;
    const uint16_t* cnt;
    __device__ __forceinline__ int key(int k) const { return order[k]; }
    __device__ __forceinline__ unsigned count(int key) const { return cnt[key - HS_CODE0]; }
};

__global__ void __launch_bounds__(128) column_rank_literal_kernel(LiteralArgs a) {
    const unsigned n = *a.n_work;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t g = a.work[i];
        int lo = 0, hi = a.n_contigs - 1;  // contig of this column
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (a.col_base[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const int c = lo;
        const int q = (int)(g - a.col_base[c]);
        const int64_t tile = a.tile_base[c] + q / HS_TILE;
        uint16_t cnt[HS_NCODES];
        uint8_t order[HS_NCODES];
        for (int k = 0; k < HS_NCODES; k++) cnt[k] = 0;
        int m = 0;
        for (int64_t l = a.tile_off[tile]; l < a.tile_off[tile + 1]; l++) {
            const int32_t r = a.tile_reads[l];
            if (a.read_start[r] <= q && q < a.read_end[r]) {
                const int code = a.codes[a.row_base[r] + q];
                if (cnt[code - HS_CODE0]++ == 0) order[m++] = (uint8_t)code;
            }
        }
        LocalAcc acc{order, cnt};
        int k0, k1;
        unsigned c0, c1, c2;
        hs_rank_literal(acc, m, k0, k1, c0, c1, c2);
        write_column(a.col, g, k0, k1, c0, c1, c2, a.col.min_reads[c]);
    }
}

// minimumNumberOfReadsToBeConsideredSuspect (:463-466) from generate_msa's float return value
__global__ void min_reads_kernel(int n_contigs, const unsigned long long* __restrict__ stats,
                                 const float* __restrict__ mean_error, int32_t* __restrict__ min_reads) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    float me;
    if (mean_error) {
        me = mean_error[c];
    } else {
        const unsigned long long dsum = stats[3 * c + 0];
        const float total_distance = (float)(dsum > 16777216ull ? 16777216ull : dsum);
        const double total_length = __dadd_rn(1.0, (double)stats[3 * c + 1]);
        me = (float)__ddiv_rn((double)total_distance, total_length);
    }
    min_reads[c] = ((double)me < 0.015) ? 3 : 5;
}

// The spacing rule `position - posoflastsnp > 5` (:470,529,535) is a greedy left-to-right scan, but its
// dependencies are short: a candidate with no other candidate in the 5 columns before it is accepted
// whatever happened earlier, so it starts an independent segment. One thread per column; the thread of a
// segment head walks its segment (it ends at the first gap of more than 5 candidate-free columns).
// Column 0 can never be accepted because posoflastsnp starts at -5 (:470).
__device__ __forceinline__ bool is_candidate(const uint8_t* __restrict__ flags, int64_t g0, int q) {
    return q > 0 && (flags[g0 + q] & HS_FLAG_CANDIDATE);
}

// SUS_TILES tiles per CTA (a CTA per tile is bound by the CTA launch rate: 39 000 tiny CTAs per step on config 2).
// The head thread of a segment also counts the accepted columns per tile (they may lie in later tiles).
#define SUS_TILES 8
__global__ void __launch_bounds__(HS_TILE * SUS_TILES) suspect_mark_kernel(int64_t n_tiles, const int32_t* __restrict__ tile_contig,
                                                                           const int64_t* __restrict__ tile_base,
                                                                           const int64_t* __restrict__ col_base,
                                                                           const int32_t* __restrict__ contig_len,
                                                                           uint8_t* __restrict__ flags,
                                                                           unsigned long long* __restrict__ tile_cnt) {
    const int64_t tile = (int64_t)blockIdx.x * SUS_TILES + (threadIdx.x / HS_TILE);
    if (tile >= n_tiles) return;
    const int c = tile_contig[tile];
    const int L = contig_len[c];
    const int64_t g0 = col_base[c];
    const int64_t t0 = tile_base[c];
    const int q = (int)(tile - t0) * HS_TILE + (threadIdx.x % HS_TILE);
    if (q >= L || !is_candidate(flags, g0, q)) return;
    for (int d = 1; d <= 5; d++)
        if (q - d > 0 && (flags[g0 + q - d] & HS_FLAG_CANDIDATE)) return;  // not a segment head
    flags[g0 + q] |= HS_FLAG_SUSPECT;
    int last = q, pc = q, p = q;
    int64_t cur_tile = tile;
    unsigned cur_n = 1;
    for (;;) {
        p++;
        if (p >= L || p - pc > 5) break;
        if (flags[g0 + p] & HS_FLAG_CANDIDATE) {
            if (p - last > 5) {
                flags[g0 + p] |= HS_FLAG_SUSPECT;
                last = p;
                const int64_t tl = t0 + p / HS_TILE;
                if (tl != cur_tile) {
                    atomicAdd(tile_cnt + cur_tile, (unsigned long long)cur_n);
                    cur_tile = tl;
                    cur_n = 0;
                }
                cur_n++;
            }
            pc = p;
        }
    }
    atomicAdd(tile_cnt + cur_tile, (unsigned long long)cur_n);
}

// ordered compaction of the accepted columns from the scanned per-tile counts; the first tile of a contig also
// writes the contig's total
__global__ void __launch_bounds__(HS_TILE * SUS_TILES) suspect_fill_kernel(int64_t n_tiles, const int32_t* __restrict__ tile_contig,
                                                                           const int64_t* __restrict__ tile_base,
                                                                           const int64_t* __restrict__ col_base,
                                                                           const int32_t* __restrict__ contig_len,
                                                                           const int64_t* __restrict__ suspect_base,
                                                                           const uint8_t* __restrict__ flags,
                                                                           const int64_t* __restrict__ tile_off,
                                                                           int32_t* __restrict__ suspect_pos,
                                                                           uint8_t* __restrict__ suspect_auto,
                                                                           int32_t* __restrict__ n_suspects) {
    __shared__ int s_warp[SUS_TILES * 4];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t tile = (int64_t)blockIdx.x * SUS_TILES + (tid / HS_TILE);
    const bool live = tile < n_tiles;
    int c = 0, L = 0, q = 0;
    int64_t g0 = 0;
    if (live) {
        c = tile_contig[tile];
        L = contig_len[c];
        g0 = col_base[c];
        q = (int)(tile - tile_base[c]) * HS_TILE + (tid % HS_TILE);
    }
    const unsigned f = (live && q < L) ? flags[g0 + q] : 0u;
    const bool acc = (f & HS_FLAG_SUSPECT) != 0;
    const unsigned m = __ballot_sync(0xffffffffu, acc);
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    if (acc) {
        int before = __popc(m & ((1u << lane) - 1u));
        for (int w = wid & ~3; w < wid; w++) before += s_warp[w];
        const int64_t i = suspect_base[c] + (tile_off[tile] - tile_off[tile_base[c]]) + before;
        suspect_pos[i] = q;
        suspect_auto[i] = (f & HS_FLAG_AUTO) ? 1 : 0;
    }
    if (live && (tid % HS_TILE) == 0 && tile == tile_base[c])
        n_suspects[c] = (int32_t)(tile_off[tile_base[c + 1]] - tile_off[tile]);
}

// all suspect lists of a batch, packed contig after contig (hsgpu_suspects_all). One CTA per contig; hdr =
// [error flag, off[0..nc], depth_sum[0..nc-1]] so that one copy brings every per-contig scalar to the host.
__global__ void __launch_bounds__(256) suspect_gather_kernel(int nc, const int32_t* __restrict__ n_suspects,
                                                             const int64_t* __restrict__ suspect_base,
                                                             const int32_t* __restrict__ src_pos,
                                                             const uint8_t* __restrict__ src_auto,
                                                             const unsigned long long* __restrict__ depth_sum,
                                                             const int32_t* __restrict__ error_flag,
                                                             int32_t* __restrict__ pos, uint8_t* __restrict__ is_auto,
                                                             int64_t* __restrict__ hdr) {
    __shared__ long long s_part[8];
    const int c = blockIdx.x, tid = threadIdx.x;
    long long before = 0;
    for (int i = tid; i < c; i += 256) before += n_suspects[i];
    before = hs_warp_sum64(before);
    if ((tid & 31) == 0) s_part[tid >> 5] = before;
    __syncthreads();
    before = 0;
    for (int w = 0; w < 8; w++) before += s_part[w];
    const int n = n_suspects[c];
    const int64_t src = suspect_base[c];
    for (int i = tid; i < n; i += 256) {
        pos[before + i] = src_pos[src + i];
        is_auto[before + i] = src_auto[src + i];
    }
    if (tid == 0) {
        hdr[1 + c] = before;
        hdr[nc + 2 + c] = (int64_t)depth_sum[c];
        if (c == nc - 1) {
            hdr[1 + nc] = before + n;
            hdr[0] = *error_flag;
        }
    }
}

// ---- export in the reference's column-major layout ------------------------------------------------
__global__ void __launch_bounds__(HS_TILE) export_kernel(int contig, int64_t tile0, const int64_t* __restrict__ col_base,
                                                         const int32_t* __restrict__ contig_len,
                                                         const int64_t* __restrict__ contig_read_off,
                                                         const int64_t* __restrict__ tile_off,
                                                         const int32_t* __restrict__ tile_reads,
                                                         const int32_t* __restrict__ read_start,
                                                         const int32_t* __restrict__ read_end,
                                                         const int64_t* __restrict__ row_base,
                                                         const uint8_t* __restrict__ codes,
                                                         const int64_t* __restrict__ col_off,
                                                         uint32_t* __restrict__ out_idx, uint8_t* __restrict__ out_code) {
    __shared__ uint4 s_tile[COL_ROWS * HS_TILE / 16];
    __shared__ int32_t s_reads[COL_ROWS];
    const int tid = threadIdx.x;
    const int64_t tile = tile0 + blockIdx.x;
    const int q0 = blockIdx.x * HS_TILE;
    const int L = contig_len[contig];
    const int64_t list_off = tile_off[tile];
    const int nlist = (int)(tile_off[tile + 1] - list_off);
    const int64_t g0 = col_base[contig];
    const int64_t r0 = contig_read_off[contig];
    const int q = q0 + tid;
    int64_t w = (q < L) ? col_off[g0 + q] - col_off[g0] : 0;
    const unsigned char* s_bytes = reinterpret_cast<const unsigned char*>(s_tile);
    for (int b0 = 0; b0 < nlist; b0 += COL_ROWS) {
        const int nrows = min(COL_ROWS, nlist - b0);
        __syncthreads();
        stage_rows(s_tile, tile_reads, list_off, b0, nrows, q0, read_start, read_end, row_base, codes, tid, HS_TILE);
        if (tid < nrows) s_reads[tid] = tile_reads[list_off + b0 + tid];
        __syncthreads();
        if (q < L) {
            for (int row = 0; row < nrows; row++) {
                const int code = s_bytes[row * HS_TILE + tid];
                if (code) {
                    out_idx[w] = (uint32_t)(s_reads[row] - r0);
                    out_code[w] = (uint8_t)code;
                    w++;
                }
            }
        }
    }
}

// one warp per requested column
__global__ void __launch_bounds__(256) extract_kernel(int n_cols, const int32_t* __restrict__ pos, int64_t tile0,
                                                      int64_t read0, const int64_t* __restrict__ tile_off,
                                                      const int32_t* __restrict__ tile_reads,
                                                      const int32_t* __restrict__ read_start,
                                                      const int32_t* __restrict__ read_end,
                                                      const int64_t* __restrict__ row_base,
                                                      const uint8_t* __restrict__ codes, const int64_t* __restrict__ off,
                                                      uint32_t* __restrict__ out_idx, uint8_t* __restrict__ out_code) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n_cols) return;
    const int q = pos[i];
    const int64_t tile = tile0 + q / HS_TILE;
    const int64_t l0 = tile_off[tile], l1 = tile_off[tile + 1];
    int64_t w = off[i];
    for (int64_t lb = l0; lb < l1; lb += 32) {
        const int64_t l = lb + lane;
        bool hit = false;
        int32_t r = 0;
        if (l < l1) {
            r = tile_reads[l];
            hit = read_start[r] <= q && q < read_end[r];
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int64_t o = w + __popc(m & ((1u << lane) - 1u));
            out_idx[o] = (uint32_t)(r - read0);
            out_code[o] = codes[row_base[r] + q];
        }
        w += __popc(m);
    }
}

__global__ void gather_depth_kernel(int n, const int32_t* __restrict__ pos, int64_t g0, const uint32_t* __restrict__ depth,
                                    int64_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = depth[g0 + pos[i]];
}


extern "C" {

int hsgpu_column_rank(hsgpu_pileup* p, const float* mean_error, float automatic_snp_threshold) {
    if (!p) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_column_rank: call hsgpu_pileup_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nc = p->n_contigs;
    if (!p->d_rank_block) {  // everything this stage keeps, out of one allocation
        HsCarve cv;
        cv.add(&p->d_k0, p->n_cols);
        cv.add(&p->d_k1, p->n_cols);
        cv.add(&p->d_flags, p->n_cols);
        cv.add(&p->d_counts, 3 * p->n_cols);
        cv.add(&p->d_depth, p->n_cols);
        cv.add(&p->d_min_reads, nc + 1);  // last entry = device error flag
        cv.add(&p->d_suspect_pos, p->h_suspect_base[nc]);
        cv.add(&p->d_suspect_auto, p->h_suspect_base[nc]);
        cv.add(&p->d_n_suspects, nc);
        cv.add(&p->d_depth_sum, nc);
        // d_work: [0..3] counters (arena words, arena items, re-read items), then the re-read list
        cv.add(&p->d_work, p->n_cols + 4);
        p->arena_words = (unsigned int)std::min<int64_t>(2 * p->n_cols + 1024, 0x7fffffff);
        cv.add(&p->d_arena, p->arena_words);
        cv.add(&p->d_item_off, 2 * p->n_cols + 512);  // items, then the literal list (at most one entry per item)
        cv.add(&p->d_tile_sus, p->n_tiles + 1);
        HS_CUDA(ctx, cv.alloc(ctx, &p->d_rank_block));
    }
    p->ranked = false;
    p->have_col_off = false;
    p->auto_threshold = automatic_snp_threshold;
    float* d_me = nullptr;
    HsTemps temps(ctx);
    temps.own(d_me);
    if (mean_error) {
        HS_CUDA(ctx, hs_alloc(ctx, &d_me, nc));
        HS_CUDA(ctx, hs_h2d(ctx, d_me, mean_error, nc));
    }
    HS_CUDA(ctx, cudaMemsetAsync(p->d_depth_sum, 0, sizeof(unsigned long long) * nc, ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(p->d_min_reads + nc, 0, sizeof(int32_t), ctx->stream));
    HS_KERNEL(ctx, "min_reads_kernel", min_reads_kernel<<<(nc + 127) / 128, 128, 0, ctx->stream>>>(nc, p->d_stats, d_me, p->d_min_reads));
    if (mean_error) {
        HS_CUDA(ctx, hs_stream_sync(ctx));  // mean_error is caller memory
        hs_free(ctx, d_me);
    }
    static std::atomic<bool> attr_set[64];  // per device: function attributes belong to the device's context
    if (!attr_set[ctx->device & 63]) {
        HS_CUDA(ctx, cudaFuncSetAttribute(column_rank_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          column_smem<uint16_t>()));
        attr_set[ctx->device & 63] = true;
    }
    if (p->n_tiles > 0) {
        ColumnArgs a;
        a.n_tiles = p->n_tiles;
        a.tile_contig = p->d_tile_contig;
        a.tile_base = p->d_tile_base;
        a.col_base = p->d_col_base;
        a.contig_len = p->d_contig_len;
        a.tile_off = p->d_tile_off;
        a.tile_reads = p->d_tile_reads;
        a.read_start = p->d_read_start;
        a.read_end = p->d_read_end;
        a.row_base = p->d_row_base;
        a.codes = p->d_codes;
        a.min_reads = p->d_min_reads;
        a.auto_threshold = automatic_snp_threshold;
        a.k0 = p->d_k0;
        a.k1 = p->d_k1;
        a.flags = p->d_flags;
        a.counts = p->d_counts;
        a.depth = p->d_depth;
        a.depth_sum = p->d_depth_sum;
        a.error_flag = p->d_min_reads + nc;
        a.lut = (const HsRankLut*)ctx->d_rank_lut;
        a.arena = p->d_arena;
        a.arena_words = p->arena_words;
        a.counters = reinterpret_cast<unsigned int*>(p->d_work);
        a.item_off = p->d_item_off;
        a.reread = p->d_work + 4;
        HS_CUDA(ctx, cudaMemsetAsync(p->d_work, 0, 4 * sizeof(int32_t), ctx->stream));
        HS_KERNEL(ctx, "column_rank_kernel", column_rank_kernel<uint8_t><<<(unsigned)p->n_tiles, HS_TILE, column_smem<uint8_t>(), ctx->stream>>>(a));
        {
            const int rc_max = hs_resolve_max_tile_reads(p);
            if (rc_max) return rc_max;
        }
        if (p->max_tile_reads > CR_U8_MAX)  // deep tiles (amplicons): 16-bit bins
            HS_KERNEL(ctx, "column_rank_kernel<u16>", column_rank_kernel<uint16_t><<<(unsigned)p->n_tiles, HS_TILE, column_smem<uint16_t>(), ctx->stream>>>(a));
        uint32_t* lit_items = p->d_item_off + p->n_cols;
        HS_KERNEL(ctx, "column_rank_deferred_kernel",
                  column_rank_deferred_kernel<false><<<ctx->sm_count * 16, 128, 0, ctx->stream>>>(a, nc, lit_items));  // every resident slot: the items are long, divergent and few per thread
        HS_KERNEL(ctx, "column_rank_deferred_kernel<literal>",
                  column_rank_deferred_kernel<true><<<ctx->sm_count * 4, 128, 0, ctx->stream>>>(a, nc, lit_items));
        LiteralArgs la;
        la.work = a.reread;
        la.n_work = a.counters + 2;
        la.n_contigs = nc;
        la.col_base = p->d_col_base;
        la.tile_base = p->d_tile_base;
        la.tile_off = p->d_tile_off;
        la.tile_reads = p->d_tile_reads;
        la.read_start = p->d_read_start;
        la.read_end = p->d_read_end;
        la.row_base = p->d_row_base;
        la.codes = p->d_codes;
        la.col = a;
        HS_KERNEL(ctx, "column_rank_literal_kernel",
                  column_rank_literal_kernel<<<ctx->sm_count * 2, 128, 0, ctx->stream>>>(la));
    }
    if (p->n_tiles > 0) {
        const unsigned sus_grid = (unsigned)((p->n_tiles + SUS_TILES - 1) / SUS_TILES);
        HS_CUDA(ctx, cudaMemsetAsync(p->d_tile_sus, 0, sizeof(int64_t) * (size_t)(p->n_tiles + 1), ctx->stream));
        HS_CUDA(ctx, cudaMemsetAsync(p->d_n_suspects, 0, sizeof(int32_t) * nc, ctx->stream));  // contigs without tiles
        HS_KERNEL(ctx, "suspect_mark_kernel", suspect_mark_kernel<<<sus_grid, HS_TILE * SUS_TILES, 0, ctx->stream>>>(
            p->n_tiles, p->d_tile_contig, p->d_tile_base, p->d_col_base, p->d_contig_len, p->d_flags,
            reinterpret_cast<unsigned long long*>(p->d_tile_sus)));
        int rc = hs_exclusive_scan_i64(ctx, p->d_tile_sus, p->d_tile_sus, p->n_tiles, p->d_tile_sus + p->n_tiles);
        if (rc) return rc;
        HS_KERNEL(ctx, "suspect_fill_kernel", suspect_fill_kernel<<<sus_grid, HS_TILE * SUS_TILES, 0, ctx->stream>>>(
            p->n_tiles, p->d_tile_contig, p->d_tile_base, p->d_col_base, p->d_contig_len, p->d_suspect_base, p->d_flags,
            p->d_tile_sus, p->d_suspect_pos, p->d_suspect_auto, p->d_n_suspects));
    } else {
        HS_CUDA(ctx, cudaMemsetAsync(p->d_n_suspects, 0, sizeof(int32_t) * nc, ctx->stream));
    }
    p->ranked = true;
    return HSGPU_OK;
}

int hsgpu_column_counts(hsgpu_pileup* p, int32_t* n_suspects, int64_t* depth_sum) {
    if (!p) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_column_counts: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t err = 0;
    HS_CUDA(ctx, hs_d2h(ctx, &err, p->d_min_reads + p->n_contigs, 1));
    if (n_suspects) HS_CUDA(ctx, hs_d2h(ctx, n_suspects, p->d_n_suspects, p->n_contigs));
    if (depth_sum) HS_CUDA(ctx, hs_d2h(ctx, (unsigned long long*)depth_sum, p->d_depth_sum, p->n_contigs));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    if (err) HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_column_rank: more than 65000 reads over one 128-column tile");
    return HSGPU_OK;
}

int hsgpu_suspects(hsgpu_pileup* p, int32_t contig, int32_t capacity, int32_t* pos, uint8_t* is_automatic) {
    if (!p || contig < 0 || contig >= p->n_contigs) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_suspects: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t n = 0;
    HS_CUDA(ctx, hs_d2h(ctx, &n, p->d_n_suspects + contig, 1));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    if (n > capacity) HS_FAIL(ctx, HSGPU_ERR_CAPACITY, "hsgpu_suspects: capacity too small");
    if (pos) HS_CUDA(ctx, hs_d2h(ctx, pos, p->d_suspect_pos + p->h_suspect_base[contig], n));
    if (is_automatic) HS_CUDA(ctx, hs_d2h(ctx, is_automatic, p->d_suspect_auto + p->h_suspect_base[contig], n));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    return HSGPU_OK;
}

int hsgpu_suspects_all(hsgpu_pileup* p, int64_t capacity, int32_t* pos, uint8_t* is_automatic, int64_t* off,
                       int64_t* depth_sum) {
    if (!p || !off) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_suspects_all: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nc = p->n_contigs;
    const int64_t cap_all = p->h_suspect_base[nc];
    int32_t* d_pos = nullptr;
    uint8_t* d_auto = nullptr;
    int64_t* d_hdr = nullptr;
    HsTemps temps(ctx);
    temps.own(d_pos, d_auto, d_hdr);
    HS_CUDA(ctx, hs_alloc(ctx, &d_pos, cap_all));
    HS_CUDA(ctx, hs_alloc(ctx, &d_auto, cap_all));
    HS_CUDA(ctx, hs_alloc(ctx, &d_hdr, 2 * nc + 2));
    HS_KERNEL(ctx, "suspect_gather_kernel", suspect_gather_kernel<<<nc, 256, 0, ctx->stream>>>(
        nc, p->d_n_suspects, p->d_suspect_base, p->d_suspect_pos, p->d_suspect_auto, p->d_depth_sum,
        p->d_min_reads + nc, d_pos, d_auto, d_hdr));
    const size_t hdr_bytes = sizeof(int64_t) * (size_t)(2 * nc + 2);
    int64_t* h_hdr = reinterpret_cast<int64_t*>(hs_host_stage(ctx, hdr_bytes));
    if (!h_hdr) HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu_suspects_all: pinned staging allocation failed");
    HS_CUDA(ctx, cudaMemcpyAsync(h_hdr, d_hdr, hdr_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    const int64_t err = h_hdr[0];
    std::copy(h_hdr + 1, h_hdr + nc + 2, off);
    if (depth_sum) std::copy(h_hdr + nc + 2, h_hdr + 2 * nc + 2, depth_sum);
    const int64_t total = off[nc];
    int rc = HSGPU_OK;
    if (err) {
        hs_set_error(ctx, "hsgpu_column_rank: more than 65000 reads over one 128-column tile");
        rc = HSGPU_ERR_LIMIT;
    } else if (total > capacity && (pos || is_automatic)) {
        hs_set_error(ctx, "hsgpu_suspects_all: capacity too small");
        rc = HSGPU_ERR_CAPACITY;
    } else if (total > 0 && (pos || is_automatic)) {
        // both lists through the pinned staging area: one synchronisation, then plain copies into the caller's arrays
        const size_t pos_bytes = pos ? sizeof(int32_t) * (size_t)total : 0;
        uint8_t* h = reinterpret_cast<uint8_t*>(hs_host_stage(ctx, pos_bytes + (size_t)total));
        if (!h) HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu_suspects_all: pinned staging allocation failed");
        if (pos) HS_CUDA(ctx, cudaMemcpyAsync(h, d_pos, pos_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (is_automatic) HS_CUDA(ctx, cudaMemcpyAsync(h + pos_bytes, d_auto, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        if (pos) memcpy(pos, h, pos_bytes);
        if (is_automatic) memcpy(is_automatic, h + pos_bytes, (size_t)total);
    }
    hs_free(ctx, d_pos);
    hs_free(ctx, d_auto);
    hs_free(ctx, d_hdr);
    return rc;
}

int hsgpu_column_summary(hsgpu_pileup* p, int32_t contig, uint8_t* ref_base, uint8_t* second_base, uint32_t* counts,
                         uint32_t* depth) {
    if (!p || contig < 0 || contig >= p->n_contigs) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_column_summary: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t g0 = p->h_col_base[contig], L = p->h_contig_len[contig];
    if (ref_base) HS_CUDA(ctx, hs_d2h(ctx, ref_base, p->d_k0 + g0, L));
    if (second_base) HS_CUDA(ctx, hs_d2h(ctx, second_base, p->d_k1 + g0, L));
    if (counts) HS_CUDA(ctx, hs_d2h(ctx, counts, p->d_counts + 3 * g0, 3 * L));
    if (depth) HS_CUDA(ctx, hs_d2h(ctx, depth, p->d_depth + g0, L));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    return HSGPU_OK;
}

static int ensure_col_off(hsgpu_pileup* p) {
    hsgpu_ctx* ctx = p->ctx;
    if (p->have_col_off) return HSGPU_OK;
    if (!p->d_col_off) HS_CUDA(ctx, hs_alloc(ctx, &p->d_col_off, p->n_cols + 1));
    int rc = hs_exclusive_scan_u32_to_i64(ctx, p->d_depth, p->d_col_off, p->n_cols, p->d_col_off + p->n_cols);
    if (rc) return rc;
    p->have_col_off = true;
    return HSGPU_OK;
}

int hsgpu_pileup_export(hsgpu_pileup* p, int32_t contig, int64_t cell_capacity, int64_t* col_off, uint32_t* read_idx,
                        uint8_t* code) {
    if (!p || contig < 0 || contig >= p->n_contigs || !col_off) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pileup_export: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_col_off(p);
    if (rc) return rc;
    const int64_t g0 = p->h_col_base[contig], L = p->h_contig_len[contig];
    HS_CUDA(ctx, hs_d2h(ctx, col_off, p->d_col_off + g0, L + 1));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    const int64_t first = col_off[0];
    for (int64_t i = 0; i <= L; i++) col_off[i] -= first;
    const int64_t n = col_off[L];
    if (n > cell_capacity || !read_idx || !code) {
        if (n == 0) return HSGPU_OK;
        HS_FAIL(ctx, HSGPU_ERR_CAPACITY, "hsgpu_pileup_export: cell_capacity too small");
    }
    if (n == 0) return HSGPU_OK;
    uint32_t* d_idx = nullptr;
    uint8_t* d_code = nullptr;
    HsTemps temps(ctx);
    temps.own(d_idx, d_code);
    HS_CUDA(ctx, hs_alloc(ctx, &d_idx, n));
    HS_CUDA(ctx, hs_alloc(ctx, &d_code, n));
    const int64_t ntile = (L + HS_TILE - 1) / HS_TILE;
    HS_KERNEL(ctx, "export_kernel", export_kernel<<<(unsigned)ntile, HS_TILE, 0, ctx->stream>>>(contig, p->h_tile_base[contig], p->d_col_base,
                                                               p->d_contig_len, p->d_contig_read_off, p->d_tile_off,
                                                               p->d_tile_reads, p->d_read_start, p->d_read_end,
                                                               p->d_row_base, p->d_codes, p->d_col_off, d_idx, d_code));
    HS_CUDA(ctx, hs_d2h(ctx, read_idx, d_idx, n));
    HS_CUDA(ctx, hs_d2h(ctx, code, d_code, n));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    hs_free(ctx, d_idx);
    hs_free(ctx, d_code);
    return HSGPU_OK;
}

int hsgpu_pileup_extract_columns(hsgpu_pileup* p, int32_t contig, int32_t n_cols, const int32_t* pos,
                                 int64_t cell_capacity, int64_t* off, uint32_t* read_idx, uint8_t* code) {
    if (!p || contig < 0 || contig >= p->n_contigs || n_cols < 0 || !off) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->ranked) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pileup_extract_columns: call hsgpu_column_rank first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    off[0] = 0;
    if (n_cols == 0) return HSGPU_OK;
    const int64_t L = p->h_contig_len[contig];
    for (int i = 0; i < n_cols; i++)
        if (pos[i] < 0 || pos[i] >= L) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_extract_columns: position out of range");
    int32_t* d_pos = nullptr;
    int64_t* d_off = nullptr;
    HsTemps temps(ctx);
    temps.own(d_pos, d_off);
    HS_CUDA(ctx, hs_alloc(ctx, &d_pos, n_cols));
    HS_CUDA(ctx, hs_alloc(ctx, &d_off, n_cols + 1));
    HS_CUDA(ctx, hs_h2d(ctx, d_pos, pos, n_cols));
    HS_KERNEL(ctx, "gather_depth_kernel", gather_depth_kernel<<<(n_cols + 255) / 256, 256, 0, ctx->stream>>>(n_cols, d_pos, p->h_col_base[contig], p->d_depth,
                                                                       d_off));
    HS_CUDA(ctx, hs_d2h(ctx, off + 1, d_off, n_cols));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    for (int i = 0; i < n_cols; i++) off[i + 1] += off[i];  // depths -> offsets (n_cols is small)
    const int64_t n = off[n_cols];
    int rc = HSGPU_OK;
    if (n > cell_capacity || !read_idx || !code) {
        if (n > 0) {
            hs_set_error(ctx, "hsgpu_pileup_extract_columns: cell_capacity too small");
            rc = HSGPU_ERR_CAPACITY;
        }
    } else if (n > 0) {
        uint32_t* d_idx = nullptr;
        uint8_t* d_code = nullptr;
        HsTemps cell_temps(ctx);
        cell_temps.own(d_idx, d_code);
        HS_CUDA(ctx, hs_alloc(ctx, &d_idx, n));
        HS_CUDA(ctx, hs_alloc(ctx, &d_code, n));
        HS_CUDA(ctx, hs_h2d(ctx, d_off, off, n_cols + 1));
        HS_KERNEL(ctx, "extract_kernel", extract_kernel<<<(n_cols + 7) / 8, 256, 0, ctx->stream>>>(n_cols, d_pos, p->h_tile_base[contig],
                                                                  p->h_contig_read_off[contig], p->d_tile_off,
                                                                  p->d_tile_reads, p->d_read_start, p->d_read_end,
                                                                  p->d_row_base, p->d_codes, d_off, d_idx, d_code));
        HS_CUDA(ctx, hs_d2h(ctx, read_idx, d_idx, n));
        HS_CUDA(ctx, hs_d2h(ctx, code, d_code, n));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        hs_free(ctx, d_idx);
        hs_free(ctx, d_code);
    }
    HS_CUDA(ctx, hs_stream_sync(ctx));
    hs_free(ctx, d_pos);
    hs_free(ctx, d_off);
    return rc;
}

}  // extern "C"
