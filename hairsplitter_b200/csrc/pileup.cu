// Pileup construction: the body of generate_msa (reference src/call_variants.cpp:163-364) on the GPU.
//
// Layout in HBM. The reference keeps a vector<Column> (column-major, (readIdx u32, code u8) pairs,
// ~5 B/cell, pointer-chasing). Here the pileup is a ragged, READ-major byte matrix: read r owns one
// contiguous row of codes, one byte per contig column it covers, addressed codes[row_base[r] + q]
// with row_base % 16 == 0, zero-padded to 16-byte boundaries in column space. Rows are written with
// fully coalesced stores by the warp that walks the read's CIGAR, need no read index per cell
// (1 B/cell instead of 5), and a 128-column tile of any row is eight aligned 16-byte vectors, which
// is what the column kernels (column.cu, contingency.cu) stage through shared memory. A per-tile
// index lists the reads overlapping each tile in ascending neighbour order, which is exactly the
// in-column order of the reference.
#include <algorithm>
#include <cstring>

#include "common.cuh"

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

__device__ __forceinline__ bool op_consumes_q(int ty) { return ty == OP_M || ty == OP_EQ || ty == OP_X || ty == OP_D; }
__device__ __forceinline__ bool op_consumes_t(int ty) {
    return ty == OP_M || ty == OP_EQ || ty == OP_X || ty == OP_I || ty == OP_S || ty == OP_H;
}

// ---- K1: contig span of every read (sum of M/=/X/D lengths, clipped at the contig end, :217) ----
__global__ void __launch_bounds__(256) span_kernel(int64_t n_reads, const uint32_t* __restrict__ cigar,
                                                   const int64_t* __restrict__ cigar_off,
                                                   const int32_t* __restrict__ read_start,
                                                   const int32_t* __restrict__ read_contig,
                                                   const int32_t* __restrict__ contig_len,
                                                   int32_t* __restrict__ read_end, int64_t* __restrict__ row_alloc) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_reads) return;
    long long sum = 0;
    for (int64_t k = cigar_off[r] + lane; k < cigar_off[r + 1]; k += 32) {
        uint32_t op = __ldg(cigar + k);
        if (op_consumes_q((int)(op & 15))) sum += op >> 4;
    }
    sum = hs_warp_sum64(sum);
    if (lane == 0) {
        const int L = contig_len[read_contig[r]];
        const int start = read_start[r];
        long long end = start;
        if (start < L) end = (start + sum < (long long)L) ? start + sum : L;
        read_end[r] = (int)end;
        long long alloc = 0;
        if (end > start) alloc = ((end + HS_ALIGN - 1) & ~(long long)(HS_ALIGN - 1)) - (start & ~(HS_ALIGN - 1));
        row_alloc[r] = alloc;
    }
}

// ---- K3: the CIGAR walk. One warp per read. ------------------------------------------------------
// Ops are taken 32 at a time (one per lane); warp scans give every op its first contig column, read
// offset and position in the expanded alignment; the expanded alignment is then processed 32
// positions at a time, each lane locating its op by binary search in shared memory. The 3-mer
// context of a cell is the two previously pushed symbols (M/=/X/I push the read base, D pushes '-',
// S/H push nothing, :234-236,283-285,332-334), fetched from the lanes below via ballot + shuffle.
struct PileupArgs {
    int64_t n_reads;
    const int32_t* contig_len;
    const uint32_t* contig_bases;
    const int64_t* contig_word_off;
    const int32_t* read_contig;
    const uint32_t* read_bases;
    const int64_t* read_word_off;
    const int32_t* read_len;
    const uint32_t* cigar;
    const int64_t* cigar_off;
    const int32_t* read_start;
    const uint8_t* read_strand;
    const int32_t* read_end;
    const int64_t* row_off;
    int64_t* row_base;
    uint8_t* codes;
    unsigned long long* stats;
};

__global__ void __launch_bounds__(256) pileup_kernel(PileupArgs a) {
    __shared__ int s_e[8][33];
    __shared__ int s_q[8][32];
    __shared__ int s_t[8][32];
    __shared__ int s_ty[8][32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + wid;
    if (r >= a.n_reads) return;  // whole warps leave together; only __syncwarp is used below
    const int c = a.read_contig[r];
    const int L = a.contig_len[c];
    const uint32_t* __restrict__ cb = a.contig_bases + a.contig_word_off[c];
    const uint32_t* __restrict__ rb = a.read_bases + a.read_word_off[r];
    const int rlen = a.read_len[r];
    const int strand = a.read_strand[r];
    const int start = a.read_start[r];
    const int end = a.read_end[r];
    const int64_t row_base = a.row_off[r] - (int64_t)(start & ~(HS_ALIGN - 1));
    if (lane == 0) a.row_base[r] = row_base;
    uint8_t* __restrict__ row = a.codes + row_base;
    if (end > start) {  // zero the alignment pads so that tiles can be staged with whole 16-byte vectors
        for (int q = (start & ~(HS_ALIGN - 1)) + lane; q < start; q += 32) row[q] = 0;
        const int b = (end + HS_ALIGN - 1) & ~(HS_ALIGN - 1);
        for (int q = end + lane; q < b; q += 32) row[q] = 0;
    }
    const int64_t k0 = a.cigar_off[r], k1 = a.cigar_off[r + 1];
    int q = start, t = 0;
    int carry1 = 2, carry2 = 1;  // context 'A','C','G': most recent = G, before it C (:212-214)
    unsigned int dist = 0, alen = 0;
    for (int64_t kb = k0; kb < k1 && q < L; kb += 32) {
        const int64_t k = kb + lane;
        uint32_t op = (k < k1) ? __ldg(a.cigar + k) : (uint32_t)OP_P;
        const int len = (int)(op >> 4), ty = (int)(op & 15);
        const bool cq = op_consumes_q(ty), ct = op_consumes_t(ty);
        const int qa = cq ? len : 0, ta = ct ? len : 0, ea = (cq || ct) ? len : 0;
        const int qi = hs_warp_incl_scan(qa, lane), ti = hs_warp_incl_scan(ta, lane), ei = hs_warp_incl_scan(ea, lane);
        s_q[wid][lane] = q + qi - qa;
        s_t[wid][lane] = t + ti - ta;
        s_e[wid][lane] = ei - ea;
        s_ty[wid][lane] = ty;
        const int E = __shfl_sync(0xffffffffu, ei, 31);
        __syncwarp();
        for (int e0 = 0; e0 < E; e0 += 32) {
            const int e = e0 + lane;
            int lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; it++) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_e[wid][mid] <= e) lo = mid; else hi = mid - 1;
            }
            const int off = e - s_e[wid][lo];
            const int oty = s_ty[wid][lo];
            const bool ocq = op_consumes_q(oty), oct = op_consumes_t(oty);
            const int qp = s_q[wid][lo] + (ocq ? off : 0);
            const int tp = s_t[wid][lo] + (oct ? off : 0);
            const bool active = (e < E) && (qp < L);
            const bool push = active && (ocq || oty == OP_I);
            int sym = 0;
            if (push) {
                if (oty == OP_D) sym = 4;
                else if (tp < rlen) {  // a CIGAR longer than the read is malformed; the reference reads past the string
                    sym = strand ? hs_base2(rb, tp) : 3 - hs_base2(rb, (int64_t)rlen - 1 - tp);
                }
            }
            const unsigned pm = __ballot_sync(0xffffffffu, push);
            const unsigned below = pm & ((1u << lane) - 1u);
            const int l1 = below ? 31 - __clz(below) : -1;
            const unsigned below2 = (l1 >= 0) ? (below & ~(1u << l1)) : 0u;
            const int l2 = below2 ? 31 - __clz(below2) : -1;
            const int s1 = __shfl_sync(0xffffffffu, sym, l1 < 0 ? 0 : l1);
            const int s2 = __shfl_sync(0xffffffffu, sym, l2 < 0 ? 0 : l2);
            const int prev1 = (l1 < 0) ? carry1 : s1;
            const int prev2 = (l1 < 0) ? carry2 : ((l2 < 0) ? carry1 : s2);
            if (active) {
                if (ocq) {
                    row[qp] = (uint8_t)(HS_CODE0 + 5 * prev2 + prev1 + 25 * sym);  // :238,287
                    alen++;
                    if (oty == OP_D) dist++;
                    else if (sym != hs_base2(cb, qp)) dist++;  // :254-256
                } else if (oty == OP_I) {
                    dist++;  // :337-338
                    alen++;
                }
            }
            if (pm) {  // warp-uniform
                const int last = 31 - __clz(pm);
                const unsigned rest = pm & ~(1u << last);
                const int sl = __shfl_sync(0xffffffffu, sym, last);
                const int sr = __shfl_sync(0xffffffffu, sym, rest ? 31 - __clz(rest) : 0);
                carry2 = rest ? sr : carry1;
                carry1 = sl;
            }
        }
        q += __shfl_sync(0xffffffffu, qi, 31);
        t += __shfl_sync(0xffffffffu, ti, 31);
        __syncwarp();
    }
    long long d64 = hs_warp_sum64((long long)dist), a64 = hs_warp_sum64((long long)alen);
    if (lane == 0) {
        if (d64) atomicAdd(a.stats + 3 * c + 0, (unsigned long long)d64);
        if (a64) atomicAdd(a.stats + 3 * c + 1, (unsigned long long)a64);
        if (end > start) atomicAdd(a.stats + 3 * c + 2, (unsigned long long)(end - start));
    }
}

// ---- tile index: for every 128-column tile, the reads overlapping it in ascending order ----------
// One warp per tile scans the reads of the tile's contig (a few thousand) with an ordered
// ballot-compaction; pass 0 counts, pass 1 fills. No atomics, no sort, deterministic.
template <bool FILL>
__global__ void __launch_bounds__(256) tile_index_kernel(int64_t n_tiles, const int32_t* __restrict__ tile_contig,
                                                         const int64_t* __restrict__ tile_base,
                                                         const int64_t* __restrict__ contig_read_off,
                                                         const int32_t* __restrict__ read_start,
                                                         const int32_t* __restrict__ read_end,
                                                         int64_t* __restrict__ tile_cnt_or_off,
                                                         int32_t* __restrict__ tile_reads) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int c = tile_contig[tile];
    const int q0 = (int)(tile - tile_base[c]) * HS_TILE, q1 = q0 + HS_TILE;
    const int64_t r0 = contig_read_off[c], r1 = contig_read_off[c + 1];
    int64_t out = FILL ? tile_cnt_or_off[tile] : 0;
    for (int64_t rb = r0; rb < r1; rb += 32) {
        const int64_t r = rb + lane;
        bool hit = false;
        if (r < r1) {
            const int s = __ldg(read_start + r), e = __ldg(read_end + r);
            hit = (e > s) && (s < q1) && (e > q0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (FILL && hit) tile_reads[out + __popc(m & ((1u << lane) - 1u))] = (int32_t)r;
        out += __popc(m);
    }
    if (!FILL && lane == 0) tile_cnt_or_off[tile] = out;
}

// ---- host side --------------------------------------------------------------------------------
extern "C" {

void hsgpu_pileup_destroy(hsgpu_pileup* p) {
    if (!p) return;
    hsgpu_ctx* ctx = p->ctx;
    cudaSetDevice(ctx->device);
    hs_free(ctx, p->d_contig_len);
    hs_free(ctx, p->d_contig_bases);
    hs_free(ctx, p->d_contig_word_off);
    hs_free(ctx, p->d_contig_read_off);
    hs_free(ctx, p->d_col_base);
    hs_free(ctx, p->d_tile_base);
    hs_free(ctx, p->d_tile_contig);
    hs_free(ctx, p->d_read_contig);
    hs_free(ctx, p->d_read_bases);
    hs_free(ctx, p->d_read_word_off);
    hs_free(ctx, p->d_read_len);
    hs_free(ctx, p->d_cigar);
    hs_free(ctx, p->d_cigar_off);
    hs_free(ctx, p->d_read_start);
    hs_free(ctx, p->d_read_strand);
    hs_free(ctx, p->d_read_end);
    hs_free(ctx, p->d_row_alloc);
    hs_free(ctx, p->d_row_base);
    hs_free(ctx, p->d_codes);
    hs_free(ctx, p->d_stats);
    hs_free(ctx, p->d_tile_off);
    hs_free(ctx, p->d_tile_reads);
    hs_free(ctx, p->d_k0);
    hs_free(ctx, p->d_k1);
    hs_free(ctx, p->d_flags);
    hs_free(ctx, p->d_counts);
    hs_free(ctx, p->d_depth);
    hs_free(ctx, p->d_min_reads);
    hs_free(ctx, p->d_suspect_pos);
    hs_free(ctx, p->d_suspect_auto);
    hs_free(ctx, p->d_n_suspects);
    hs_free(ctx, p->d_depth_sum);
    hs_free(ctx, p->d_suspect_base);
    hs_free(ctx, p->d_col_off);
    hs_free(ctx, p->d_tile_sus);
    hs_free(ctx, p->d_work);
    delete p;
}

int hsgpu_pileup_create(hsgpu_ctx* ctx, const hsgpu_pileup_input* in, hsgpu_pileup** out) {
    if (!ctx || !in || !out) return HSGPU_ERR_ARG;
    *out = nullptr;
    if (in->n_contigs <= 0 || in->n_reads < 0) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: empty batch");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int32_t nc = in->n_contigs;
    const int64_t nr = in->n_reads;
    if (in->contig_read_off[0] != 0 || in->contig_read_off[nc] != nr)
        HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: contig_read_off must span [0, n_reads]");
    hsgpu_pileup* p = new hsgpu_pileup();
    p->ctx = ctx;
    p->n_contigs = nc;
    p->n_reads = nr;
    p->h_contig_len.assign(in->contig_len, in->contig_len + nc);
    p->h_contig_read_off.assign(in->contig_read_off, in->contig_read_off + nc + 1);
    p->h_col_base.resize(nc + 1);
    p->h_tile_base.resize(nc + 1);
    p->h_suspect_base.resize(nc + 1);
    p->h_stats.assign((size_t)3 * nc, 0);
    int64_t cols = 0, tiles = 0, sus = 0;
    for (int c = 0; c < nc; c++) {
        if (in->contig_len[c] < 0 || in->contig_read_off[c + 1] < in->contig_read_off[c]) {
            delete p;
            HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: negative contig length or read range");
        }
        p->h_col_base[c] = cols;
        p->h_tile_base[c] = tiles;
        p->h_suspect_base[c] = sus;
        cols += in->contig_len[c];
        tiles += (in->contig_len[c] + HS_TILE - 1) / HS_TILE;
        sus += in->contig_len[c] / 6 + 2;  // suspects are > 5 columns apart (:529)
    }
    p->h_col_base[nc] = cols;
    p->h_tile_base[nc] = tiles;
    p->h_suspect_base[nc] = sus;
    p->n_cols = cols;
    p->n_tiles = tiles;
    p->n_cigar = in->cigar_off[nr];
    for (int64_t r = 0; r < nr; r++) {
        if (in->read_start[r] < 0 || in->read_len[r] < 0) {
            delete p;
            HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pileup_create: negative read start or length");
        }
    }
    std::vector<int32_t> read_contig((size_t)nr);
    for (int c = 0; c < nc; c++)
        for (int64_t r = in->contig_read_off[c]; r < in->contig_read_off[c + 1]; r++) read_contig[r] = c;
    std::vector<int32_t> tile_contig((size_t)tiles);
    for (int c = 0; c < nc; c++)
        for (int64_t t = p->h_tile_base[c]; t < p->h_tile_base[c + 1]; t++) tile_contig[t] = c;

    const int64_t contig_words = in->contig_word_off[nc];
    const int64_t read_words = in->read_word_off[nr];
#define A(ptr, n) HS_CUDA(ctx, hs_alloc(ctx, &p->ptr, (n)))
    A(d_contig_len, nc); A(d_contig_bases, contig_words); A(d_contig_word_off, nc + 1);
    A(d_contig_read_off, nc + 1); A(d_col_base, nc + 1); A(d_tile_base, nc + 1); A(d_tile_contig, tiles);
    A(d_read_contig, nr); A(d_read_bases, read_words); A(d_read_word_off, nr + 1); A(d_read_len, nr);
    A(d_cigar, p->n_cigar); A(d_cigar_off, nr + 1); A(d_read_start, nr); A(d_read_strand, nr);
    A(d_read_end, nr); A(d_row_alloc, nr + 1); A(d_row_base, nr); A(d_stats, 3 * nc);
    A(d_tile_off, tiles + 1); A(d_suspect_base, nc + 1);
#undef A
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_len, in->contig_len, nc));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_bases, in->contig_bases, contig_words));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_word_off, in->contig_word_off, nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_contig_read_off, in->contig_read_off, nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_col_base, p->h_col_base.data(), nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_tile_base, p->h_tile_base.data(), nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_suspect_base, p->h_suspect_base.data(), nc + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_tile_contig, tile_contig.data(), tiles));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_contig, read_contig.data(), nr));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_bases, in->read_bases, read_words));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_word_off, in->read_word_off, nr + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_len, in->read_len, nr));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_cigar, in->cigar, p->n_cigar));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_cigar_off, in->cigar_off, nr + 1));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_start, in->read_start, nr));
    HS_CUDA(ctx, hs_h2d(ctx, p->d_read_strand, in->read_strand, nr));
    // the local vectors must outlive the async copies from pageable memory
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = p;
    return HSGPU_OK;
}

int hsgpu_pileup_build(hsgpu_pileup* p) {
    if (!p) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nr = p->n_reads;
    const unsigned rblocks = (unsigned)((nr + 7) / 8);
    // rebuilt from scratch on every call (bench steps call this repeatedly on resident inputs)
    hs_free(ctx, p->d_codes);
    hs_free(ctx, p->d_tile_reads);
    p->built = p->ranked = p->have_col_off = false;
    HS_CUDA(ctx, cudaMemsetAsync(p->d_stats, 0, sizeof(unsigned long long) * 3 * p->n_contigs, ctx->stream));
    int64_t* d_totals = nullptr;
    HS_CUDA(ctx, hs_alloc(ctx, &d_totals, 2));
    if (nr > 0) {
        HS_KERNEL(ctx, "span_kernel", span_kernel<<<rblocks, 256, 0, ctx->stream>>>(nr, p->d_cigar, p->d_cigar_off, p->d_read_start, p->d_read_contig,
                                                      p->d_contig_len, p->d_read_end, p->d_row_alloc));
    }
    int rc = hs_exclusive_scan_i64(ctx, p->d_row_alloc, p->d_row_alloc, nr, d_totals);
    if (rc) return rc;
    if (p->n_tiles > 0) {
        HS_KERNEL(ctx, "tile_index_kernel<false>", tile_index_kernel<false><<<(unsigned)((p->n_tiles + 7) / 8), 256, 0, ctx->stream>>>(
            p->n_tiles, p->d_tile_contig, p->d_tile_base, p->d_contig_read_off, p->d_read_start, p->d_read_end,
            p->d_tile_off, nullptr));
    }
    rc = hs_exclusive_scan_i64(ctx, p->d_tile_off, p->d_tile_off, p->n_tiles, d_totals + 1);
    if (rc) return rc;
    int64_t totals[2] = {0, 0};
    HS_CUDA(ctx, hs_d2h(ctx, totals, d_totals, 2));
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the only host round trip of the build: two sizes
    hs_free(ctx, d_totals);
    p->codes_bytes = totals[0];
    p->tile_entries = totals[1];
    HS_CUDA(ctx, hs_alloc(ctx, &p->d_codes, p->codes_bytes + HS_ALIGN));
    HS_CUDA(ctx, hs_alloc(ctx, &p->d_tile_reads, p->tile_entries));
    HS_CUDA(ctx, cudaMemcpyAsync(p->d_tile_off + p->n_tiles, &p->tile_entries, sizeof(int64_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
    if (nr > 0) {
        PileupArgs a;
        a.n_reads = nr;
        a.contig_len = p->d_contig_len;
        a.contig_bases = p->d_contig_bases;
        a.contig_word_off = p->d_contig_word_off;
        a.read_contig = p->d_read_contig;
        a.read_bases = p->d_read_bases;
        a.read_word_off = p->d_read_word_off;
        a.read_len = p->d_read_len;
        a.cigar = p->d_cigar;
        a.cigar_off = p->d_cigar_off;
        a.read_start = p->d_read_start;
        a.read_strand = p->d_read_strand;
        a.read_end = p->d_read_end;
        a.row_off = p->d_row_alloc;
        a.row_base = p->d_row_base;
        a.codes = p->d_codes;
        a.stats = p->d_stats;
        HS_KERNEL(ctx, "pileup_kernel", pileup_kernel<<<rblocks, 256, 0, ctx->stream>>>(a));
    }
    if (p->n_tiles > 0) {
        HS_KERNEL(ctx, "tile_index_kernel<true>", tile_index_kernel<true><<<(unsigned)((p->n_tiles + 7) / 8), 256, 0, ctx->stream>>>(
            p->n_tiles, p->d_tile_contig, p->d_tile_base, p->d_contig_read_off, p->d_read_start, p->d_read_end,
            p->d_tile_off, p->d_tile_reads));
    }
    p->built = true;
    return HSGPU_OK;
}

int hsgpu_pileup_stats(hsgpu_pileup* p, int64_t* n_cells, int64_t* distance_sum, int64_t* aligned_sum) {
    if (!p) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pileup_stats: call hsgpu_pileup_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "");
    HS_CUDA(ctx, hs_d2h(ctx, (unsigned long long*)p->h_stats.data(), p->d_stats, 3 * p->n_contigs));
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < p->n_contigs; c++) {
        if (distance_sum) distance_sum[c] = p->h_stats[3 * c + 0];
        if (aligned_sum) aligned_sum[c] = p->h_stats[3 * c + 1];
        if (n_cells) n_cells[c] = p->h_stats[3 * c + 2];
    }
    return HSGPU_OK;
}

float hsgpu_mean_distance(int64_t distance_sum, int64_t aligned_sum) {
    // `float totalDistance` grows by 1.0f per event and therefore saturates at 2^24; the length is a
    // double that starts at 1 (src/call_variants.cpp:67-68); the quotient is returned as float (:434)
    float total_distance = (float)(distance_sum > 16777216 ? 16777216 : distance_sum);
    double total_length = 1.0 + (double)aligned_sum;
    return (float)(total_distance / total_length);
}

int hsgpu_pileup_read_ends(hsgpu_pileup* p, int32_t* read_end) {
    if (!p || !read_end) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = p->ctx;
    if (!p->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pileup_read_ends: call hsgpu_pileup_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    HS_CUDA(ctx, hs_d2h(ctx, read_end, p->d_read_end, p->n_reads));
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return HSGPU_OK;
}

}  // extern "C"
