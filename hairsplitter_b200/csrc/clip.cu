// Read clipping and window extraction of modify_GFA (reference src/create_new_contigs.cpp:383-447): for every
// (read, contig interval) pair, the part of the read and of its CIGAR that lies on [leftToPolish, rightToPolish] --
// what is handed to the polisher for each window and cluster. The reference walks the expanded CIGAR of the read
// one character at a time from its first character, once per interval the read takes part in.
//
// Here a warp takes one (read, interval) item and works on the run-length ops: three prefix sums per 32 ops give
// every op its expanded index, read cursor and interval cursor; the walk's three events -- the first character at
// or past the left end, the first character at the right end, the first clip character after the start -- each
// have a closed form inside an op, so they are minima over the ops. Same family as the pileup's CIGAR walk
// (pileup.cu), but on the BAM letters themselves: the reference's loop moves its cursors for 'M', 'D', 'I' and
// clips only, while '=', 'X', 'N', 'P' are looked at and move nothing (:426-435), which the byte-sized CIGAR of the
// pileup does not distinguish.
#include "common.cuh"

struct ClipArgs {
    int64_t n_items;
    const uint32_t* cigar;
    const int64_t* cigar_off;
    const int32_t* pos_2_1;
    const int64_t* item_read;
    const int32_t* left;
    const int32_t* right;
    hsgpu_clip* out;
};

struct ClipEvent {  // a character of the expanded CIGAR: its index, the op it belongs to, its offset in the op, the read cursor
    long long idx;
    int op, off, rpos;
};

// keeps the earlier of `best` and the warp's earliest candidate (cand.idx < 0: this lane has none)
__device__ __forceinline__ void clip_take_min(ClipEvent& best, const ClipEvent& cand) {
    long long v = cand.idx < 0 ? 0x7fffffffffffffffll : cand.idx;
    long long m = v;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (m == 0x7fffffffffffffffll) return;
    const unsigned who = __ballot_sync(0xffffffffu, v == m);
    const int src = __ffs(who) - 1;
    ClipEvent e;
    e.idx = m;
    e.op = __shfl_sync(0xffffffffu, cand.op, src);
    e.off = __shfl_sync(0xffffffffu, cand.off, src);
    e.rpos = __shfl_sync(0xffffffffu, cand.rpos, src);
    if (best.idx < 0 || e.idx < best.idx) best = e;
}

// BAM op letters: M I D N S H P = X
__device__ __forceinline__ bool clip_is_clip(int ty) { return ty == 4 || ty == 5; }
__device__ __forceinline__ bool clip_adv_read(int ty) { return ty == 0 || ty == 1 || ty == 4 || ty == 5; }  // M, I and (until the start) S/H
__device__ __forceinline__ bool clip_adv_interval(int ty) { return ty == 0 || ty == 2; }                    // M, D

__global__ void __launch_bounds__(256) clip_kernel(ClipArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= a.n_items) return;  // whole warps leave together
    const int64_t r = a.item_read[item];
    const int64_t k0 = a.cigar_off[r], k1 = a.cigar_off[r + 1];
    const int left = a.left[item], right = a.right[item];
    ClipEvent start, end_eq, end_clip;
    start.idx = end_eq.idx = end_clip.idx = -1;
    start.op = start.off = start.rpos = 0;
    end_eq = start;
    end_clip = start;
    long long tot_len = 0;
    int tot_read = 0;
    for (int pass = 0; pass < 2; pass++) {
        // pass 0: the start and the end-by-position events; pass 1 (needs the start): the first clip after the start
        long long base_idx = 0;
        int base_r = 0, base_i = a.pos_2_1[r];
        for (int64_t kb = k0; kb < k1; kb += 32) {
            const int64_t k = kb + lane;
            const uint32_t op = k < k1 ? __ldg(a.cigar + k) : 0u;
            const int n = (int)(op >> 4), ty = (int)(op & 15u);
            const int ar = clip_adv_read(ty) ? n : 0, ai = clip_adv_interval(ty) ? n : 0;
            const int in = hs_warp_incl_scan(n, lane), ir = hs_warp_incl_scan(ar, lane), ii = hs_warp_incl_scan(ai, lane);
            const long long idx0 = base_idx + in - n;  // state in front of the op's first character
            const int r0 = base_r + ir - ar, i0 = base_i + ii - ai;
            ClipEvent c;
            c.idx = -1;
            c.op = (int)(k - k0);
            c.off = 0;
            c.rpos = 0;
            if (pass == 0) {
                if (n > 0 && !clip_is_clip(ty)) {
                    // first character with posOnInterval >= leftToPolish (:417-421)
                    int j = -1;
                    if (i0 >= left) j = 0;
                    else if (ai && left - i0 < n) j = left - i0;
                    if (j >= 0) {
                        c.idx = idx0 + j;
                        c.off = j;
                        c.rpos = r0 + (ar ? j : 0);
                    }
                }
                clip_take_min(start, c);
                c.idx = -1;
                if (n > 0 && !clip_is_clip(ty)) {
                    // first character with posOnInterval == rightToPolish (:422-426)
                    int j = -1;
                    if (ai) {
                        if (right >= i0 && right - i0 < n) j = right - i0;
                    } else if (i0 == right) {
                        j = 0;
                    }
                    if (j >= 0) {
                        c.idx = idx0 + j;
                        c.off = j;
                        c.rpos = r0 + (ar ? j : 0);
                    }
                }
                clip_take_min(end_eq, c);
            } else {
                // the first 'S' / 'H' character once the start has been found ends the walk (:408-413)
                if (n > 0 && clip_is_clip(ty) && idx0 > start.idx) {
                    c.idx = idx0;
                    c.rpos = r0;
                }
                clip_take_min(end_clip, c);
            }
            base_idx += __shfl_sync(0xffffffffu, in, 31);
            base_r += __shfl_sync(0xffffffffu, ir, 31);
            base_i += __shfl_sync(0xffffffffu, ii, 31);
        }
        tot_len = base_idx;
        tot_read = base_r;
        if (pass == 0 && start.idx < 0) break;  // posOnReadStart stays -1
    }
    if (lane != 0) return;
    hsgpu_clip o;
    o.status = 0;
    o.read_start = o.read_end = o.cigar_start = o.cigar_end = -1;
    o.op_first = o.op_first_skip = 0;
    o.op_last = -1;
    o.op_last_take = 0;
    if (start.idx < 0 || (end_eq.idx >= 0 && end_eq.idx < start.idx)) {
        o.status = -2;  // the start was never set (the walk ended first): "can happen when within a deletion" (:442-447)
    } else {
        ClipEvent e;
        e.idx = -1;
        if (end_eq.idx >= 0) e = end_eq;
        if (end_clip.idx >= 0 && (e.idx < 0 || end_clip.idx < e.idx)) e = end_clip;
        o.read_start = start.rpos;
        o.cigar_start = (int32_t)start.idx;
        o.op_first = start.op;
        o.op_first_skip = start.off;
        if (e.idx >= 0) {
            o.read_end = e.rpos;
            o.cigar_end = (int32_t)e.idx;
            o.op_last = e.op;
            o.op_last_take = e.off;
        } else {  // the walk ran to the end of the CIGAR (:437-440)
            o.read_end = tot_read;
            o.cigar_end = (int32_t)tot_len;
            o.op_last = (int32_t)(k1 - k0);
            o.op_last_take = 0;
        }
        if (o.read_start > o.read_end) o.status = -2;
    }
    a.out[item] = o;
}

extern "C" int hsgpu_clip_reads(hsgpu_ctx* ctx, int64_t n_reads, const uint32_t* cigar, const int64_t* cigar_off,
                                const int32_t* pos_2_1, int64_t n_items, const int64_t* item_read,
                                const int32_t* left_to_polish, const int32_t* right_to_polish, hsgpu_clip* out) {
    if (!ctx || n_reads < 0 || n_items < 0) return HSGPU_ERR_ARG;
    if (n_items == 0) return HSGPU_OK;
    if (!cigar_off || !pos_2_1 || !item_read || !left_to_polish || !right_to_polish || !out) return HSGPU_ERR_ARG;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n_ops = cigar_off[n_reads];
    if (n_ops > 0 && !cigar) return HSGPU_ERR_ARG;
    for (int64_t i = 0; i < n_items; i++)
        if (item_read[i] < 0 || item_read[i] >= n_reads) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_clip_reads: item_read out of range");
    for (int64_t r = 0; r < n_reads; r++)
        if (cigar_off[r + 1] < cigar_off[r]) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_clip_reads: cigar_off is not ascending");
    void* block = nullptr;
    ClipArgs a;
    uint32_t* d_cigar;
    int64_t *d_off, *d_item;
    int32_t *d_pos, *d_left, *d_right;
    hsgpu_clip* d_out;
    HsCarve cv;
    cv.add(&d_cigar, n_ops);
    cv.add(&d_off, n_reads + 1);
    cv.add(&d_pos, n_reads);
    cv.add(&d_item, n_items);
    cv.add(&d_left, n_items);
    cv.add(&d_right, n_items);
    cv.add(&d_out, n_items);
    HS_CUDA(ctx, cv.alloc(ctx, &block));
    cudaError_t e = hs_h2d(ctx, d_cigar, cigar, n_ops);
    if (e == cudaSuccess) e = hs_h2d(ctx, d_off, cigar_off, n_reads + 1);
    if (e == cudaSuccess) e = hs_h2d(ctx, d_pos, pos_2_1, n_reads);
    if (e == cudaSuccess) e = hs_h2d(ctx, d_item, item_read, n_items);
    if (e == cudaSuccess) e = hs_h2d(ctx, d_left, left_to_polish, n_items);
    if (e == cudaSuccess) e = hs_h2d(ctx, d_right, right_to_polish, n_items);
    if (e != cudaSuccess) {
        hs_free(ctx, block);
        return hs_cuda_fail(ctx, e, "hsgpu_clip_reads: upload", __FILE__, __LINE__);
    }
    a.n_items = n_items;
    a.cigar = d_cigar;
    a.cigar_off = d_off;
    a.pos_2_1 = d_pos;
    a.item_read = d_item;
    a.left = d_left;
    a.right = d_right;
    a.out = d_out;
    if (ctx->profiling) hs_prof_begin(ctx, "clip_kernel");
    clip_kernel<<<(unsigned)((n_items + 7) / 8), 256, 0, ctx->stream>>>(a);
    if (ctx->profiling) hs_prof_end(ctx);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = hs_d2h(ctx, out, d_out, n_items);
    if (e == cudaSuccess) e = hs_stream_sync(ctx);
    hs_free(ctx, block);
    if (e != cudaSuccess) return hs_cuda_fail(ctx, e, "hsgpu_clip_reads", __FILE__, __LINE__);
    return HSGPU_OK;
}
