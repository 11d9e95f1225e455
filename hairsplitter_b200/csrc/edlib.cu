// Batched realignment with edlib's semantics: edlibAlign (reference src/edlib/src/edlib.cpp:142-297,
// header src/edlib/include/edlib.h:146-271) for a batch of (query, target) pairs.
//
// One warp per pair. The query is cut into 64-row blocks, one block per lane (Myers/Hyyro bit-vectors,
// calculateBlock :411-446); the lanes run the columns of the DP matrix as a skewed wavefront -- at step
// s lane b computes column s-b -- so that the horizontal delta leaving block b-1 reaches block b through
// one warp shuffle together with the target symbol. edlib's Ukkonen band only prunes work, it does not
// change results, so the kernel computes the full matrix and reproduces edlib's observable rules
// (restated and pinned against the vendored edlib in oracle/hs_oracle_edlib.c):
//   * HW/SHW: every target position whose bottom-row score equals the minimum, ascending; position -1
//     (the empty prefix) competes only when the query length is not a multiple of 64, because edlib pads
//     the query with W wildcard rows and reads column c-W of the padded bottom row (:665-702);
//   * HW start locations: reverse SHW pass per end location, LAST minimal position (:226-259);
//   * path: NW on target[start0..end0], traceback priority up (insert) > left (delete) > diagonal
//     (obtainAlignmentTraceback :947-1146). The traceback only needs "vertical delta is +1" and
//     "horizontal delta is +1" per cell, so the NW pass stores the two bit-vectors Pv and Ph per
//     (block, column), 2 bits per cell, in a per-warp scratch slab laid out [block][column] so that the
//     walk loads 32 consecutive columns of the block it is in with one coalesced request.
//   * paths at or above edlib's 1 MiB switch ((20*blocks+8)*columns >= 1 MiB, :1193-1195) follow
//     obtainAlignmentHirschberg (:1236-1401): the target is halved, the last score column of the left half and of the
//     reversed right half are rebuilt from the final Pv/Mv bit-vectors of each block, the first query row (ascending)
//     where the two add up to the distance is the split, both parts go back on a per-warp stack (right part first,
//     because the ops are produced back to front) until a part passes the 1 MiB test and is traced back as above.
// Queries longer than 2048 (more than one block per lane) are swept in strips of 32 blocks: the horizontal deltas that
// leave the bottom row of a strip are kept per column (one byte) and enter the next strip as its top row.
// The batch kernels come in two instantiations: the ordinary one (queries up to 2048, paths below the switch; the
// throughput path), and the LONG one that runs over a list of the remaining pairs -- launched only when there are any.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

#define ED_WARPS 4          // warps (pairs in flight) per CTA
#define ED_MAXBLOCKS 32     // one 64-row block per lane -> queries up to 2048 in one strip
#define ED_MAX_QUERY (1 << 20)
#define ED_STACK 40         // Hirschberg parts waiting per warp (the target halves at every level)
#define ED_PACK_MAXBLOCKS 32  // queries up to 2048 share a warp with others of the same block count
#define ED_PACK_MAXBPL 4      // blocks of 64 rows stacked in one lane
#define ED_PACK_MAXLOC 15     // ... as long as they have few end locations (the start-location sweeps run in step;
                              // phase B groups pairs of equal location count)
enum { ED_ROUTE_ORDINARY = 0, ED_ROUTE_PACKED = 1, ED_ROUTE_LONG = 2, ED_ROUTE_NONE = 3 };
#define ED_SMEM_SYMS 8      // alphabets up to this size keep Peq in shared memory
#define ED_TRACE_BYTES (52429ll * 16)  // per-warp traceback slab: blocks*columns < 2^20/20 entries of 16 B

struct EdArgs {
    int n_pairs;
    const uint8_t* q;
    const int64_t* q_off;
    const uint8_t* t;
    const int64_t* t_off;
    int k, mode, task;
    hsgpu_edlib_result* res;
    unsigned int* bitmask;        // end-position bits per pair
    const int64_t* bm_off;        // [n_pairs+1] word offsets
    int32_t* ends;
    int32_t* starts;
    uint8_t* aln_tmp;             // per pair region of qlen + tlen bytes (ops in reverse order)
    const int64_t* aln_tmp_off;
    uint64_t* peq_big;            // per resident warp: 256 * 32 words, for alphabets > ED_SMEM_SYMS
    uint8_t* trace;               // per resident warp: ED_TRACE_BYTES
    unsigned int* counter;        // dynamic pair scheduler
    const uint8_t* route;         // per pair: which launch works on it (ED_ROUTE_*)
    // packed launches only: several short queries share a warp
    const int32_t* alpha_len;     // per pair alphabet size (edlib_alphabet_kernel)
    const unsigned int* batch_alpha;  // 256-bit set of the bytes present anywhere in the batch
    const int32_t* plist;         // pairs in task order
    const int32_t* tasks;         // per task: first index into plist, (blocks << 8) | pairs
    int n_tasks;
    uint8_t* ptrace;              // per resident warp of the packed phase B: ptrace_stride bytes shared by its groups
    int64_t ptrace_stride;
    // LONG instantiations only
    const int32_t* list;          // pairs to work on
    int n_list;
    uint8_t* hbuf;                // per warp: 2 * hbuf_stride bytes (horizontal deltas between strips, ping-pong)
    int64_t hbuf_stride;
    ulonglong2* colv;             // per warp: 2 * col_stride final (Pv, Mv) per block: left half, reversed right half
    int* colpre;                  // per warp: 2 * (col_stride + 1) prefix sums of the blocks' vertical deltas
    int64_t col_stride;
};

struct WarpCtx {
    uint64_t* peq;        // [sym][32]
    const uint8_t* lut;   // byte -> symbol
    int lane;
    // strip mode (MULTI passes build Peq themselves, 32 blocks at a time)
    const uint8_t* q;
    bool qrev;
    int n_sym;
    uint8_t* hbuf;
    int64_t hbuf_stride;
};

// symbol table of a pair: lut[byte] = rank of the byte among the bytes present (any consistent numbering
// gives the same Peq behaviour; edlib numbers by first appearance, :1440-1459). Returns alphabet size.
__device__ int build_alphabet(const uint8_t* q, int m, const uint8_t* t, int n, uint8_t* lut, int lane) {
    unsigned int present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < m; i += 32) { const int c = q[i]; present[c >> 5] |= 1u << (c & 31); }
    for (int i = lane; i < n; i += 32) { const int c = t[i]; present[c >> 5] |= 1u << (c & 31); }
    int total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        unsigned int v = present[w];
        for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        present[w] = v;
    }
    int base = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const unsigned int v = present[w];
        lut[w * 32 + lane] = (uint8_t)(base + __popc(v & ((1u << lane) - 1u)));
        base += __popc(v);
    }
    total = base;
    __syncwarp();
    return total;
}

// Peq[sym][block]: bit i of block b set iff query row 64*b+i carries sym (buildPeq :357-385; padding rows
// stay 0 -- they never influence the real rows because carries only travel towards higher rows)
__device__ void build_peq(const WarpCtx& w, const uint8_t* q, int m, int nb, int n_sym, bool reversed, int row0 = 0) {
    for (int i = w.lane; i < n_sym * 32; i += 32) w.peq[i] = 0ull;
    __syncwarp();
    if (w.lane < nb) {
        const int r0 = row0 + w.lane * 64;
        const int r1 = min(m, r0 + 64);
        for (int r = r0; r < r1; r++) {
            const int c = reversed ? q[m - 1 - r] : q[r];
            w.peq[w.lut[c] * 32 + w.lane] |= 1ull << (r - r0);
        }
    }
    __syncwarp();
}

enum { PASS_SEMIGLOBAL = 0, PASS_NW_SCORE = 1, PASS_REV_SHW = 2, PASS_NW_STORE = 3, PASS_NW_COLUMN = 4 };

struct PassOut {
    int best;       // minimum bottom-row score (semi-global) / D[m][n] (NW)
    int first_j;    // first column index j (0..n) reaching best
    int last_j;     // last column index reaching best
    int count;      // number of columns with score == best, j >= first_j
};

// One sweep over the columns. tget(c) = target symbol index source: forward t[c], or t[rev_end - c].
// start_hin: +1 when the top row is penalised (NW, SHW), 0 for HW. consider_j0: position -1 competes.
// MULTI: any query length, in strips of 32 blocks (w.q / w.qrev / w.n_sym / w.hbuf set by the caller);
// otherwise the caller has built Peq for the at most 32 blocks. PASS_NW_COLUMN leaves the vertical deltas of the last
// column in colv[block] = (Pv, Mv).
template <int KIND, bool MULTI>
__device__ PassOut dp_pass(const WarpCtx& w, int m, int nb_total, const uint8_t* t, int n, int rev_end, int start_hin,
                           bool consider_j0, unsigned int* bitmask, ulonglong2* trace, ulonglong2* colv = nullptr) {
    const int lane = w.lane;
    const int lb = (m - 1) & 63;
    int score = m;  // D[m][0]
    PassOut o;
    o.best = consider_j0 ? m : 0x3fffffff;
    o.first_j = 0;
    o.last_j = 0;
    o.count = consider_j0 ? 1 : 0;
    unsigned int bits = consider_j0 ? 1u : 0u;  // bit (j & 31) of the word being assembled
    const int n_strips = MULTI ? (nb_total + 31) >> 5 : 1;
    for (int st = 0; st < n_strips; st++) {
        const int nb = MULTI ? min(32, nb_total - st * 32) : nb_total;
        const bool last = !MULTI || st == n_strips - 1;
        const uint8_t* hsrc = nullptr;
        uint8_t* hdst = nullptr;
        if (MULTI) {
            build_peq(w, w.q, m, nb, w.n_sym, w.qrev, st * 64 * 32);
            hsrc = w.hbuf + (size_t)(st & 1) * w.hbuf_stride;
            hdst = w.hbuf + (size_t)((st + 1) & 1) * w.hbuf_stride;
        }
        uint64_t Pv = ~0ull, Mv = 0ull;
        int packed = 0;
        int tchunk = 0, hchunk = 0;
        const int steps = n + nb - 1;
        for (int s = 0; s < steps; s++) {
            if ((s & 31) == 0) {
                const int idx = s + lane;
                tchunk = 0;
                if (idx < n) tchunk = w.lut[rev_end >= 0 ? t[rev_end - idx] : t[idx]];
                if (MULTI && st > 0) hchunk = idx < n ? hsrc[idx] : 1;
            }
            const int sym0 = __shfl_sync(0xffffffffu, tchunk, s & 31);
            int hin0 = start_hin;
            if (MULTI && st > 0) hin0 = __shfl_sync(0xffffffffu, hchunk, s & 31) - 1;
            const int pk = __shfl_up_sync(0xffffffffu, packed, 1);
            const int sym = lane == 0 ? sym0 : (pk >> 2);
            const int hin = lane == 0 ? hin0 : ((pk & 3) - 1);
            const int c = s - lane;
            if (lane < nb && c >= 0 && c < n) {
                uint64_t Eq = w.peq[sym * 32 + lane];
                const uint64_t hneg = hin < 0 ? 1ull : 0ull, hpos = hin > 0 ? 1ull : 0ull;
                const uint64_t Xv = Eq | Mv;
                Eq |= hneg;
                const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                uint64_t Ph = Mv | ~(Xh | Pv);
                uint64_t Mh = Pv & Xh;
                const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
                packed = (sym << 2) | (hout + 1);
                if (lane == nb - 1) {
                    if (!last) {
                        hdst[c] = (uint8_t)(hout + 1);
                    } else {
                        score += (int)((Ph >> lb) & 1ull) - (int)((Mh >> lb) & 1ull);
                        const int j = c + 1;
                        if (KIND == PASS_SEMIGLOBAL) {
                            if ((j & 31) == 0) bits = 0;
                            if (score < o.best) { o.best = score; o.first_j = j; o.count = 0; }
                            if (score == o.best) { o.count++; o.last_j = j; bits |= 1u << (j & 31); }
                            if ((j & 31) == 31 || j == n) bitmask[j >> 5] = bits;
                        } else if (KIND == PASS_REV_SHW) {
                            if (score < o.best) { o.best = score; o.first_j = j; }
                            if (score == o.best) o.last_j = j;
                        } else {
                            o.best = score;  // NW: value after the last column is D[m][n]
                        }
                    }
                }
                const uint64_t Phs = (Ph << 1) | hpos;
                const uint64_t Mhs = (Mh << 1) | hneg;
                Pv = Mhs | ~(Xv | Phs);
                Mv = Phs & Xv;
                if (KIND == PASS_NW_STORE) trace[(size_t)(st * 32 + lane) * n + c] = make_ulonglong2(Pv, Ph);
            }
        }
        if (KIND == PASS_NW_COLUMN && lane < nb) colv[st * 32 + lane] = make_ulonglong2(Pv, Mv);
        if (MULTI) __syncwarp();
    }
    // results live in the lane of the last block
    const int rl = (nb_total - 1) & 31;
    o.best = __shfl_sync(0xffffffffu, o.best, rl);
    o.first_j = __shfl_sync(0xffffffffu, o.first_j, rl);
    o.last_j = __shfl_sync(0xffffffffu, o.last_j, rl);
    o.count = __shfl_sync(0xffffffffu, o.count, rl);
    return o;
}

// ---- phase A: distance + end positions ------------------------------------------------------------
template <bool LONG>
__global__ void __launch_bounds__(ED_WARPS * 32) edlib_phase_a_kernel(EdArgs a) {
    __shared__ uint64_t s_peq[ED_WARPS][ED_SMEM_SYMS * 32];
    __shared__ uint8_t s_lut[ED_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * ED_WARPS + wid;
    WarpCtx w;
    w.lane = lane;
    w.lut = s_lut[wid];
    w.qrev = false;
    w.hbuf = LONG ? a.hbuf + (size_t)gw * 2 * a.hbuf_stride : nullptr;
    w.hbuf_stride = a.hbuf_stride;
    const int n_work = LONG ? a.n_list : a.n_pairs;
    for (;;) {
        int pair = 0;
        if (lane == 0) pair = (int)atomicAdd(a.counter, 1u);
        pair = __shfl_sync(0xffffffffu, pair, 0);
        if (pair >= n_work) break;
        if (LONG) pair = a.list[pair];
        const int m = (int)(a.q_off[pair + 1] - a.q_off[pair]);
        const int n = (int)(a.t_off[pair + 1] - a.t_off[pair]);
        if (!LONG && a.route[pair] != ED_ROUTE_ORDINARY) continue;  // left to the packed / LONG launches
        const uint8_t* q = a.q + a.q_off[pair];
        const uint8_t* t = a.t + a.t_off[pair];
        unsigned int* bm = a.bitmask + a.bm_off[pair];
        __syncwarp();
        const int n_sym = build_alphabet(q, m, t, n, s_lut[wid], lane);
        hsgpu_edlib_result r;
        r.status = 0;
        r.edit_distance = -1;
        r.n_locations = 0;
        r.alignment_length = 0;
        r.alphabet_length = n_sym;
        r.has_start_locations = 0;
        r.loc_off = 0;
        r.aln_off = 0;
        if (m == 0 || n == 0) {  // :162-180
            if (a.mode == 0) r.edit_distance = max(m, n);
            else r.edit_distance = m;
            r.n_locations = 1;
            if (lane == 0) {
                // bit j <-> position j-1: NW reports n-1, SHW/HW report -1
                const int j = a.mode == 0 ? n : 0;
                bm[j >> 5] = 1u << (j & 31);
                a.res[pair] = r;
            }
            continue;
        }
        const int nb = (m + 63) >> 6;
        w.peq = n_sym <= ED_SMEM_SYMS ? s_peq[wid] : a.peq_big + (size_t)gw * 256 * 32;
        w.q = q;
        w.n_sym = n_sym;
        if (!LONG) build_peq(w, q, m, nb, n_sym, false);
        const bool unbounded = a.k < 0;
        if (a.mode == 0) {
            int kk = unbounded ? 0x3fffffff : a.k;
            if (!(kk < abs(n - m))) {  // :741-744
                kk = min(kk, max(m, n));
                const PassOut o = dp_pass<PASS_NW_SCORE, LONG>(w, m, nb, t, n, -1, 1, false, nullptr, nullptr);
                if (o.best <= kk) {
                    r.edit_distance = o.best;
                    r.n_locations = 1;
                    if (lane == 0) bm[n >> 5] = 1u << (n & 31);
                }
            }
        } else {
            const bool j0 = (m & 63) != 0;  // W > 0
            const PassOut o = dp_pass<PASS_SEMIGLOBAL, LONG>(w, m, nb, t, n, -1, a.mode == 2 ? 0 : 1, j0, bm, nullptr);
            int kk = unbounded ? 0x3fffffff : a.k;
            if (a.mode == 2) kk = min(kk, m);  // :565-567
            if (o.best <= kk) {
                r.edit_distance = o.best;
                r.n_locations = o.count;
                // remember where the valid bits start: stale bits of worse minima lie before first_j
                r.loc_off = o.first_j;
            }
        }
        if (lane == 0) a.res[pair] = r;
    }
}

// ---- phase B: locations, start locations, path -------------------------------------------------------
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// warp-uniform traceback from (m, n) over the stored Pv/Ph bit-vectors; ops are written in reverse order
__device__ int traceback(const ulonglong2* trace, const uint8_t* q, int m, const uint8_t* t, int n, uint8_t* out,
                         int lane) {
    int i = m, j = n, len = 0;
    unsigned int pend = 0;  // op of output position len - (len & 31) + lane
    int win_block = -1, win_j0 = -1;
    uint64_t wPv = 0, wPh = 0;
    while (i > 0 && j > 0) {
        const int b = (i - 1) >> 6;
        if (b != win_block || j > win_j0 || j <= win_j0 - 32) {
            win_block = b;
            win_j0 = j;
            const int jj = j - lane;  // lane l holds column j0 - l
            ulonglong2 v = make_ulonglong2(0ull, 0ull);
            if (jj >= 1) v = trace[(size_t)b * n + (jj - 1)];
            wPv = v.x;
            wPh = v.y;
        }
        const int src = win_j0 - j;
        const uint64_t pv = shfl64(wPv, src), ph = shfl64(wPh, src);
        const int bit = (i - 1) & 63;
        int op;
        if ((pv >> bit) & 1ull) { op = 1; i--; }            // up: insertion (:1024-1056)
        else if ((ph >> bit) & 1ull) { op = 2; j--; }       // left: deletion (:1057-1088)
        else { op = (q[i - 1] == t[j - 1]) ? 0 : 3; i--; j--; }  // diagonal (:1089-1135)
        if ((len & 31) == lane) pend = (unsigned)op;
        len++;
        if ((len & 31) == 0) out[len - 32 + lane] = (uint8_t)pend;
    }
    // boundary: the rest of the query is inserted / the rest of the target deleted
    const int rest = i > 0 ? i : j;
    const int rest_op = i > 0 ? 1 : 2;
    for (int r = 0; r < rest; r++) {
        if ((len & 31) == lane) pend = (unsigned)rest_op;
        len++;
        if ((len & 31) == 0) out[len - 32 + lane] = (uint8_t)pend;
    }
    if (lane < (len & 31)) out[(len & ~31) + lane] = (uint8_t)pend;
    return len;
}

// ---- packed form: several pairs per warp, several blocks per lane ------------------------------------------
// The in-pipeline realignment is a <= 300-base query against a ~2.3 kb window (reference src/create_new_contigs.cpp:
// 557-630): 5 blocks, so one pair per warp leaves 27 lanes idle; the 1536-base chunks of the benchmark shape fill 24.
// Here a pair of `nb` blocks takes gl = ceil(nb / BPL) lanes (BPL = 1..4 blocks of 64 rows per lane, stacked: the
// horizontal delta between the blocks of a lane stays in a register) and a warp runs floor(32 / gl) pairs of the same
// block count, all groups in step: 1536 bases = 24 blocks = 8 lanes x 3 blocks, four pairs per warp, no idle lane,
// one shuffle per three block steps. One symbol table serves the whole batch (the numbering of the symbols never
// shows in edlib's results; the per-pair alphabet size comes from edlib_alphabet_kernel).
// Traceback vectors are kept only where the path can be: a cell (row i, column j) of an alignment of distance d has
// |i - j| <= d, so block b (rows 64b+1..64b+64) needs columns 64b-d .. 64b+63+d (0-based) -- band_w = min(n, 64 + 2d)
// columns per block starting at band_first, instead of all n.
__device__ __forceinline__ int band_first(int b, int d, int band_w, int n) { return min(max(64 * b - d, 0), n - band_w); }

struct PackLane {
    int nb, gl, bl, ghead;  // blocks per pair, lanes per pair, this lane's place in its group, first lane of the group
    bool active;
};

__global__ void __launch_bounds__(256) edlib_alphabet_kernel(int n_pairs, const uint8_t* __restrict__ q,
                                                             const int64_t* __restrict__ q_off, const uint8_t* __restrict__ t,
                                                             const int64_t* __restrict__ t_off, int32_t* __restrict__ alpha_len,
                                                             unsigned int* __restrict__ batch_alpha) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    unsigned int all[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int pair = gw; pair < n_pairs; pair += nw) {
        unsigned int present[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int64_t q0 = q_off[pair], q1 = q_off[pair + 1], t0 = t_off[pair], t1 = t_off[pair + 1];
        for (int64_t i = q0 + lane; i < q1; i += 32) { const int c = q[i]; present[c >> 5] |= 1u << (c & 31); }
        for (int64_t i = t0 + lane; i < t1; i += 32) { const int c = t[i]; present[c >> 5] |= 1u << (c & 31); }
        int total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const unsigned int v = __reduce_or_sync(0xffffffffu, present[w]);
            total += __popc(v);
            all[w] |= v;
        }
        if (lane == 0) alpha_len[pair] = total;
    }
    if (lane < 8) {
        unsigned int v = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) if (w == lane) v = all[w];
        if (v) atomicOr(batch_alpha + lane, v);
    }
}

// lut over the batch's bytes (same numbering rule as build_alphabet); returns the number of symbols
__device__ int build_batch_lut(const unsigned int* batch_alpha, uint8_t* lut, int lane) {
    int base = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const unsigned int v = batch_alpha[w];
        lut[w * 32 + lane] = (uint8_t)(base + __popc(v & ((1u << lane) - 1u)));
        base += __popc(v);
    }
    __syncwarp();
    return base;
}

// Peq of the packed kernels: [sym][k][lane], k = the lane's k-th block; every lane fills its own entries
template <int BPL>
__device__ void build_peq_packed(uint64_t* peq, const uint8_t* lut, int lane, const PackLane& g, const uint8_t* q, int m,
                                 int n_sym, bool reversed) {
    for (int i = 0; i < n_sym * BPL; i++) peq[i * 32 + lane] = 0ull;
    if (g.active) {
#pragma unroll
        for (int k = 0; k < BPL; k++) {
            const int r0 = (g.bl * BPL + k) * 64;
            const int r1 = min(m, r0 + 64);
            for (int r = r0; r < r1; r++) {
                const int c = reversed ? q[m - 1 - r] : q[r];
                peq[(lut[c] * BPL + k) * 32 + lane] |= 1ull << (r - r0);
            }
        }
    }
    __syncwarp();
}

// dp_pass with per-lane pair parameters; n_max = the longest target of the warp's groups. The result is valid in the
// lanes of the group it belongs to.
// store of a traceback entry under a predicate, without a branch (the band test differs from block to block)
__device__ __forceinline__ void st_if(bool on, ulonglong2* p, uint64_t x, uint64_t y) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.global.v2.u64 [%1], {%2, %3}; }" ::"r"((unsigned)on), "l"(p),
                 "l"(x), "l"(y)
                 : "memory");
}

// KL = the block of the bottom row inside the group's last lane, (nb - 1) % BPL: the same for every pair of a launch
template <int KIND, int BPL, int KL>
__device__ PassOut dp_pass_packed(const uint64_t* peq, const uint8_t* lut, int lane, const PackLane& g, int m,
                                  const uint8_t* t, int n, int n_max, int rev_end, int start_hin, bool consider_j0,
                                  unsigned int* bitmask, ulonglong2* trace, int band_d = 0, int band_w = 0) {
    uint64_t Pv[BPL], Mv[BPL];
    int band_lo[BPL];  // PASS_NW_STORE: first column kept of each block (band_first)
#pragma unroll
    for (int k = 0; k < BPL; k++) {
        Pv[k] = ~0ull;
        Mv[k] = 0ull;
        band_lo[k] = band_first(g.bl * BPL + k, band_d, band_w, n);
        if (g.bl * BPL + k >= g.nb) band_lo[k] = 0x40000000;  // padding block: never stored
    }
    const int lb = (m - 1) & 63;
    const int gl = g.gl;
    int score = m;
    PassOut o;
    o.best = consider_j0 ? m : 0x3fffffff;
    o.first_j = 0;
    o.last_j = 0;
    o.count = consider_j0 ? 1 : 0;
    unsigned int bits = consider_j0 ? 1u : 0u;
    unsigned int packed = 0;  // symbol << 2 | (horizontal delta leaving the lane is +1) << 1 | (... is -1)
    const unsigned int head_carry = start_hin > 0 ? 2u : 0u;
    // target symbols: gl at a time per group (lane bl holds symbol chunk*gl + bl), fetched one chunk ahead
    int tchunk = 0, tnext = 0;
    if (g.active && g.bl < n) tchunk = lut[rev_end >= 0 ? t[rev_end - g.bl] : t[g.bl]];
    if (g.active && gl + g.bl < n) tnext = lut[rev_end >= 0 ? t[rev_end - gl - g.bl] : t[gl + g.bl]];
    int sc = 0;
    const int steps = n_max + gl - 1;
    for (int s = 0; s < steps; s++) {
        const unsigned int sym0 = (unsigned)__shfl_sync(0xffffffffu, tchunk, g.ghead + sc);
        unsigned int pk = __shfl_up_sync(0xffffffffu, packed, 1);
        if (g.bl == 0) pk = (sym0 << 2) | head_carry;
        const unsigned int sym = pk >> 2;
        uint64_t hneg = pk & 1u, hpos = (pk >> 1) & 1u;
        const int c = s - g.bl;
        if (g.active && c >= 0 && c < n) {
            uint64_t phs = 0, mhs = 0;  // Ph / Mh of the block that holds the bottom row
#pragma unroll
            for (int k = 0; k < BPL; k++) {
                uint64_t Eq = peq[(sym * BPL + k) * 32 + lane];
                const uint64_t Xv = Eq | Mv[k];
                Eq |= hneg;
                const uint64_t Xh = (((Eq & Pv[k]) + Pv[k]) ^ Pv[k]) | Eq;
                const uint64_t Ph = Mv[k] | ~(Xh | Pv[k]);
                const uint64_t Mh = Pv[k] & Xh;
                if (k == KL) { phs = Ph; mhs = Mh; }
                const uint64_t Phs = (Ph << 1) | hpos;
                const uint64_t Mhs = (Mh << 1) | hneg;
                hpos = Ph >> 63;
                hneg = Mh >> 63;
                Pv[k] = Mhs | ~(Xv | Phs);
                Mv[k] = Phs & Xv;
                if (KIND == PASS_NW_STORE) {
                    const int x = c - band_lo[k];
                    st_if((unsigned)x < (unsigned)band_w, trace + (size_t)(g.bl * BPL + k) * band_w + x, Pv[k], Ph);
                }
            }
            packed = (sym << 2) | ((unsigned)hpos << 1) | (unsigned)hneg;
            if (g.bl == gl - 1) {
                score += (int)((phs >> lb) & 1ull) - (int)((mhs >> lb) & 1ull);
                const int j = c + 1;
                if (KIND == PASS_SEMIGLOBAL) {
                    if ((j & 31) == 0) bits = 0;
                    if (score < o.best) { o.best = score; o.first_j = j; o.count = 0; }
                    if (score == o.best) { o.count++; o.last_j = j; bits |= 1u << (j & 31); }
                    if ((j & 31) == 31 || j == n) bitmask[j >> 5] = bits;
                } else if (KIND == PASS_REV_SHW) {
                    if (score < o.best) { o.best = score; o.first_j = j; }
                    if (score == o.best) o.last_j = j;
                } else {
                    o.best = score;
                }
            }
        }
        if (++sc == gl) {
            sc = 0;
            tchunk = tnext;
            const int idx = s + 1 + gl + g.bl;
            tnext = 0;
            if (g.active && idx < n) tnext = lut[rev_end >= 0 ? t[rev_end - idx] : t[idx]];
        }
    }
    const int rl = g.ghead + gl - 1;
    o.best = __shfl_sync(0xffffffffu, o.best, rl);
    o.first_j = __shfl_sync(0xffffffffu, o.first_j, rl);
    o.last_j = __shfl_sync(0xffffffffu, o.last_j, rl);
    o.count = __shfl_sync(0xffffffffu, o.count, rl);
    return o;
}

// ---- banded sweeps of phase B -------------------------------------------------------------------------------------
// Once the distance d is known, only cells with |row - column| <= d can lie on a path of that distance (the cells
// whose value is at most d, and they keep their exact values when everything outside is taken as too large: Ukkonen).
// Group i of a pair = its blocks i*BPL .. i*BPL + BPL - 1; it needs the columns 64*BPL*i - d .. 64*BPL*(i+1) - 1 + d and
// meets column c at step c + i, like in the full sweep. A pair therefore never has more than a few groups at work, and
// gl lanes take its groups in turn: lane r does groups r, r + gl, r + 2 gl, ... (gl is chosen so that group i + gl
// starts after group i has ended). A group that starts takes the column to its left as all +1 below the bottom score
// of the group above, which it picked up from the ring one step earlier; where the group above has already ended, the
// delta coming in from above is +1 (edlib does the same when its band moves, edlib.cpp:781-860). Every lane fetches its
// own target symbols. Peq sits in shared memory per pair as [symbol][block].
#define ED_BAND_PEQ_WORDS 1024  // per warp: pairs x symbols x blocks

template <bool REV>
__device__ void build_peq_band(uint64_t* peq_pair, const uint8_t* lut, const PackLane& g, const uint8_t* q, int m, int n_sym) {
    if (g.active) {
        for (int i = g.bl; i < n_sym * g.nb; i += g.gl) peq_pair[i] = 0ull;
    }
    __syncwarp();
    if (g.active) {
        for (int b = g.bl; b < g.nb; b += g.gl) {
            const int r0 = 64 * b, r1 = min(m, r0 + 64);
            for (int r = r0; r < r1; r++) {
                const int c = REV ? q[m - 1 - r] : q[r];
                peq_pair[lut[c] * g.nb + b] |= 1ull << (r - r0);
            }
        }
    }
    __syncwarp();
}

template <int KIND, int BPL, int KL>
__device__ PassOut dp_pass_band(const uint64_t* peq_pair, const uint8_t* lut, const PackLane& g, int m, const uint8_t* t,
                                int n, int n_max, int rev_end, bool consider_j0, int d, ulonglong2* trace, int band_w) {
    const int nb = g.nb, gl = g.gl;
    const int n_grp = (nb + BPL - 1) / BPL;
    const int i_last = n_grp - 1;
    const int lb = (m - 1) & 63;
    const int prev_lane = g.ghead + (g.bl + gl - 1) % gl;
    uint64_t Pv[BPL], Mv[BPL];
    int band_lo[BPL];
    int i_cur = g.bl, lo = 0, hi = -1;
    int top = 0;      // score of the row above the group at column lo - 1
    int bs = 0;       // score of the group's bottom row (of row m for the last group) at the column just done
    unsigned int symq = 0, symq_next = 0;  // the next four target symbols (a byte each) and the four after them
    int symq_left = 0;
    bool alive = g.active && i_cur < n_grp;
    PassOut o;
    o.best = consider_j0 ? m : 0x3fffffff;
    o.first_j = 0;
    o.last_j = 0;
    o.count = consider_j0 ? 1 : 0;
    auto fetch = [&](int c) -> unsigned int { return c < n ? (unsigned int)lut[rev_end >= 0 ? t[rev_end - c] : t[c]] : 0u; };
    auto fetch4 = [&](int c) -> unsigned int { return fetch(c) | (fetch(c + 1) << 8) | (fetch(c + 2) << 16) | (fetch(c + 3) << 24); };
    auto prime = [&](int c) {  // the group starts at column c
        symq = fetch4(c);
        symq_next = fetch4(c + 4);
        symq_left = 4;
    };
    auto enter = [&](int i) {
#pragma unroll
        for (int k = 0; k < BPL; k++) {
            Pv[k] = ~0ull;
            Mv[k] = 0ull;
            band_lo[k] = KIND == PASS_NW_STORE ? band_first(i * BPL + k, d, band_w, n) : 0;
            if (i * BPL + k >= nb) band_lo[k] = 0x40000000;  // padding block: never stored
        }
        lo = max(0, 64 * BPL * i - d);
        hi = min(n - 1, 64 * BPL * (i + 1) - 1 + d);
        if (lo == 0) {  // the group starts at the matrix's left edge: D[r][0] = r
            top = 64 * BPL * i;
            prime(0);
        }
    };
    if (alive) enter(i_cur);
    unsigned int packed = 0;  // (bottom score << 2) | (horizontal delta leaving the group is +1) << 1 | (... is -1)
    const int steps = n_max + n_grp - 1;
    for (int s = 0; s < steps; s++) {
        const unsigned int pk = __shfl_sync(0xffffffffu, packed, prev_lane);
        if (alive && s - i_cur > hi) {  // this group is done: the lane's next one
            i_cur += gl;
            alive = i_cur < n_grp;
            if (alive) enter(i_cur);
        }
        const int c = s - i_cur;
        if (alive && lo > 0 && c == lo - 1) {  // one step before the group starts: what the group above had at column lo - 1
            top = (int)(pk >> 2);
            prime(lo);
        }
        if (alive && c >= lo && c <= hi) {
            if (c == lo) bs = top + (i_cur == i_last ? 64 * KL + lb + 1 : 64 * BPL);
            const int sym = (int)(symq & 255u);
            symq >>= 8;
            if (--symq_left == 0) {  // the loads of the refill are four steps ahead of their first use
                symq = symq_next;
                symq_next = fetch4(c + 5);
                symq_left = 4;
            }
            // the delta coming in from above: the group above at this column, if it still has it
            uint64_t hneg = 0, hpos = 1;
            if (i_cur > 0 && c <= 64 * BPL * i_cur - 1 + d) {
                hneg = pk & 1u;
                hpos = (pk >> 1) & 1u;
            }
            uint64_t phs = 0, mhs = 0;
            const uint64_t* pq = peq_pair + sym * nb + i_cur * BPL;
#pragma unroll
            for (int k = 0; k < BPL; k++) {
                uint64_t Eq = i_cur * BPL + k < nb ? pq[k] : 0ull;
                const uint64_t Xv = Eq | Mv[k];
                Eq |= hneg;
                const uint64_t Xh = (((Eq & Pv[k]) + Pv[k]) ^ Pv[k]) | Eq;
                const uint64_t Ph = Mv[k] | ~(Xh | Pv[k]);
                const uint64_t Mh = Pv[k] & Xh;
                if (k == KL) { phs = Ph; mhs = Mh; }
                const uint64_t Phs = (Ph << 1) | hpos;
                const uint64_t Mhs = (Mh << 1) | hneg;
                hpos = Ph >> 63;
                hneg = Mh >> 63;
                Pv[k] = Mhs | ~(Xv | Phs);
                Mv[k] = Phs & Xv;
                if (KIND == PASS_NW_STORE) {
                    const int x = c - band_lo[k];
                    st_if((unsigned)x < (unsigned)band_w, trace + (size_t)(i_cur * BPL + k) * band_w + x, Pv[k], Ph);
                }
            }
            if (i_cur == i_last) {
                bs += (int)((phs >> lb) & 1ull) - (int)((mhs >> lb) & 1ull);
                if (KIND == PASS_REV_SHW) {
                    const int j = c + 1;
                    if (bs < o.best) { o.best = bs; o.first_j = j; }
                    if (bs == o.best) o.last_j = j;
                } else {
                    o.best = bs;
                }
            } else {
                bs += (int)hpos - (int)hneg;
            }
            packed = ((unsigned)bs << 2) | ((unsigned)hpos << 1) | (unsigned)hneg;
        }
    }
    const int rl = g.ghead + i_last % gl;  // the lane that did the last group
    o.best = __shfl_sync(0xffffffffu, o.best, rl);
    o.first_j = __shfl_sync(0xffffffffu, o.first_j, rl);
    o.last_j = __shfl_sync(0xffffffffu, o.last_j, rl);
    o.count = __shfl_sync(0xffffffffu, o.count, rl);
    return o;
}

// the groups of a warp walk their paths in step (same rules as traceback above; the window is gl columns wide)
__device__ int traceback_packed(const ulonglong2* trace, const uint8_t* q, int m, const uint8_t* t, int n, uint8_t* out,
                                const PackLane& g, int band_d, int band_w) {
    const int gl = g.gl;
    int i = g.active ? m : 0, j = g.active ? n : 0, len = 0, fl = 0;
    unsigned int pend = 0;  // op of output position len - fl + bl
    int win_block = -1, win_j0 = -1;
    uint64_t wPv = 0, wPh = 0;
    while (__any_sync(0xffffffffu, i > 0 && j > 0)) {
        const bool go = i > 0 && j > 0;
        if (go) {
            const int b = (i - 1) >> 6;
            if (b != win_block || j > win_j0 || j <= win_j0 - gl) {
                win_block = b;
                win_j0 = j;
                const int x = j - g.bl - 1 - band_first(b, band_d, band_w, n);  // lane bl of the group holds column j0 - bl
                ulonglong2 v = make_ulonglong2(0ull, 0ull);
                if ((unsigned)x < (unsigned)band_w) v = trace[(size_t)b * band_w + x];  // columns off the band are never walked
                wPv = v.x;
                wPh = v.y;
            }
        }
        const int src = g.ghead + (go ? win_j0 - j : 0);
        const uint64_t pv = shfl64(wPv, src), ph = shfl64(wPh, src);
        if (go) {
            const int bit = (i - 1) & 63;
            int op;
            if ((pv >> bit) & 1ull) op = 1;
            else if ((ph >> bit) & 1ull) op = 2;
            else op = (q[i - 1] == t[j - 1]) ? 0 : 3;
            i -= op != 2;
            j -= op != 1;
            if (fl == g.bl) pend = (unsigned)op;
            len++;
            if (++fl == gl) {
                fl = 0;
                out[len - gl + g.bl] = (uint8_t)pend;
            }
        }
    }
    if (g.active) {
        if (g.bl < fl) out[len - fl + g.bl] = (uint8_t)pend;
        const int rest = i > 0 ? i : j;
        const uint8_t rest_op = i > 0 ? 1 : 2;
        for (int x = g.bl; x < rest; x += gl) out[len + x] = rest_op;
        len += rest;
    }
    return len;
}

template <int BPL>
__device__ __forceinline__ PackLane pack_lane(int lane, int nb, int cnt, int band_lanes = 0) {
    PackLane g;
    g.nb = nb;
    g.gl = band_lanes ? band_lanes : (nb + BPL - 1) / BPL;
    const int grp = lane / g.gl;
    g.bl = lane - grp * g.gl;
    g.ghead = grp * g.gl;
    g.active = grp < cnt;
    return g;
}

template <int BPL, int KL>
__global__ void __launch_bounds__(ED_WARPS * 32) edlib_phase_a_packed_kernel(EdArgs a) {
    __shared__ uint64_t s_peq[ED_WARPS][ED_SMEM_SYMS * 32 * BPL];
    __shared__ uint8_t s_lut[ED_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * ED_WARPS + wid;
    const uint8_t* lut = s_lut[wid];
    const int n_sym = build_batch_lut(a.batch_alpha, s_lut[wid], lane);
    uint64_t* peq = n_sym <= ED_SMEM_SYMS ? s_peq[wid] : a.peq_big + (size_t)gw * 256 * 32 * ED_PACK_MAXBPL;
    for (;;) {
        int task = 0;
        if (lane == 0) task = (int)atomicAdd(a.counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const int first = a.tasks[2 * task], nbc = a.tasks[2 * task + 1];
        const PackLane g = pack_lane<BPL>(lane, nbc >> 8, nbc & 255);
        const int grp = lane / g.gl;
        const int pair = g.active ? a.plist[first + grp] : 0;
        const uint8_t* q = a.q + a.q_off[pair];
        const uint8_t* t = a.t + a.t_off[pair];
        const int m = g.active ? (int)(a.q_off[pair + 1] - a.q_off[pair]) : 1;
        const int n = g.active ? (int)(a.t_off[pair + 1] - a.t_off[pair]) : 0;
        unsigned int* bm = a.bitmask + a.bm_off[pair];
        __syncwarp();
        build_peq_packed<BPL>(peq, lut, lane, g, q, m, n_sym, false);
        hsgpu_edlib_result r;
        r.status = 0;
        r.edit_distance = -1;
        r.n_locations = 0;
        r.alignment_length = 0;
        r.alphabet_length = g.active ? a.alpha_len[pair] : 0;
        r.has_start_locations = 0;
        r.loc_off = 0;
        r.aln_off = 0;
        const bool unbounded = a.k < 0;
        int kk = unbounded ? 0x3fffffff : a.k;
        if (a.mode == 0) {
            const bool run = g.active && !(kk < abs(n - m));  // :741-744
            kk = min(kk, max(m, n));
            const int n_max = __reduce_max_sync(0xffffffffu, run ? n : 0);
            PackLane gr = g;
            gr.active = run;
            const PassOut o = dp_pass_packed<PASS_NW_SCORE, BPL, KL>(peq, lut, lane, gr, m, t, n, n_max, -1, 1, false, nullptr, nullptr);
            if (run && o.best <= kk) {
                r.edit_distance = o.best;
                r.n_locations = 1;
                if (g.bl == 0) bm[n >> 5] = 1u << (n & 31);
            }
        } else {
            const bool j0 = (m & 63) != 0;  // W > 0
            const int n_max = __reduce_max_sync(0xffffffffu, n);
            const PassOut o = dp_pass_packed<PASS_SEMIGLOBAL, BPL, KL>(peq, lut, lane, g, m, t, n, n_max, -1, a.mode == 2 ? 0 : 1, j0,
                                                                   bm, nullptr);
            if (a.mode == 2) kk = min(kk, m);  // :565-567
            if (o.best <= kk) {
                r.edit_distance = o.best;
                r.n_locations = o.count;
                r.loc_off = o.first_j;
            }
        }
        if (g.active && g.bl == 0) a.res[pair] = r;
    }
}

// BAND: the two sweeps run inside the band of the known distance (dp_pass_band); the task word carries the lanes per pair
template <int BPL, int KL, bool BAND>
__global__ void __launch_bounds__(ED_WARPS * 32) edlib_phase_b_packed_kernel(EdArgs a) {
    __shared__ uint64_t s_peq[ED_WARPS][BAND ? ED_BAND_PEQ_WORDS : ED_SMEM_SYMS * 32 * BPL];
    __shared__ uint8_t s_lut[ED_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * ED_WARPS + wid;
    const uint8_t* lut = s_lut[wid];
    const int n_sym = build_batch_lut(a.batch_alpha, s_lut[wid], lane);
    uint64_t* peq = (BAND || n_sym <= ED_SMEM_SYMS) ? s_peq[wid] : a.peq_big + (size_t)gw * 256 * 32 * ED_PACK_MAXBPL;
    for (;;) {
        int task = 0;
        if (lane == 0) task = (int)atomicAdd(a.counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= a.n_tasks) break;
        const int first = a.tasks[2 * task], nbc = a.tasks[2 * task + 1];
        const PackLane g = pack_lane<BPL>(lane, (nbc >> 8) & 255, nbc & 255, BAND ? (nbc >> 16) : 0);
        const int grp = lane / g.gl;
        uint64_t* const peq_pair = peq + (BAND ? grp * n_sym * g.nb : 0);  // BAND: this pair's [symbol][block]
        const int pair = g.active ? a.plist[first + grp] : 0;
        hsgpu_edlib_result r = a.res[pair];
        const uint8_t* q = a.q + a.q_off[pair];
        const uint8_t* t = a.t + a.t_off[pair];
        const int m = g.active ? (int)(a.q_off[pair + 1] - a.q_off[pair]) : 1;
        const int n = g.active ? (int)(a.t_off[pair + 1] - a.t_off[pair]) : 0;
        const unsigned int* bm = a.bitmask + a.bm_off[pair];
        int32_t* ends = a.ends + r.loc_off;
        int32_t* starts = a.starts + r.loc_off;
        const int n_loc = g.active ? r.n_locations : 0;
        __syncwarp();
        if (g.active && g.bl == 0) {  // end locations: set bits j >= first_j, position = j - 1
            const int first_j = r.aln_off < 0 ? 0 : (int)r.aln_off;
            int found = 0;
            for (int j0 = first_j & ~31; j0 <= n && found < n_loc; j0 += 32) {
                unsigned int word = bm[j0 >> 5];
                if (j0 < first_j) word &= ~((1u << (first_j - j0)) - 1u);
                while (word && found < n_loc) {
                    const int bpos = __ffs(word) - 1;
                    word &= word - 1;
                    ends[found++] = j0 + bpos - 1;
                }
            }
        }
        __syncwarp();
        if (a.task == 0) continue;
        r.has_start_locations = 1;
        if (a.mode == 2) {
            if (BAND) build_peq_band<true>(peq_pair, lut, g, q, m, n_sym);
            else build_peq_packed<BPL>(peq, lut, lane, g, q, m, n_sym, true);
            const bool j0c = (m & 63) != 0;
            const int loc_max = __reduce_max_sync(0xffffffffu, n_loc);
            for (int l = 0; l < loc_max; l++) {
                const int e = l < n_loc ? ends[l] : -1;
                PackLane gr = g;
                gr.active = g.active && e >= 0;
                // an alignment of distance d takes at most m + d target symbols: no later column can reach d again
                const int nrev = min(e + 1, m + r.edit_distance);
                const int n_max = __reduce_max_sync(0xffffffffu, gr.active ? nrev : 0);
                const PassOut o = BAND ? dp_pass_band<PASS_REV_SHW, BPL, KL>(peq_pair, lut, gr, m, t, nrev, n_max, e, j0c,
                                                                              r.edit_distance, nullptr, 0)
                                       : dp_pass_packed<PASS_REV_SHW, BPL, KL>(peq, lut, lane, gr, m, t, nrev, n_max, e, 1, j0c,
                                                                                nullptr, nullptr);
                if (g.bl == 0 && l < n_loc) starts[l] = e >= 0 ? e - (o.last_j - 1) : 0;  // :254-256 last position
            }
        } else if (g.active) {
            for (int l = g.bl; l < n_loc; l += g.gl) starts[l] = 0;
        }
        __syncwarp();
        if (a.task == 2) {
            const int s0 = g.active ? starts[0] : 0, e0 = g.active ? ends[0] : -1;
            const int an = e0 - s0 + 1;
            uint8_t* out = a.aln_tmp + a.aln_tmp_off[pair];
            // the groups share the warp's traceback slab
            const bool walk = g.active && an > 0;
            const int band_d = r.edit_distance, band_w = min(an, 64 + 2 * band_d);
            const long long need = walk ? (long long)g.nb * band_w : 0;  // entries of 16 bytes
            const long long mine = g.bl == 0 ? need : 0;
            long long incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const long long base = __shfl_sync(0xffffffffu, incl - mine, g.ghead);
            PackLane gt = g;
            gt.active = walk && (base + need) * 16 <= a.ptrace_stride;
            if (walk && !gt.active) r.status = 1;  // cannot happen: the host sizes the slab by an upper bound of the same sum
            ulonglong2* trace = reinterpret_cast<ulonglong2*>(a.ptrace + (size_t)gw * a.ptrace_stride) + base;
            if (g.active && an <= 0) {  // obtainAlignment's empty-target case (:1173-1180)
                for (int i = g.bl; i < m; i += g.gl) out[i] = 1;
                r.alignment_length = m;
            }
            const int n_max = __reduce_max_sync(0xffffffffu, gt.active ? an : 0);
            if (BAND) {
                build_peq_band<false>(peq_pair, lut, g, q, m, n_sym);
                dp_pass_band<PASS_NW_STORE, BPL, KL>(peq_pair, lut, gt, m, t + s0, an, n_max, -1, false, band_d, trace, band_w);
            } else {
                build_peq_packed<BPL>(peq, lut, lane, g, q, m, n_sym, false);
                dp_pass_packed<PASS_NW_STORE, BPL, KL>(peq, lut, lane, gt, m, t + s0, an, n_max, -1, 1, false, nullptr, trace,
                                                       band_d, band_w);
            }
            __syncwarp();
            const int len = traceback_packed(trace, q, m, t + s0, an, out, gt, band_d, band_w);
            if (gt.active) r.alignment_length = len;
            __syncwarp();
        }
        if (g.active && g.bl == 0) a.res[pair] = r;
    }
}

// score of row i in a last column kept as per-block (Pv, Mv) with the blocks' prefix sums: D[i][.] = top + the
// vertical deltas of rows 0..i-1
__device__ __forceinline__ int column_score(const ulonglong2* col, const int* pre, int top, int i) {
    const int b = i >> 6, k = i & 63;
    int v = top + pre[b];
    if (k) {
        const ulonglong2 c = col[b];
        const uint64_t mask = (1ull << k) - 1ull;
        v += __popcll(c.x & mask) - __popcll(c.y & mask);
    }
    return v;
}

__device__ void column_prefix(const ulonglong2* col, int* pre, int nb, int lane) {
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        int v = 0;
        if (b < nb) {
            const ulonglong2 c = col[b];
            v = __popcll(c.x) - __popcll(c.y);
        }
        const int inc = hs_warp_incl_scan(v, lane);
        if (b < nb) pre[b + 1] = carry + inc;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) pre[0] = 0;
    __syncwarp();
}

struct HbPart {
    int q0, m, t0, n, score;
};

// obtainAlignment (:1168-1230) for the NW problem q[0..m) x t[0..n) with distance `best`: ops in reverse order at
// out, returns their number or -1 (no split row / stack exhausted: edlib's EDLIB_STATUS_ERROR)
__device__ int path_any_size(WarpCtx& w, const EdArgs& a, int gw, const uint8_t* q, int m, const uint8_t* t, int n, int best,
                             int n_sym, uint8_t* out, HbPart* stack) {
    const int lane = w.lane;
    ulonglong2* trace = reinterpret_cast<ulonglong2*>(a.trace + (size_t)gw * ED_TRACE_BYTES);
    ulonglong2* colL = a.colv + (size_t)gw * 2 * a.col_stride;
    ulonglong2* colR = colL + a.col_stride;
    int* preL = a.colpre + (size_t)gw * 2 * (a.col_stride + 1);
    int* preR = preL + a.col_stride + 1;
    int sp = 0, len = 0;
    if (lane == 0) stack[0] = HbPart{0, m, 0, n, best};
    sp = 1;
    __syncwarp();
    w.n_sym = n_sym;
    while (sp > 0) {
        const HbPart p = stack[--sp];
        __syncwarp();
        const uint8_t* pq = q + p.q0;
        const uint8_t* pt = t + p.t0;
        if (p.m == 0 || p.n == 0) {  // :1173-1180
            const uint8_t op = p.m == 0 ? 2 : 1;
            for (int i = lane; i < p.m + p.n; i += 32) out[len + i] = op;
            len += p.m + p.n;
            continue;
        }
        const int nb = (p.m + 63) >> 6;
        if ((2ll * 8 + 4) * nb * p.n + 8ll * p.n < 1024 * 1024) {  // traceback (:1193-1195)
            w.q = pq;
            w.qrev = false;
            dp_pass<PASS_NW_STORE, true>(w, p.m, nb, pt, p.n, -1, 1, false, nullptr, trace);
            __syncwarp();
            len += traceback(trace, pq, p.m, pt, p.n, out + len, lane);
            __syncwarp();
            continue;
        }
        // Hirschberg (:1236-1401)
        const int left_w = p.n / 2, right_w = p.n - left_w;
        w.q = pq;
        w.qrev = false;
        dp_pass<PASS_NW_COLUMN, true>(w, p.m, nb, pt, left_w, -1, 1, false, nullptr, nullptr, colL);
        w.qrev = true;
        dp_pass<PASS_NW_COLUMN, true>(w, p.m, nb, pt + left_w, right_w, right_w - 1, 1, false, nullptr, nullptr, colR);
        __syncwarp();
        column_prefix(colL, preL, nb, lane);
        column_prefix(colR, preR, nb, lane);
        // first i = row + 1 in 1..m-1 with D_left[i][left_w] + D_right_reversed[m - i][right_w] == score (:1312-1323)
        int split = -1, ls = 0, rs = 0;
        for (int i0 = 1; i0 < p.m && split < 0; i0 += 32) {
            const int i = i0 + lane;
            int l = 0, r = 0;
            bool hit = false;
            if (i < p.m) {
                l = column_score(colL, preL, left_w, i);
                r = column_score(colR, preR, right_w, p.m - i);
                hit = l + r == p.score;
            }
            const unsigned int who = __ballot_sync(0xffffffffu, hit);
            if (who) {
                const int src = __ffs(who) - 1;
                split = i0 + src;
                ls = __shfl_sync(0xffffffffu, l, src);
                rs = __shfl_sync(0xffffffffu, r, src);
            }
        }
        if (split < 0) {
            const int r_all = column_score(colR, preR, right_w, p.m);
            const int l_all = column_score(colL, preL, left_w, p.m);
            if (left_w + r_all == p.score) {  // the whole query goes right (:1325-1333)
                split = 0;
                ls = left_w;
                rs = r_all;
            } else if (l_all + right_w == p.score) {  // the whole query goes left (:1334-1343)
                split = p.m;
                ls = l_all;
                rs = right_w;
            }
        }
        if (split < 0 || sp + 2 > ED_STACK) return -1;
        __syncwarp();
        if (lane == 0) {  // the right part is popped first: ops come out back to front
            stack[sp] = HbPart{p.q0, split, p.t0, left_w, ls};
            stack[sp + 1] = HbPart{p.q0 + split, p.m - split, p.t0 + left_w, right_w, rs};
        }
        sp += 2;
        __syncwarp();
    }
    return len;
}

template <bool LONG>
__global__ void __launch_bounds__(ED_WARPS * 32) edlib_phase_b_kernel(EdArgs a) {
    __shared__ uint64_t s_peq[ED_WARPS][ED_SMEM_SYMS * 32];
    __shared__ uint8_t s_lut[ED_WARPS][256];
    __shared__ HbPart s_stack[LONG ? ED_WARPS : 1][LONG ? ED_STACK : 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * ED_WARPS + wid;
    WarpCtx w;
    w.lane = lane;
    w.lut = s_lut[wid];
    w.qrev = false;
    w.hbuf = LONG ? a.hbuf + (size_t)gw * 2 * a.hbuf_stride : nullptr;
    w.hbuf_stride = a.hbuf_stride;
    const int n_work = LONG ? a.n_list : a.n_pairs;
    for (;;) {
        int pair = 0;
        if (lane == 0) pair = (int)atomicAdd(a.counter, 1u);
        pair = __shfl_sync(0xffffffffu, pair, 0);
        if (pair >= n_work) break;
        if (LONG) pair = a.list[pair];
        const int m = (int)(a.q_off[pair + 1] - a.q_off[pair]);
        const int n = (int)(a.t_off[pair + 1] - a.t_off[pair]);
        if (!LONG && a.route[pair] != ED_ROUTE_ORDINARY) continue;
        hsgpu_edlib_result r = a.res[pair];
        if (r.edit_distance < 0) continue;
        r.status = 0;
        const uint8_t* q = a.q + a.q_off[pair];
        const uint8_t* t = a.t + a.t_off[pair];
        const unsigned int* bm = a.bitmask + a.bm_off[pair];
        int32_t* ends = a.ends + r.loc_off;
        int32_t* starts = a.starts + r.loc_off;
        const int first_j = r.aln_off < 0 ? 0 : (int)r.aln_off;  // phase A parked first_j here (moved by the host)
        // end locations: set bits j >= first_j, position = j - 1
        __syncwarp();
        int found = 0;
        for (int j0 = first_j & ~31; j0 <= n && found < r.n_locations; j0 += 32) {
            unsigned int word = bm[j0 >> 5];
            if (j0 < first_j) word &= ~((1u << (first_j - j0)) - 1u);
            if (lane == 0) {
                while (word && found < r.n_locations) {
                    const int bpos = __ffs(word) - 1;
                    word &= word - 1;
                    ends[found++] = j0 + bpos - 1;
                }
            }
            found = __shfl_sync(0xffffffffu, found, 0);
        }
        __syncwarp();
        if (m == 0 || n == 0 || a.task == 0) continue;
        r.has_start_locations = 1;
        const int nb = (m + 63) >> 6;
        const int n_sym = build_alphabet(q, m, t, n, s_lut[wid], lane);
        w.peq = n_sym <= ED_SMEM_SYMS ? s_peq[wid] : a.peq_big + (size_t)gw * 256 * 32;
        w.q = q;
        w.n_sym = n_sym;
        if (a.mode == 2) {
            if (!LONG) build_peq(w, q, m, nb, n_sym, true);
            w.qrev = true;
            const bool j0c = (m & 63) != 0;
            for (int l = 0; l < r.n_locations; l++) {
                const int e = ends[l];
                int st = 0;
                if (e >= 0) {
                    // an alignment of distance d takes at most m + d target symbols: no later column can reach d again
                    const PassOut o = dp_pass<PASS_REV_SHW, LONG>(w, m, nb, t, min(e + 1, m + r.edit_distance), e, 1, j0c,
                                                                  nullptr, nullptr);
                    st = e - (o.last_j - 1);  // :254-256 last position
                }
                if (lane == 0) starts[l] = st;
            }
            w.qrev = false;
        } else {
            for (int l = lane; l < r.n_locations; l += 32) starts[l] = 0;
        }
        __syncwarp();
        if (a.task == 2) {
            const int s0 = starts[0], e0 = ends[0];
            const int an = e0 - s0 + 1;
            uint8_t* out = a.aln_tmp + a.aln_tmp_off[pair];
            if (an <= 0) {  // obtainAlignment's empty-target case (:1173-1180)
                for (int i = lane; i < m; i += 32) out[i] = 1;
                r.alignment_length = m;
            } else if (LONG) {
                const int len = path_any_size(w, a, gw, q, m, t + s0, an, r.edit_distance, n_sym, out, s_stack[wid]);
                if (len < 0) r.status = 1;
                else r.alignment_length = len;
            } else if ((2ll * 8 + 4) * nb * an + 8ll * an >= 1024 * 1024) {
                r.status = 2;  // Hirschberg regime: redone by the LONG launch
            } else {
                build_peq(w, q, m, nb, n_sym, false);
                ulonglong2* trace = reinterpret_cast<ulonglong2*>(a.trace + (size_t)gw * ED_TRACE_BYTES);
                dp_pass<PASS_NW_STORE, false>(w, m, nb, t + s0, an, -1, 1, false, nullptr, trace);
                __syncwarp();
                r.alignment_length = traceback(trace, q, m, t + s0, an, out, lane);
            }
        }
        if (lane == 0) a.res[pair] = r;
    }
}

// final[aln_off + x] = tmp[len - 1 - x]
__global__ void edlib_gather_kernel(int n_pairs, const hsgpu_edlib_result* __restrict__ res,
                                    const uint8_t* __restrict__ tmp, const int64_t* __restrict__ tmp_off,
                                    uint8_t* __restrict__ out) {
    const int pair = blockIdx.x;
    const int len = res[pair].alignment_length;
    const uint8_t* src = tmp + tmp_off[pair];
    uint8_t* dst = out + res[pair].aln_off;
    for (int x = threadIdx.x; x < len; x += blockDim.x) dst[x] = src[len - 1 - x];
}

extern "C" int hsgpu_edlib_align_batch(hsgpu_ctx* ctx, int32_t n_pairs, const char* queries, const int64_t* query_off,
                                       const char* targets, const int64_t* target_off, int32_t k, int32_t mode,
                                       int32_t task, hsgpu_edlib_result* results, int32_t* end_locations,
                                       int32_t* start_locations, int64_t loc_capacity, uint8_t* alignment,
                                       int64_t aln_capacity) {
    if (!ctx || n_pairs < 0 || !query_off || !target_off || !results) return HSGPU_ERR_ARG;
    if (mode < 0 || mode > 2 || task < 0 || task > 2) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_edlib_align_batch: bad mode/task");
    if (n_pairs == 0) return HSGPU_OK;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int64_t> bm_off((size_t)n_pairs + 1), tmp_off((size_t)n_pairs + 1);
    std::vector<int32_t> long_list;  // queries of more than one strip: both phases run in their LONG instantiation
    std::vector<uint8_t> route((size_t)n_pairs);
    static const bool pack = !getenv("HSGPU_EDLIB_PACK") || atoi(getenv("HSGPU_EDLIB_PACK")) != 0;
    int64_t n_packed = 0, n_ordinary = 0;
    int64_t bmw = 0, tmpb = 0;
    for (int i = 0; i < n_pairs; i++) {
        const int64_t m = query_off[i + 1] - query_off[i], n = target_off[i + 1] - target_off[i];
        if (m < 0 || n < 0) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_edlib_align_batch: offsets must be non-decreasing");
        if (m > ED_MAX_QUERY || n > (1 << 30))
            HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_edlib_align_batch: query longer than 2^20 or target longer than 2^30");
        if (m > 64 * ED_MAXBLOCKS) {
            long_list.push_back(i);
            route[i] = ED_ROUTE_LONG;
        } else if (pack && m >= 1 && m <= 64 * ED_PACK_MAXBLOCKS && n >= 1) {
            route[i] = ED_ROUTE_PACKED;
            n_packed++;
        } else {
            route[i] = ED_ROUTE_ORDINARY;
            n_ordinary++;
        }
        bm_off[i] = bmw;
        tmp_off[i] = tmpb;
        bmw += (n + 1 + 31) / 32 + 1;
        tmpb += (m + n + 32 + 15) & ~15ll;
    }
    bm_off[n_pairs] = bmw;
    tmp_off[n_pairs] = tmpb;
    const int64_t qbytes = query_off[n_pairs], tbytes = target_off[n_pairs];
    static const int ctas_per_sm = getenv("HSGPU_EDLIB_CTAS") ? std::max(1, atoi(getenv("HSGPU_EDLIB_CTAS"))) : 5;
    const int grid = ctx->sm_count * ctas_per_sm;
    const int n_warps = grid * ED_WARPS;
    uint8_t *d_q = nullptr, *d_t = nullptr, *d_aln_tmp = nullptr, *d_trace = nullptr, *d_aln = nullptr, *d_ptrace = nullptr;
    int64_t *d_qo = nullptr, *d_to = nullptr, *d_bmo = nullptr, *d_tmpo = nullptr, *d_scan = nullptr;
    hsgpu_edlib_result* d_res = nullptr;
    unsigned int *d_bm = nullptr, *d_counter = nullptr;
    int32_t *d_ends = nullptr, *d_starts = nullptr;
    uint64_t* d_peq = nullptr;
    HsTemps temps(ctx);  // an early error return releases whatever has been allocated by then
    temps.own(d_q, d_t, d_aln_tmp, d_trace, d_aln, d_ptrace, d_qo, d_to, d_bmo, d_tmpo, d_scan, d_res, d_bm, d_counter, d_ends,
              d_starts, d_peq);
    HS_CUDA(ctx, hs_alloc(ctx, &d_q, qbytes));
    HS_CUDA(ctx, hs_alloc(ctx, &d_t, tbytes));
    HS_CUDA(ctx, hs_alloc(ctx, &d_qo, n_pairs + 1));
    HS_CUDA(ctx, hs_alloc(ctx, &d_to, n_pairs + 1));
    HS_CUDA(ctx, hs_alloc(ctx, &d_bmo, n_pairs + 1));
    HS_CUDA(ctx, hs_alloc(ctx, &d_res, n_pairs));
    // phase A fills some of the fields; the whole struct travels to the host after it
    HS_CUDA(ctx, cudaMemsetAsync(d_res, 0, sizeof(hsgpu_edlib_result) * (size_t)std::max(n_pairs, 1), ctx->stream));
    HS_CUDA(ctx, hs_alloc(ctx, &d_bm, bmw));
    HS_CUDA(ctx, hs_alloc(ctx, &d_counter, 48));
    uint8_t* d_route = nullptr;
    int32_t *d_alpha_len = nullptr, *d_plist = nullptr, *d_tasks = nullptr;
    unsigned int* d_batch_alpha = nullptr;
    temps.own(d_route, d_alpha_len, d_plist, d_tasks, d_batch_alpha);
    HS_CUDA(ctx, hs_alloc(ctx, &d_route, n_pairs));
    HS_CUDA(ctx, hs_h2d(ctx, d_route, route.data(), n_pairs));
    EdArgs a;
    // tasks of the packed launches: pairs of one block count, floor(32 / lanes per pair) to a warp. Blocks per lane:
    // what a block step costs (core work per block + the per-step overhead of a lane, shared by the warp's pairs)
    static int bpl_of[ED_PACK_MAXBLOCKS + 1];
    static std::once_flag bpl_once;
    std::call_once(bpl_once, [] {
        const int forced = getenv("HSGPU_EDLIB_BPL") ? atoi(getenv("HSGPU_EDLIB_BPL")) : 0;
        for (int nb = 1; nb <= ED_PACK_MAXBLOCKS; nb++) {
            double best = 1e30;
            for (int b = 1; b <= ED_PACK_MAXBPL; b++) {
                const int gl = (nb + b - 1) / b;
                const double cost = (34.0 * b + 28.0) / (32 / gl);
                if (cost < best - 1e-9) { best = cost; bpl_of[nb] = b; }
            }
            if (forced >= 1 && forced <= ED_PACK_MAXBPL) bpl_of[nb] = forced;
        }
    });
    std::vector<int32_t> plist, tasks;
    // a launch = one (blocks per lane, block of the bottom row in the last lane) class: 1 + 2 + 3 + 4 of them
    constexpr int NCLS = ED_PACK_MAXBPL * (ED_PACK_MAXBPL + 1) / 2;
    // phase B may run its two sweeps inside the band of the known distance: a plan per pair = (banded?, blocks per
    // lane, lanes per pair); launches are per (banded?, blocks per lane, bottom-row block) class
    static const bool band_on = getenv("HSGPU_EDLIB_BAND") && atoi(getenv("HSGPU_EDLIB_BAND")) != 0;  // off until validated on the GPU
    int n_sym_batch = 256;  // symbols of the batch (known after phase A): bounds the banded kernels' Peq in shared memory
    struct Plan { int cls, nb, gl, per; };  // class index (0 .. 2 NCLS - 1), blocks, lanes per pair (0 = all groups), pairs per task
    auto plan_of = [&](int i, bool phase_b) -> Plan {
        const int nb = (int)((query_off[i + 1] - query_off[i] + 63) / 64);
        const int bu = bpl_of[nb];
        Plan pl{bu * (bu - 1) / 2 + (nb - 1) % bu, nb, 0, 32 / ((nb + bu - 1) / bu)};
        if (!phase_b || !band_on || (mode != 2 && task != 2)) return pl;
        double best = (34.0 * bu + 28.0) / pl.per;
        const int64_t d = results[i].edit_distance;
        for (int b = bu; b <= bu; b++) {  // the same blocks per lane as the full sweeps (one class per batch shape)
            const int n_grp = (nb + b - 1) / b;
            int gl = 2;
            while (gl <= 8 && 64ll * b * (gl - 1) + gl < 2 * d + 4) gl++;  // group i + gl starts after group i has ended
            if (gl > 8 || gl >= n_grp) continue;
            const int per = std::min(32 / gl, ED_BAND_PEQ_WORDS / std::max(1, n_sym_batch * nb));
            if (per < 1) continue;
            const double cost = (34.0 * b + 44.0) / per;
            if (cost < best - 1e-9) {
                best = cost;
                pl = Plan{NCLS + b * (b - 1) / 2 + (nb - 1) % b, nb, gl, per};
            }
        }
        return pl;
    };
    int task_first[2 * NCLS + 1];  // tasks of class c: [task_first[c], task_first[c + 1])
    int64_t ptrace_stride = 0;
    // key of a pair: (class, blocks, lanes per pair, end locations) -- counting sort, then tasks of equal shape
    auto build_tasks = [&](bool phase_b) -> cudaError_t {
        const int NL = ED_PACK_MAXLOC + 1, NB = ED_PACK_MAXBLOCKS + 1, NG = 9;
        const int NK = 2 * NCLS * NB * NG * NL;
        std::vector<int64_t> cnt((size_t)NK + 1, 0);
        std::vector<int32_t> keys((size_t)n_pairs, -1);
        for (int i = 0; i < n_pairs; i++) {
            if (route[i] != ED_ROUTE_PACKED) continue;
            const Plan pl = plan_of(i, phase_b);
            keys[i] = (((pl.cls * NB + pl.nb) * NG + pl.gl) * NL) + (phase_b ? results[i].n_locations : 0);
            cnt[(size_t)keys[i] + 1]++;
        }
        for (int c = 1; c <= NK; c++) cnt[c] += cnt[c - 1];
        const int64_t total = cnt[NK];
        plist.assign((size_t)total, 0);
        std::vector<int64_t> fill(cnt);
        for (int i = 0; i < n_pairs; i++)
            if (keys[i] >= 0) plist[(size_t)fill[(size_t)keys[i]]++] = i;
        tasks.clear();
        ptrace_stride = 0;
        for (int c = 0; c < 2 * NCLS; c++) {
            task_first[c] = (int)(tasks.size() / 2);
            for (int nb = 1; nb <= ED_PACK_MAXBLOCKS; nb++) {
                for (int gl = 0; gl < NG; gl++) {
                    const int64_t lo = cnt[(size_t)((c * NB + nb) * NG + gl) * NL], hi = cnt[(size_t)((c * NB + nb) * NG + gl + 1) * NL];
                    if (lo == hi) continue;
                    const int per = plan_of(plist[(size_t)lo], phase_b).per;  // the same for every pair of the bucket but for d: take the first
                    for (int64_t f = lo; f < hi; f += per) {
                        const int n_in = (int)std::min<int64_t>(per, hi - f);
                        tasks.push_back((int32_t)f);
                        tasks.push_back((gl << 16) | (nb << 8) | n_in);
                        if (phase_b && task == 2) {  // the task's groups share one traceback slab
                            int64_t need = 0;
                            for (int x = 0; x < n_in; x++) {
                                const int i = plist[(size_t)(f + x)];
                                const int64_t m = query_off[i + 1] - query_off[i], n = target_off[i + 1] - target_off[i];
                                const int64_t d = results[i].edit_distance;  // band of the path: see band_first
                                need += nb * std::min<int64_t>(std::min<int64_t>(n, m + d), 64 + 2 * d) * 16;
                            }
                            ptrace_stride = std::max(ptrace_stride, need);
                        }
                    }
                }
            }
        }
        task_first[2 * NCLS] = (int)(tasks.size() / 2);
        ptrace_stride = (ptrace_stride + 255) & ~255ll;
        if (total == 0) return cudaSuccess;
        cudaError_t e = hs_h2d(ctx, d_plist, plist.data(), total);
        if (e == cudaSuccess) e = hs_h2d(ctx, d_tasks, tasks.data(), (int64_t)tasks.size());
        return e;
    };
    typedef void (*PackedKernel)(EdArgs);
    static const PackedKernel phase_a_packed[NCLS] = {
        edlib_phase_a_packed_kernel<1, 0>, edlib_phase_a_packed_kernel<2, 0>, edlib_phase_a_packed_kernel<2, 1>,
        edlib_phase_a_packed_kernel<3, 0>, edlib_phase_a_packed_kernel<3, 1>, edlib_phase_a_packed_kernel<3, 2>,
        edlib_phase_a_packed_kernel<4, 0>, edlib_phase_a_packed_kernel<4, 1>, edlib_phase_a_packed_kernel<4, 2>,
        edlib_phase_a_packed_kernel<4, 3>};
    static const PackedKernel phase_b_packed[2 * NCLS] = {
        edlib_phase_b_packed_kernel<1, 0, false>, edlib_phase_b_packed_kernel<2, 0, false>, edlib_phase_b_packed_kernel<2, 1, false>,
        edlib_phase_b_packed_kernel<3, 0, false>, edlib_phase_b_packed_kernel<3, 1, false>, edlib_phase_b_packed_kernel<3, 2, false>,
        edlib_phase_b_packed_kernel<4, 0, false>, edlib_phase_b_packed_kernel<4, 1, false>, edlib_phase_b_packed_kernel<4, 2, false>,
        edlib_phase_b_packed_kernel<4, 3, false>,
        edlib_phase_b_packed_kernel<1, 0, true>, edlib_phase_b_packed_kernel<2, 0, true>, edlib_phase_b_packed_kernel<2, 1, true>,
        edlib_phase_b_packed_kernel<3, 0, true>, edlib_phase_b_packed_kernel<3, 1, true>, edlib_phase_b_packed_kernel<3, 2, true>,
        edlib_phase_b_packed_kernel<4, 0, true>, edlib_phase_b_packed_kernel<4, 1, true>, edlib_phase_b_packed_kernel<4, 2, true>,
        edlib_phase_b_packed_kernel<4, 3, true>};
    static const char* const phase_a_names[NCLS] = {
        "edlib_phase_a_kernel<packed,1,0>", "edlib_phase_a_kernel<packed,2,0>", "edlib_phase_a_kernel<packed,2,1>",
        "edlib_phase_a_kernel<packed,3,0>", "edlib_phase_a_kernel<packed,3,1>", "edlib_phase_a_kernel<packed,3,2>",
        "edlib_phase_a_kernel<packed,4,0>", "edlib_phase_a_kernel<packed,4,1>", "edlib_phase_a_kernel<packed,4,2>",
        "edlib_phase_a_kernel<packed,4,3>"};
    static const char* const phase_b_names[2 * NCLS] = {
        "edlib_phase_b_kernel<packed,1,0>", "edlib_phase_b_kernel<packed,2,0>", "edlib_phase_b_kernel<packed,2,1>",
        "edlib_phase_b_kernel<packed,3,0>", "edlib_phase_b_kernel<packed,3,1>", "edlib_phase_b_kernel<packed,3,2>",
        "edlib_phase_b_kernel<packed,4,0>", "edlib_phase_b_kernel<packed,4,1>", "edlib_phase_b_kernel<packed,4,2>",
        "edlib_phase_b_kernel<packed,4,3>",
        "edlib_phase_b_kernel<band,1,0>", "edlib_phase_b_kernel<band,2,0>", "edlib_phase_b_kernel<band,2,1>",
        "edlib_phase_b_kernel<band,3,0>", "edlib_phase_b_kernel<band,3,1>", "edlib_phase_b_kernel<band,3,2>",
        "edlib_phase_b_kernel<band,4,0>", "edlib_phase_b_kernel<band,4,1>", "edlib_phase_b_kernel<band,4,2>",
        "edlib_phase_b_kernel<band,4,3>"};
    if (n_packed > 0) {
        HS_CUDA(ctx, hs_alloc(ctx, &d_alpha_len, n_pairs));
        HS_CUDA(ctx, hs_alloc(ctx, &d_plist, n_packed));
        HS_CUDA(ctx, hs_alloc(ctx, &d_tasks, 2 * n_packed));
        HS_CUDA(ctx, hs_alloc(ctx, &d_batch_alpha, 8));
        HS_CUDA(ctx, cudaMemsetAsync(d_batch_alpha, 0, 8 * sizeof(unsigned int), ctx->stream));
    }
    HS_CUDA(ctx, hs_alloc(ctx, &d_peq, (int64_t)n_warps * 256 * 32 * ED_PACK_MAXBPL));
    HS_CUDA(ctx, hs_h2d(ctx, d_q, (const uint8_t*)queries, qbytes));
    HS_CUDA(ctx, hs_h2d(ctx, d_t, (const uint8_t*)targets, tbytes));
    HS_CUDA(ctx, hs_h2d(ctx, d_qo, query_off, n_pairs + 1));
    HS_CUDA(ctx, hs_h2d(ctx, d_to, target_off, n_pairs + 1));
    HS_CUDA(ctx, hs_h2d(ctx, d_bmo, bm_off.data(), n_pairs + 1));
    HS_CUDA(ctx, cudaMemsetAsync(d_bm, 0, sizeof(unsigned int) * bmw, ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(d_counter, 0, 48 * sizeof(unsigned int), ctx->stream));
    a.n_pairs = n_pairs;
    a.q = d_q;
    a.q_off = d_qo;
    a.t = d_t;
    a.t_off = d_to;
    a.k = k;
    a.mode = mode;
    a.task = task;
    a.res = d_res;
    a.bitmask = d_bm;
    a.bm_off = d_bmo;
    a.ends = nullptr;
    a.starts = nullptr;
    a.aln_tmp = nullptr;
    a.aln_tmp_off = nullptr;
    a.peq_big = d_peq;
    a.trace = nullptr;
    a.counter = d_counter;
    a.list = nullptr;
    a.n_list = 0;
    a.hbuf = nullptr;
    a.hbuf_stride = 0;
    a.colv = nullptr;
    a.colpre = nullptr;
    a.col_stride = 0;
    a.route = d_route;
    a.alpha_len = d_alpha_len;
    a.batch_alpha = d_batch_alpha;
    a.plist = d_plist;
    a.tasks = d_tasks;
    a.n_tasks = 0;
    a.ptrace = nullptr;
    a.ptrace_stride = 0;
    if (n_ordinary > 0)
        HS_KERNEL(ctx, "edlib_phase_a_kernel", edlib_phase_a_kernel<false><<<grid, ED_WARPS * 32, 0, ctx->stream>>>(a));
    if (n_packed > 0) {
        HS_KERNEL(ctx, "edlib_alphabet_kernel",
                  edlib_alphabet_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(n_pairs, d_q, d_qo, d_t, d_to, d_alpha_len,
                                                                                  d_batch_alpha));
        HS_CUDA(ctx, build_tasks(false));
        for (int c = 0; c < NCLS; c++) {
            a.tasks = d_tasks + 2 * task_first[c];
            a.n_tasks = task_first[c + 1] - task_first[c];
            if (a.n_tasks == 0) continue;
            a.counter = d_counter + 8 + c;
            const int gp = (int)std::min<int64_t>(grid, (a.n_tasks + ED_WARPS - 1) / ED_WARPS);
            HS_KERNEL(ctx, phase_a_names[c], phase_a_packed[c]<<<gp, ED_WARPS * 32, 0, ctx->stream>>>(a));
        }
    }
    // scratch of the LONG launches, sized for the pairs of the list at hand
    int32_t* d_list = nullptr;
    uint8_t* d_hbuf = nullptr;
    ulonglong2* d_colv = nullptr;
    int* d_colpre = nullptr;
    temps.own(d_list, d_hbuf, d_colv, d_colpre);
    auto long_setup = [&](const std::vector<int32_t>& list, bool with_columns, int* grid_long) -> cudaError_t {
        hs_free(ctx, d_list); hs_free(ctx, d_hbuf); hs_free(ctx, d_colv); hs_free(ctx, d_colpre);
        d_list = nullptr; d_hbuf = nullptr; d_colv = nullptr; d_colpre = nullptr;
        int64_t max_n = 1, max_nb = 1;
        for (int32_t i : list) {
            max_n = std::max<int64_t>(max_n, target_off[i + 1] - target_off[i] + 1);
            max_nb = std::max<int64_t>(max_nb, (query_off[i + 1] - query_off[i] + 63) / 64);
        }
        *grid_long = (int)std::min<int64_t>(grid, ((int64_t)list.size() + ED_WARPS - 1) / ED_WARPS);
        const int64_t nw = (int64_t)*grid_long * ED_WARPS;
        a.hbuf_stride = (max_n + 15) & ~15ll;
        a.col_stride = max_nb;
        cudaError_t e = hs_alloc(ctx, &d_list, (int64_t)list.size());
        if (e == cudaSuccess) e = hs_alloc(ctx, &d_hbuf, nw * 2 * a.hbuf_stride);
        if (e == cudaSuccess && with_columns) e = hs_alloc(ctx, &d_colv, nw * 2 * a.col_stride);
        if (e == cudaSuccess && with_columns) e = hs_alloc(ctx, &d_colpre, nw * 2 * (a.col_stride + 1));
        if (e == cudaSuccess) e = hs_h2d(ctx, d_list, list.data(), (int64_t)list.size());
        a.list = d_list;
        a.n_list = (int)list.size();
        a.hbuf = d_hbuf;
        a.colv = d_colv;
        a.colpre = d_colpre;
        return e;
    };
    if (!long_list.empty()) {
        int grid_long = 0;
        HS_CUDA(ctx, long_setup(long_list, false, &grid_long));
        a.counter = d_counter + 2;
        HS_KERNEL(ctx, "edlib_phase_a_kernel<long>", edlib_phase_a_kernel<true><<<grid_long, ED_WARPS * 32, 0, ctx->stream>>>(a));
    }
    // sizes -> offsets on the host (n_pairs structs; the location lists are usually 1-3 entries each)
    HS_CUDA(ctx, hs_d2h(ctx, results, d_res, n_pairs));
    unsigned int h_alpha[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (d_batch_alpha) HS_CUDA(ctx, hs_d2h(ctx, h_alpha, d_batch_alpha, 8));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    if (d_batch_alpha) {
        n_sym_batch = 0;
        for (int w = 0; w < 8; w++) n_sym_batch += __builtin_popcount(h_alpha[w]);
    }
    int64_t nloc = 0;
    for (int i = 0; i < n_pairs; i++) {
        results[i].aln_off = results[i].loc_off;  // first_j parked by phase A
        results[i].loc_off = nloc;
        nloc += results[i].n_locations;
    }
    int rc = HSGPU_OK;
    if (nloc > loc_capacity || !end_locations || (task >= 1 && !start_locations)) {
        hs_set_error(ctx, "hsgpu_edlib_align_batch: loc_capacity too small");
        rc = HSGPU_ERR_CAPACITY;
    }
    if (rc == HSGPU_OK) {
        HS_CUDA(ctx, hs_alloc(ctx, &d_ends, nloc));
        HS_CUDA(ctx, hs_alloc(ctx, &d_starts, nloc));
        // pairs for which edlib reports no start locations (empty sequences) leave their slots untouched: -1
        HS_CUDA(ctx, cudaMemsetAsync(d_ends, 0xff, sizeof(int32_t) * (size_t)std::max<int64_t>(nloc, 1), ctx->stream));
        HS_CUDA(ctx, cudaMemsetAsync(d_starts, 0xff, sizeof(int32_t) * (size_t)std::max<int64_t>(nloc, 1), ctx->stream));
        HS_CUDA(ctx, hs_h2d(ctx, d_res, results, n_pairs));
        if (task == 2) {
            HS_CUDA(ctx, hs_alloc(ctx, &d_aln_tmp, tmpb));
            HS_CUDA(ctx, hs_alloc(ctx, &d_tmpo, n_pairs + 1));
            HS_CUDA(ctx, hs_h2d(ctx, d_tmpo, tmp_off.data(), n_pairs + 1));
        }
        a.ends = d_ends;
        a.starts = d_starts;
        a.aln_tmp = d_aln_tmp;
        a.aln_tmp_off = d_tmpo;
        a.trace = d_trace;
        // who does phase B: the packed launch keeps the pairs whose few locations and short path fit its shared slab
        n_ordinary = 0;
        n_packed = 0;
        for (int i = 0; i < n_pairs; i++) {
            if (route[i] == ED_ROUTE_LONG) continue;
            const bool was_packed = route[i] == ED_ROUTE_PACKED;
            route[i] = ED_ROUTE_ORDINARY;
            if (results[i].edit_distance < 0) {
                route[i] = ED_ROUTE_NONE;
                continue;
            }
            if (was_packed && results[i].n_locations <= ED_PACK_MAXLOC) {
                // the path must lie below edlib's 1 MiB switch whatever the start location turns out to be
                const int64_t m = query_off[i + 1] - query_off[i], n = target_off[i + 1] - target_off[i];
                const int64_t nb = (m + 63) / 64, bound = std::min<int64_t>(n, m + results[i].edit_distance);
                if (task < 2 || (2 * 8 + 4) * nb * bound + 8 * bound < 1024 * 1024) route[i] = ED_ROUTE_PACKED;
            }
            if (route[i] == ED_ROUTE_PACKED) n_packed++; else n_ordinary++;
        }
        HS_CUDA(ctx, hs_h2d(ctx, d_route, route.data(), n_pairs));
        a.counter = d_counter + 1;
        if (n_ordinary > 0) {
            if (task == 2) HS_CUDA(ctx, hs_alloc(ctx, &d_trace, (int64_t)n_warps * ED_TRACE_BYTES));
            a.trace = d_trace;
            HS_KERNEL(ctx, "edlib_phase_b_kernel", edlib_phase_b_kernel<false><<<grid, ED_WARPS * 32, 0, ctx->stream>>>(a));
        }
        if (n_packed > 0) {
            HS_CUDA(ctx, build_tasks(true));
            // one traceback slab per resident warp, shared by the warp's pairs; at most 16 GB of them
            const int64_t max_warps = std::max<int64_t>(ED_WARPS, (16ll << 30) / std::max<int64_t>(ptrace_stride, 1));
            const int grid_cap = (int)std::min<int64_t>(grid, max_warps / ED_WARPS);
            if (task == 2) HS_CUDA(ctx, hs_alloc(ctx, &d_ptrace, (int64_t)grid_cap * ED_WARPS * ptrace_stride));
            a.ptrace = d_ptrace;
            a.ptrace_stride = ptrace_stride;
            for (int c = 0; c < 2 * NCLS; c++) {
                a.tasks = d_tasks + 2 * task_first[c];
                a.n_tasks = task_first[c + 1] - task_first[c];
                if (a.n_tasks == 0) continue;
                a.counter = d_counter + 20 + c;
                const int gp = (int)std::min<int64_t>(grid_cap, (a.n_tasks + ED_WARPS - 1) / ED_WARPS);
                HS_KERNEL(ctx, phase_b_names[c], phase_b_packed[c]<<<gp, ED_WARPS * 32, 0, ctx->stream>>>(a));
            }
        }
        HS_CUDA(ctx, hs_d2h(ctx, results, d_res, n_pairs));
        if (task == 2 || !long_list.empty()) {
            // the LONG launch: long queries, and the pairs whose path turned out to lie at or above edlib's 1 MiB switch
            HS_CUDA(ctx, hs_stream_sync(ctx));
            std::vector<int32_t> list = long_list;
            for (int i = 0; i < n_pairs; i++)
                if (results[i].status == 2) list.push_back(i);
            if (!list.empty()) {
                int grid_long = 0;
                HS_CUDA(ctx, long_setup(list, task == 2, &grid_long));
                if (task == 2 && !d_trace) HS_CUDA(ctx, hs_alloc(ctx, &d_trace, (int64_t)n_warps * ED_TRACE_BYTES));
                a.trace = d_trace;
                a.counter = d_counter + 3;
                HS_KERNEL(ctx, "edlib_phase_b_kernel<long>",
                          edlib_phase_b_kernel<true><<<grid_long, ED_WARPS * 32, 0, ctx->stream>>>(a));
                HS_CUDA(ctx, hs_d2h(ctx, results, d_res, n_pairs));
            }
        }
        HS_CUDA(ctx, hs_d2h(ctx, end_locations, d_ends, nloc));
        if (task >= 1) HS_CUDA(ctx, hs_d2h(ctx, start_locations, d_starts, nloc));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        int64_t naln = 0;
        for (int i = 0; i < n_pairs; i++) {
            results[i].aln_off = naln;
            naln += results[i].alignment_length;
        }
        if (task == 2) {
            if (naln > aln_capacity || !alignment) {
                hs_set_error(ctx, "hsgpu_edlib_align_batch: aln_capacity too small");
                rc = HSGPU_ERR_CAPACITY;
            } else if (naln > 0) {
                HS_CUDA(ctx, hs_alloc(ctx, &d_aln, naln));
                HS_CUDA(ctx, hs_h2d(ctx, d_res, results, n_pairs));
                HS_KERNEL(ctx, "edlib_gather_kernel",
                          edlib_gather_kernel<<<n_pairs, 128, 0, ctx->stream>>>(n_pairs, d_res, d_aln_tmp, d_tmpo, d_aln));
                HS_CUDA(ctx, hs_d2h(ctx, alignment, d_aln, naln));
                HS_CUDA(ctx, hs_stream_sync(ctx));
            }
        }
    }
    hs_free(ctx, d_q); hs_free(ctx, d_t); hs_free(ctx, d_qo); hs_free(ctx, d_to); hs_free(ctx, d_bmo);
    hs_free(ctx, d_res); hs_free(ctx, d_bm); hs_free(ctx, d_counter); hs_free(ctx, d_peq); hs_free(ctx, d_ends);
    hs_free(ctx, d_starts); hs_free(ctx, d_aln_tmp); hs_free(ctx, d_tmpo); hs_free(ctx, d_trace); hs_free(ctx, d_aln);
    hs_free(ctx, d_ptrace); hs_free(ctx, d_route); hs_free(ctx, d_alpha_len); hs_free(ctx, d_plist); hs_free(ctx, d_tasks); hs_free(ctx, d_batch_alpha);
    hs_free(ctx, d_scan); hs_free(ctx, d_list); hs_free(ctx, d_hbuf); hs_free(ctx, d_colv); hs_free(ctx, d_colpre);
    return rc;
}

extern "C" hsgpu_EdlibAlignResult hsgpu_edlibAlign(hsgpu_ctx* ctx, const char* query, int queryLength, const char* target,
                                                   int targetLength, hsgpu_EdlibAlignConfig config) {
    hsgpu_EdlibAlignResult out;
    out.status = 1;  // EDLIB_STATUS_ERROR
    out.editDistance = -1;
    out.endLocations = out.startLocations = nullptr;
    out.numLocations = 0;
    out.alignment = nullptr;
    out.alignmentLength = 0;
    out.alphabetLength = 0;
    if (!ctx || queryLength < 0 || targetLength < 0 || (!query && queryLength > 0) || (!target && targetLength > 0)) return out;
    if (config.additionalEqualitiesLength > 0) {
        hs_set_error(ctx, "hsgpu_edlibAlign: additional equalities are not supported");
        return out;
    }
    const int64_t qo[2] = {0, queryLength}, to[2] = {0, targetLength};
    const int64_t loc_cap = (int64_t)targetLength + 2, aln_cap = (int64_t)queryLength + targetLength + 8;
    int32_t* ends = static_cast<int32_t*>(malloc(sizeof(int32_t) * (size_t)loc_cap));
    int32_t* starts = static_cast<int32_t*>(malloc(sizeof(int32_t) * (size_t)loc_cap));
    uint8_t* aln = static_cast<uint8_t*>(malloc((size_t)aln_cap));
    hsgpu_edlib_result r;
    memset(&r, 0, sizeof(r));
    const int rc = (ends && starts && aln)
                       ? hsgpu_edlib_align_batch(ctx, 1, query ? query : "", qo, target ? target : "", to, config.k, config.mode,
                                                 config.task, &r, ends, starts, loc_cap, aln, aln_cap)
                       : HSGPU_ERR_CUDA;
    if (rc != HSGPU_OK) {
        free(ends);
        free(starts);
        free(aln);
        return out;
    }
    out.status = r.status;
    out.editDistance = r.edit_distance;
    out.numLocations = r.n_locations;
    out.alphabetLength = r.alphabet_length;
    if (r.n_locations > 0) out.endLocations = ends; else free(ends);
    if (r.n_locations > 0 && r.has_start_locations) out.startLocations = starts; else free(starts);
    if (r.alignment_length > 0) {
        out.alignment = aln;
        out.alignmentLength = r.alignment_length;
    } else {
        free(aln);
    }
    return out;
}

extern "C" void hsgpu_edlibFreeAlignResult(hsgpu_EdlibAlignResult result) {
    free(result.endLocations);
    free(result.startLocations);
    free(result.alignment);
}
