// Read graph and chinese-whispers clustering of the windows of a batch of contigs:
//   create_read_graph_matrix      reference src/separate_reads.cpp:706-828
//   chinese_whispers_high_memory  reference src/cluster_graph.cpp:240-310
//
// The reference walks the windows of a contig one after the other; for every read that spans the
// window ("masked" read) it fills three n_reads-long vectors from the sparse similarity/difference
// matrices, sorts all n_reads distances and links the read to its closest neighbours; then it runs
// chinese whispers once per SNP of the window, each node scan allocating another n_reads-long vector.
//
// B200 formulation. Everything is local to the m masked reads of a window (links only ever join two
// masked reads, and labels only ever flow along links), so a window is an m x m problem with m ~ depth:
//   read_graph_kernel   one warp per (window, masked read): gathers the m similarity/difference counts
//                       of the read from the dense device-resident matrices hsgpu_pairs left in HBM,
//                       computes the m distances, the three order statistics the thresholds need by rank
//                       counting in shared memory (the reference's full sort is only ever read at its
//                       top), and writes the read's selected neighbours as one bit row.
//   graph_degree/fill   symmetrise (link i-j if either end selected the other) into a CSR over local
//                       indices -- Eigen's setFromTriplets + "set every value to 1".
//   window_order_kernel the node order of a sweep = the host's shuffled order of all reads restricted to
//                       the masked ones; one warp per (window, order) ranks the m reads.
//   whispers_kernel     one warp per clustering run (window x starting labels): nodes sequentially in
//                       sweep order (each update is visible to the next node, as in the reference), the
//                       neighbour-label vote of one node in parallel over the lanes: match_any + redux
//                       when the node has <= 32 neighbours, shared-memory counters otherwise.
// Exactness. All float expressions use round-to-nearest intrinsics in the reference's operation order (no
// FMA contraction). The only place where the reference's result depends on std::sort's treatment of equal
// keys is the "first five neighbours" rule when a group of equal distances straddles the fifth place; such
// reads are flagged by the kernel and resolved by replaying the reference's sort on the host with the
// same libstdc++ std::sort (hsgpu_graph_build reports how many).
#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"
#include "pairs_view.cuh"

#define GR_WARPS 8  // warps (= masked reads, runs) per CTA

struct hsgpu_graph {
    hsgpu_ctx* ctx = nullptr;
    hsgpu_pairs* pairs = nullptr;
    int32_t n_windows = 0, n_contigs = 0;
    int64_t total_masked = 0, sel_words = 0, n_adj = 0;
    int32_t max_m = 0;
    float error_rate = 0;
    int64_t n_flagged = 0;
    bool built = false;
    std::vector<int32_t> h_win_contig, h_win_reads, h_contig_n;
    std::vector<int64_t> h_win_off, h_sel_off;
    // device
    int32_t *d_win_contig = nullptr, *d_win_reads = nullptr, *d_read_win = nullptr;
    int64_t *d_win_off = nullptr, *d_sel_off = nullptr;
    HsPairView* d_views = nullptr;  // per contig: where its similarity / difference blocks are (pairs_view.cuh)
    uint8_t* d_win_low = nullptr;   // per window: 1 = the distance rule of create_read_graph_low_memory
    bool any_low = false;
    int32_t* d_contig_n = nullptr;
    uint32_t* d_sel = nullptr;   // per window m rows of ceil(m/32) words: bit j of row i = read i selected j
    uint8_t* d_flag = nullptr;   // per masked read: the selection depends on the order of equal distances
    uint32_t* d_deg = nullptr;
    int64_t* d_adj_off = nullptr;  // [total_masked+1]
    int32_t* d_adj = nullptr;      // neighbours as local indices, ascending
};

int hs_pairs_view(hsgpu_pairs* h, int32_t contig, hsgpu_ctx** ctx, HsPairView* view);
int hs_pairs_contigs(hsgpu_pairs* h);

// ---- distances of one masked read to the others of its window ----------------------------------------------
// dist[j] as the reference computes it: create_read_graph_matrix (:752-766) or, with `low`, the pairwise loop of
// create_read_graph_low_memory (:583-628), which has no `sims > 0` guard (a pair without a common SNP gives
// 1 - 0 / float(0) = NaN, a pair with differences only gets a distance too) and skips the reads that appear in no
// SNP column. Returns true when a NaN survives the overlap filter: the neighbour choice then hangs on what
// std::sort does with it, and the read is replayed on the host.
__device__ __forceinline__ bool gr_distances(const HsPairView& v, const int32_t* __restrict__ M, int m, int i, int lane,
                                             float* dist, bool low) {
    int max_compat = 0;
    const int ri = M[i];
    for (int j = lane; j < m; j += 32) {
        float ds = 0.f;
        if (j != i) {
            const int r = M[j];
            int s, d;
            hs_pair_get(v, ri, r, s, d);
            if (low ? (__ldg(v.has_cells + r) != 0) : (s > 0)) {
                const float diff = (float)max(0, d - 1);                       // :754 / :618
                ds = __fsub_rn(1.0f, __fdiv_rn(diff, (float)(s + d)));         // :755 / :618
                max_compat = max(max_compat, s);
            }
        }
        dist[j] = ds;
    }
    max_compat = __reduce_max_sync(0xffffffffu, max_compat);
    const double lim = 0.7 * (double)max_compat;  // :763 / :628, double arithmetic
    bool nan = false;
    for (int j = lane; j < m; j += 32) {
        if (j != i) {
            const int r = M[j];
            int s, d;
            hs_pair_get(v, ri, r, s, d);
            if ((double)(s + d) < lim) dist[j] = 0.f;
            nan |= dist[j] != dist[j];
        }
    }
    __syncwarp();
    return __any_sync(0xffffffffu, nan);
}

__device__ __forceinline__ int gr_warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// value at position idx (0-based) of the descending order of the R-long distance vector, of which only the m
// entries in dist can be non-zero
__device__ float gr_kth(const float* dist, int m, int idx, int lane) {
    float found = 0.f;
    int have = 0;
    for (int c0 = 0; c0 < m; c0 += 32) {
        const int j = c0 + lane;
        const float v = j < m ? dist[j] : 0.f;
        int gt = 0, ge = 0;
        if (j < m && v > 0.f) {
            for (int t = 0; t < m; t++) {
                const float u = dist[t];
                gt += u > v;
                ge += u >= v;
            }
        }
        const bool hit = j < m && v > 0.f && gt <= idx && idx < ge;
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b) {
            found = __shfl_sync(0xffffffffu, v, __ffs(b) - 1);
            have = 1;
            break;
        }
    }
    return have ? found : 0.f;  // beyond the positive entries lie the zeros
}

// MODE 0: selection bit rows + flags for every masked read. MODE 1: distance rows of the listed reads.
template <int MODE>
__global__ void __launch_bounds__(GR_WARPS * 32)
read_graph_kernel(int64_t n_items, const int64_t* __restrict__ items, const int64_t* __restrict__ item_off,
                  const int32_t* __restrict__ read_win, const int64_t* __restrict__ win_off,
                  const int32_t* __restrict__ win_reads, const int32_t* __restrict__ win_contig,
                  const HsPairView* __restrict__ views, const uint8_t* __restrict__ win_low,
                  const int32_t* __restrict__ contig_n, const int64_t* __restrict__ sel_off, float error_rate, int max_m, uint32_t* __restrict__ sel,
                  uint8_t* __restrict__ flag, float* __restrict__ rows) {
    extern __shared__ float gr_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * GR_WARPS + warp;
    if (item >= n_items) return;
    const int64_t g = MODE == 0 ? item : items[item];
    float* dist = gr_smem + (size_t)warp * max_m;
    const int w = read_win[g];
    const int64_t g0 = win_off[w];
    const int m = (int)(win_off[w + 1] - g0), i = (int)(g - g0);
    const int32_t* M = win_reads + g0;
    const int c = win_contig[w];
    const int R = contig_n[c];
    const HsPairView view = views[c];
    const bool low = win_low != nullptr && win_low[w] != 0;
    if (low && __ldg(view.has_cells + M[i]) == 0) {
        // create_read_graph_low_memory skips a read that appears in no SNP column (mask_extend, :569-576)
        if (MODE == 0) {
            uint32_t* out0 = sel + sel_off[w] + (int64_t)i * ((m + 31) >> 5);
            for (int k = lane; k < ((m + 31) >> 5); k += 32) out0[k] = 0;
            if (lane == 0) flag[g] = 0;
        }
        return;
    }
    const bool has_nan = gr_distances(view, M, m, i, lane, dist, low);
    if (MODE == 1) {
        float* out = rows + item_off[item];
        for (int j = lane; j < m; j += 32) out[j] = dist[j];
        return;
    }
    // ---- thresholds (:778-795) ----
    const float below = __fsub_rn(1.0f, __fmul_rn(error_rate, 2.0f));
    float s0 = 0.f;
    for (int j = lane; j < m; j += 32) s0 = fmaxf(s0, dist[j]);
    s0 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(s0)));  // all values >= 0
    int n_top = 0, n_one = 0, n_pos = 0;
    float s1 = 0.f;
    for (int j = lane; j < m; j += 32) {
        const float v = dist[j];
        n_top += v == s0;
        n_one += v == 1.0f;
        n_pos += v > 0.f;
        if (v < s0) s1 = fmaxf(s1, v);
    }
    n_top = gr_warp_sum(n_top);
    n_one = gr_warp_sum(n_one);
    n_pos = gr_warp_sum(n_pos);
    s1 = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(s1)));
    if (s0 == 0.f) n_top = R;  // every entry of the vector is a zero
    if (n_top >= 2) s1 = s0;
    float above = 1.0f;
    if (R > 1) above = __fsub_rn(s0, __fmul_rn(__fsub_rn(s0, s1), 3.0f));
    if (above == 1.0f) {
        int idx = n_one;
        if (idx < R) {
            idx = min(idx + 4, R - 1);
            above = idx < n_pos ? gr_kth(dist, m, idx, lane) : 0.f;
        }
    }
    // ---- selection (:808-817): entries >= above (or == 1) always link, the rest while fewer than 5 are linked ----
    bool ambiguous = below < 0.f || has_nan;  // zeros would qualify: order among all n_reads zeros matters -> host
    int n_a = 0, n_rest = 0;
    for (int j = lane; j < m; j += 32) {
        const float v = dist[j];
        const bool linkable = v > below;
        const bool in_a = v >= above || v == 1.0f;
        n_a += linkable && in_a;
        n_rest += linkable && !in_a;
    }
    n_a = gr_warp_sum(n_a);
    n_rest = gr_warp_sum(n_rest);
    float cut = above;  // link everything linkable with dist >= cut
    if (n_a < 5 && n_rest > 0) {
        const int need = 5 - n_a;
        if (n_rest <= need) {
            cut = -1.f;
        } else {
            // the need-th largest of the rest
            float vcut = 0.f;
            int ge_cut = 0, got = 0;
            for (int c0 = 0; c0 < m && !got; c0 += 32) {
                const int j = c0 + lane;
                const float v = j < m ? dist[j] : 0.f;
                const bool cand = j < m && v > below && !(v >= above || v == 1.0f);
                int gt = 0, ge = 0;
                if (cand) {
                    for (int t = 0; t < m; t++) {
                        const float u = dist[t];
                        const bool rest = u > below && !(u >= above || u == 1.0f);
                        gt += rest && u > v;
                        ge += rest && u >= v;
                    }
                }
                const bool hit = cand && gt < need && need <= ge;
                const unsigned b = __ballot_sync(0xffffffffu, hit);
                if (b) {
                    const int src = __ffs(b) - 1;
                    vcut = __shfl_sync(0xffffffffu, v, src);
                    ge_cut = __shfl_sync(0xffffffffu, ge, src);
                    got = 1;
                }
            }
            cut = vcut;
            if (ge_cut != need) ambiguous = true;
        }
    }
    const int words = (m + 31) >> 5;
    uint32_t* out = sel + sel_off[w] + (int64_t)i * words;
    for (int c0 = 0; c0 < m; c0 += 32) {
        const int j = c0 + lane;
        bool on = false;
        if (j < m) {
            const float v = dist[j];
            on = v > below && (v >= above || v == 1.0f || v >= cut);
        }
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (lane == 0) out[c0 >> 5] = b;
    }
    if (lane == 0) flag[g] = ambiguous ? 1 : 0;
}

// ---- symmetrise the selections into a CSR ------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(GR_WARPS * 32)
graph_csr_kernel(int64_t total, const int32_t* __restrict__ read_win, const int64_t* __restrict__ win_off,
                 const int64_t* __restrict__ sel_off, const uint32_t* __restrict__ sel, uint32_t* __restrict__ deg,
                 const int64_t* __restrict__ adj_off, int32_t* __restrict__ adj) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * GR_WARPS + warp;
    if (g >= total) return;
    const int w = read_win[g];
    const int64_t g0 = win_off[w];
    const int m = (int)(win_off[w + 1] - g0), i = (int)(g - g0);
    const int words = (m + 31) >> 5;
    const uint32_t* S = sel + sel_off[w];
    int64_t pos = FILL ? adj_off[g] : 0;
    int count = 0;
    for (int c0 = 0; c0 < m; c0 += 32) {
        const int j = c0 + lane;
        const bool t = j < m && ((S[(int64_t)j * words + (i >> 5)] >> (i & 31)) & 1u);
        const unsigned both = __ballot_sync(0xffffffffu, t) | S[(int64_t)i * words + (c0 >> 5)];
        if (FILL) {
            if ((both >> lane) & 1u) adj[pos + __popc(both & ((1u << lane) - 1u))] = j;
            pos += __popc(both);
        } else {
            count += __popc(both);
        }
    }
    if (!FILL && lane == 0) deg[g] = (uint32_t)count;
}

// ---- sweep orders ---------------------------------------------------------------------------------------------
// order[(win_off[w] * n_orders) + k * m + t] = local index of the t-th masked read of window w in sweep order k,
// given rank[order_base[c] + k * n + r] = position of read r of contig c in the host's k-th shuffled order
__global__ void __launch_bounds__(GR_WARPS * 32)
window_order_kernel(int64_t n_items, int n_orders, const int64_t* __restrict__ win_off,
                    const int32_t* __restrict__ win_reads, const int32_t* __restrict__ win_contig,
                    const int32_t* __restrict__ contig_n, const int64_t* __restrict__ order_base,
                    const int32_t* __restrict__ rank, int32_t* __restrict__ order) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * GR_WARPS + warp;
    if (item >= n_items) return;
    const int w = (int)(item / n_orders), k = (int)(item % n_orders);
    const int64_t g0 = win_off[w];
    const int m = (int)(win_off[w + 1] - g0);
    const int c = win_contig[w];
    const int32_t* rk = rank + order_base[c] + (int64_t)k * contig_n[c];
    const int32_t* M = win_reads + g0;
    int32_t* out = order + g0 * n_orders + (int64_t)k * m;
    for (int i = lane; i < m; i += 32) {
        const int mine = __ldg(rk + M[i]);
        int before = 0;
        for (int t = 0; t < m; t++) before += __ldg(rk + M[t]) < mine;
        out[before] = i;
    }
}

// ---- chinese whispers -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GR_WARPS * 32)
whispers_kernel(int64_t n_runs, const int32_t* __restrict__ run_window, const int64_t* __restrict__ run_off,
                const int32_t* __restrict__ init, const int64_t* __restrict__ win_off,
                const int64_t* __restrict__ adj_off, const int32_t* __restrict__ adj, int n_orders,
                const int32_t* __restrict__ order, int max_m, int32_t* __restrict__ labels_out) {
    extern __shared__ int32_t cw_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t run = (int64_t)blockIdx.x * GR_WARPS + warp;
    if (run >= n_runs) return;
    int32_t* lab = cw_smem + (size_t)warp * 2 * max_m;
    int32_t* cnt = lab + max_m;
    const int w = run_window[run];
    const int64_t g0 = win_off[w];
    const int m = (int)(win_off[w + 1] - g0);
    const int32_t* in = init + run_off[run];
    for (int j = lane; j < m; j += 32) {
        lab[j] = in[j];
        cnt[j] = 0;
    }
    __syncwarp();
    int changes = 3, sweeps = 0;
    while (changes > 2 && sweeps < 15) {  // cluster_graph.cpp:248
        changes = 0;
        const int32_t* ord = order + g0 * n_orders + (int64_t)min(sweeps, n_orders - 1) * m;
        for (int t = 0; t < m; t++) {
            const int i = __ldg(ord + t);
            const int64_t e0 = __ldg(adj_off + g0 + i), e1 = __ldg(adj_off + g0 + i + 1);
            const int deg = (int)(e1 - e0);
            if (deg == 0) continue;
            unsigned best;
            if (deg <= 32) {
                const int l = lane < deg ? lab[__ldg(adj + e0 + lane)] : -1;
                const unsigned peers = __match_any_sync(0xffffffffu, l);
                const unsigned key = l >= 0 ? ((unsigned)__popc(peers) << 16) | (unsigned)(0xffff - l) : 0u;
                best = __reduce_max_sync(0xffffffffu, key);
            } else {
                for (int e = lane; e < deg; e += 32) {
                    const int l = lab[__ldg(adj + e0 + e)];
                    if (l >= 0) atomicAdd(&cnt[l], 1);
                }
                __syncwarp();
                unsigned key = 0;
                for (int e = lane; e < deg; e += 32) {
                    const int l = lab[__ldg(adj + e0 + e)];
                    if (l >= 0) key = max(key, ((unsigned)cnt[l] << 16) | (unsigned)(0xffff - l));
                }
                best = __reduce_max_sync(0xffffffffu, key);
                __syncwarp();
                for (int e = lane; e < deg; e += 32) {
                    const int l = lab[__ldg(adj + e0 + e)];
                    if (l >= 0) cnt[l] = 0;
                }
            }
            if (best >> 16) {  // max_value > 0: lowest label among the most frequent (:272-287)
                const int pick = 0xffff - (int)(best & 0xffffu);
                if (lab[i] != pick) changes++;
                __syncwarp();
                if (lane == 0) lab[i] = pick;
            }
            __syncwarp();
        }
        sweeps++;
    }
    int32_t* out = labels_out + run_off[run];
    for (int j = lane; j < m; j += 32) out[j] = lab[j];
}

// ---- host replay of the reference's neighbour selection for one read (:768-817) ------------------------------
static void replay_selection(int R, int m, const int32_t* M, const float* dist, float error_rate, uint32_t* sel_row) {
    std::vector<std::pair<int, float>> smallest((size_t)R);
    for (int r = 0; r < R; r++) smallest[r] = std::make_pair(r, 0.0f);
    std::vector<int> local((size_t)R, -1);
    for (int j = 0; j < m; j++) {
        smallest[M[j]].second = dist[j];
        local[M[j]] = j;
    }
    std::sort(smallest.begin(), smallest.end(),
              [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
    int nb = 0;
    const float below = 1 - error_rate * 2;
    float above = 1;
    if (smallest.size() > 1) above = smallest[0].second - (smallest[0].second - smallest[1].second) * 3;
    if (above == 1) {
        int idx = 0;
        while (idx < (int)smallest.size() && smallest[idx].second == 1) idx += 1;
        if (idx < (int)smallest.size()) {
            idx = std::min(idx + 4, (int)smallest.size() - 1);
            above = smallest[idx].second;
        }
    }
    const int words = (m + 31) >> 5;
    for (int k = 0; k < words; k++) sel_row[k] = 0;
    for (const auto& nbr : smallest) {
        if (nbr.second > below && (nb < 5 || nbr.second == 1 || nbr.second >= above) && local[nbr.first] >= 0) {
            nb++;
            const int j = local[nbr.first];
            sel_row[j >> 5] |= 1u << (j & 31);
        }
    }
}

static void graph_release(hsgpu_graph* g) {
    hsgpu_ctx* ctx = g->ctx;
    hs_free(ctx, g->d_win_contig); hs_free(ctx, g->d_win_reads); hs_free(ctx, g->d_read_win);
    hs_free(ctx, g->d_win_off); hs_free(ctx, g->d_sel_off); hs_free(ctx, g->d_views); hs_free(ctx, g->d_win_low);
    hs_free(ctx, g->d_contig_n); hs_free(ctx, g->d_sel); hs_free(ctx, g->d_flag);
    hs_free(ctx, g->d_deg); hs_free(ctx, g->d_adj_off); hs_free(ctx, g->d_adj);
}

#define GR_TRY(call)                                                          \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) {                                              \
            graph_release(g);                                                 \
            delete g;                                                         \
            return hs_cuda_fail(ctx, _e, #call, __FILE__, __LINE__);          \
        }                                                                     \
    } while (0)

extern "C" {

int hsgpu_graph_create(hsgpu_pairs* pairs, int32_t n_windows, const int32_t* win_contig, const int64_t* win_off,
                       const int32_t* win_reads, float error_rate, hsgpu_graph** out) {
    return hsgpu_graph_create_ex(pairs, n_windows, win_contig, win_off, win_reads, nullptr, error_rate, out);
}

int hsgpu_graph_create_ex(hsgpu_pairs* pairs, int32_t n_windows, const int32_t* win_contig, const int64_t* win_off,
                          const int32_t* win_reads, const uint8_t* win_low_memory, float error_rate, hsgpu_graph** out) {
    if (!pairs || !out || n_windows < 0 || (n_windows > 0 && (!win_contig || !win_off))) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = nullptr;
    const int n_contigs = hs_pairs_contigs(pairs);
    {
        HsPairView v0;
        if (n_contigs <= 0 || hs_pairs_view(pairs, 0, &ctx, &v0) != HSGPU_OK) return HSGPU_ERR_ARG;
    }
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    hsgpu_graph* g = new hsgpu_graph;
    g->ctx = ctx;
    g->pairs = pairs;
    g->n_windows = n_windows;
    g->n_contigs = n_contigs;
    g->error_rate = error_rate;
    g->total_masked = n_windows ? win_off[n_windows] : 0;
    if (g->total_masked > 0 && !win_reads) { delete g; return HSGPU_ERR_ARG; }
    std::vector<HsPairView> views(n_contigs);
    std::vector<int32_t> cn(n_contigs);
    for (int c = 0; c < n_contigs; c++) {
        hsgpu_ctx* cx;
        if (hs_pairs_view(pairs, c, &cx, &views[c]) != HSGPU_OK) {
            delete g;
            HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_graph_create: call hsgpu_pairs_compute first");
        }
        cn[c] = views[c].n;
    }
    g->h_contig_n = cn;
    g->h_win_contig.assign(win_contig, win_contig + n_windows);
    g->h_win_off.assign(win_off, win_off + n_windows + (n_windows ? 1 : 0));
    if (n_windows == 0) g->h_win_off.assign(1, 0);
    g->h_win_reads.assign(win_reads, win_reads + g->total_masked);
    g->h_sel_off.assign((size_t)n_windows + 1, 0);
    std::vector<int32_t> read_win((size_t)g->total_masked);
    for (int w = 0; w < n_windows; w++) {
        const int64_t m = win_off[w + 1] - win_off[w];
        const int c = win_contig[w];
        if (m < 0 || c < 0 || c >= n_contigs) { delete g; HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_graph_create: bad window"); }
        if (m > 65535) { delete g; HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_graph_create: more than 65535 reads span one window"); }
        for (int64_t j = 0; j < m; j++) {
            const int32_t r = win_reads[win_off[w] + j];
            if (r < 0 || r >= cn[c] || (j > 0 && r <= win_reads[win_off[w] + j - 1])) {
                delete g;
                HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_graph_create: window reads must be ascending read indices of the contig");
            }
            read_win[win_off[w] + j] = w;
        }
        g->max_m = std::max<int32_t>(g->max_m, (int32_t)m);
        g->h_sel_off[w + 1] = g->h_sel_off[w] + m * ((m + 31) / 32);
    }
    g->sel_words = g->h_sel_off[n_windows];
    // read_graph_kernel keeps one distance row per warp in shared memory (whispers_kernel: two label arrays, checked
    // in hsgpu_graph_whispers)
    if ((size_t)GR_WARPS * g->max_m * sizeof(float) > 200 * 1024) {
        delete g;
        HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_graph_create: more than 6400 reads span one window");
    }
    GR_TRY(hs_alloc(ctx, &g->d_win_contig, n_windows));
    GR_TRY(hs_alloc(ctx, &g->d_win_off, n_windows + 1));
    GR_TRY(hs_alloc(ctx, &g->d_sel_off, n_windows + 1));
    GR_TRY(hs_alloc(ctx, &g->d_win_reads, g->total_masked));
    GR_TRY(hs_alloc(ctx, &g->d_read_win, g->total_masked));
    GR_TRY(hs_alloc(ctx, &g->d_views, n_contigs));
    GR_TRY(hs_alloc(ctx, &g->d_contig_n, n_contigs));
    if (win_low_memory) {
        GR_TRY(hs_alloc(ctx, &g->d_win_low, n_windows));
        GR_TRY(hs_h2d(ctx, g->d_win_low, win_low_memory, n_windows));
        for (int w = 0; w < n_windows; w++) g->any_low = g->any_low || win_low_memory[w];
    }
    GR_TRY(hs_alloc(ctx, &g->d_sel, g->sel_words));
    GR_TRY(hs_alloc(ctx, &g->d_flag, g->total_masked));
    GR_TRY(hs_alloc(ctx, &g->d_deg, g->total_masked));
    GR_TRY(hs_alloc(ctx, &g->d_adj_off, g->total_masked + 1));
    GR_TRY(hs_h2d(ctx, g->d_win_contig, g->h_win_contig.data(), n_windows));
    GR_TRY(hs_h2d(ctx, g->d_win_off, g->h_win_off.data(), n_windows + 1));
    GR_TRY(hs_h2d(ctx, g->d_sel_off, g->h_sel_off.data(), n_windows + 1));
    GR_TRY(hs_h2d(ctx, g->d_win_reads, g->h_win_reads.data(), g->total_masked));
    GR_TRY(hs_h2d(ctx, g->d_read_win, read_win.data(), g->total_masked));
    GR_TRY(hs_h2d(ctx, g->d_views, views.data(), n_contigs));
    GR_TRY(hs_h2d(ctx, g->d_contig_n, cn.data(), n_contigs));
    GR_TRY(hs_stream_sync(ctx));  // the staging vectors go out of scope
    *out = g;
    return HSGPU_OK;
}

void hsgpu_graph_destroy(hsgpu_graph* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    graph_release(g);
    delete g;
}

int hsgpu_graph_build(hsgpu_graph* g, int64_t* n_replayed) {
    if (!g) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = g->ctx;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_replayed) *n_replayed = 0;
    const int64_t total = g->total_masked;
    if (total == 0) {
        int64_t zero = 0;
        HS_CUDA(ctx, hs_h2d(ctx, g->d_adj_off, &zero, 1));
        g->built = true;
        return HSGPU_OK;
    }
    const unsigned blocks = (unsigned)((total + GR_WARPS - 1) / GR_WARPS);
    const size_t smem = (size_t)GR_WARPS * g->max_m * sizeof(float);
    if (smem > 48 * 1024) {
        HS_CUDA(ctx, cudaFuncSetAttribute(read_graph_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HS_CUDA(ctx, cudaFuncSetAttribute(read_graph_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    HS_KERNEL(ctx, "read_graph_kernel",
              read_graph_kernel<0><<<blocks, GR_WARPS * 32, smem, ctx->stream>>>(
                  total, nullptr, nullptr, g->d_read_win, g->d_win_off, g->d_win_reads, g->d_win_contig, g->d_views, g->d_win_low,
                  g->d_contig_n, g->d_sel_off, g->error_rate, g->max_m, g->d_sel, g->d_flag, nullptr));
    // reads whose selection hangs on std::sort's order of equal distances: replay the reference on the host
    std::vector<uint8_t> flag((size_t)total);
    HS_CUDA(ctx, hs_d2h(ctx, flag.data(), g->d_flag, total));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    std::vector<int64_t> items, item_off;
    int64_t n_floats = 0;
    {
        int w = 0;
        for (int64_t i = 0; i < total; i++) {
            if (!flag[i]) continue;
            while (g->h_win_off[w + 1] <= i) w++;
            items.push_back(i);
            item_off.push_back(n_floats);
            n_floats += g->h_win_off[w + 1] - g->h_win_off[w];
        }
    }
    g->n_flagged = (int64_t)items.size();
    if (n_replayed) *n_replayed = g->n_flagged;
    if (!items.empty()) {
        int64_t *d_items = nullptr, *d_item_off = nullptr;
        float* d_rows = nullptr;
        HsTemps temps(ctx);
        temps.own(d_items, d_item_off, d_rows);
        const int64_t ni = (int64_t)items.size();
        HS_CUDA(ctx, hs_alloc(ctx, &d_items, ni));
        HS_CUDA(ctx, hs_alloc(ctx, &d_item_off, ni));
        HS_CUDA(ctx, hs_alloc(ctx, &d_rows, n_floats));
        HS_CUDA(ctx, hs_h2d(ctx, d_items, items.data(), ni));
        HS_CUDA(ctx, hs_h2d(ctx, d_item_off, item_off.data(), ni));
        HS_KERNEL(ctx, "read_graph_rows_kernel",
                  read_graph_kernel<1><<<(unsigned)((ni + GR_WARPS - 1) / GR_WARPS), GR_WARPS * 32, smem, ctx->stream>>>(
                      ni, d_items, d_item_off, g->d_read_win, g->d_win_off, g->d_win_reads, g->d_win_contig, g->d_views,
                      g->d_win_low, g->d_contig_n, g->d_sel_off, g->error_rate, g->max_m, g->d_sel,
                      g->d_flag, d_rows));
        std::vector<float> rows((size_t)n_floats);
        HS_CUDA(ctx, hs_d2h(ctx, rows.data(), d_rows, n_floats));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        // patch the selection rows of the replayed reads in a host copy of the bit matrices
        std::vector<uint32_t> sel((size_t)g->sel_words);
        HS_CUDA(ctx, hs_d2h(ctx, sel.data(), g->d_sel, g->sel_words));
        HS_CUDA(ctx, hs_stream_sync(ctx));
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t t = 0; t < ni; t++) {
            const int64_t gi = items[t];
            const int w = (int)(std::upper_bound(g->h_win_off.begin(), g->h_win_off.end(), gi) - g->h_win_off.begin()) - 1;
            const int64_t g0 = g->h_win_off[w];
            const int m = (int)(g->h_win_off[w + 1] - g0);
            replay_selection(g->h_contig_n[g->h_win_contig[w]], m, g->h_win_reads.data() + g0, rows.data() + item_off[t],
                             g->error_rate, sel.data() + g->h_sel_off[w] + (gi - g0) * ((m + 31) / 32));
        }
        HS_CUDA(ctx, hs_h2d(ctx, g->d_sel, sel.data(), g->sel_words));
        HS_CUDA(ctx, hs_stream_sync(ctx));
        hs_free(ctx, d_items); hs_free(ctx, d_item_off); hs_free(ctx, d_rows);
    }
    HS_KERNEL(ctx, "graph_degree_kernel",
              graph_csr_kernel<false><<<blocks, GR_WARPS * 32, 0, ctx->stream>>>(total, g->d_read_win, g->d_win_off, g->d_sel_off,
                                                                                  g->d_sel, g->d_deg, nullptr, nullptr));
    int64_t* d_total = nullptr;
    HsTemps temps(ctx);
    temps.own(d_total);
    HS_CUDA(ctx, hs_alloc(ctx, &d_total, 1));
    int rc = hs_exclusive_scan_u32_to_i64(ctx, g->d_deg, g->d_adj_off, total, d_total);
    if (rc) return rc;
    HS_CUDA(ctx, cudaMemcpyAsync(g->d_adj_off + total, d_total, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    HS_CUDA(ctx, hs_d2h(ctx, &g->n_adj, d_total, 1));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    hs_free(ctx, d_total);
    hs_free(ctx, g->d_adj);
    HS_CUDA(ctx, hs_alloc(ctx, &g->d_adj, g->n_adj));
    HS_KERNEL(ctx, "graph_fill_kernel",
              graph_csr_kernel<true><<<blocks, GR_WARPS * 32, 0, ctx->stream>>>(total, g->d_read_win, g->d_win_off, g->d_sel_off,
                                                                                 g->d_sel, g->d_deg, g->d_adj_off, g->d_adj));
    g->built = true;
    return HSGPU_OK;
}

int hsgpu_graph_adjacency(hsgpu_graph* g, int64_t* adj_off, int64_t capacity, int32_t* adj, int64_t* n_adj) {
    if (!g) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = g->ctx;
    if (!g->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_graph_adjacency: call hsgpu_graph_build first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_adj) *n_adj = g->n_adj;
    if (adj_off) HS_CUDA(ctx, hs_d2h(ctx, adj_off, g->d_adj_off, g->total_masked + 1));
    const bool fits = capacity >= g->n_adj;
    if (adj && fits) HS_CUDA(ctx, hs_d2h(ctx, adj, g->d_adj, g->n_adj));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    return (adj && !fits) ? HSGPU_ERR_CAPACITY : HSGPU_OK;
}

int hsgpu_graph_whispers(hsgpu_graph* g, int64_t n_runs, const int32_t* run_window, const int32_t* init_labels,
                         int32_t n_orders, const int32_t* order_rank, int32_t* labels_out) {
    if (!g || n_runs < 0 || n_orders < 1 || (n_runs > 0 && (!run_window || !init_labels || !order_rank || !labels_out)))
        return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = g->ctx;
    if (!g->built) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_graph_whispers: call hsgpu_graph_build first");
    if (n_runs == 0) return HSGPU_OK;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int64_t> run_off((size_t)n_runs + 1, 0);
    int32_t run_max_m = 1;  // the largest window a run works on sizes the shared-memory label arrays
    for (int64_t i = 0; i < n_runs; i++) {
        const int w = run_window[i];
        if (w < 0 || w >= g->n_windows) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_graph_whispers: bad window index");
        run_off[i + 1] = run_off[i] + (g->h_win_off[w + 1] - g->h_win_off[w]);
        run_max_m = std::max<int32_t>(run_max_m, (int32_t)(g->h_win_off[w + 1] - g->h_win_off[w]));
    }
    if ((size_t)GR_WARPS * 2 * run_max_m * sizeof(int32_t) > 200 * 1024)
        HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_graph_whispers: a run works on a window of more than 3200 reads (shared-memory "
                                      "label arrays); cluster such windows from hsgpu_graph_adjacency on the host");
    std::vector<int64_t> order_base((size_t)g->n_contigs + 1, 0);
    for (int c = 0; c < g->n_contigs; c++) order_base[c + 1] = order_base[c] + (int64_t)n_orders * g->h_contig_n[c];
    const int64_t n_lab = run_off[n_runs];
    int32_t *d_run_window = nullptr, *d_init = nullptr, *d_rank = nullptr, *d_order = nullptr, *d_out = nullptr;
    int64_t *d_run_off = nullptr, *d_order_base = nullptr;
    HsTemps temps(ctx);
    temps.own(d_run_window, d_init, d_rank, d_order, d_out, d_run_off, d_order_base);
    HS_CUDA(ctx, hs_alloc(ctx, &d_run_window, n_runs));
    HS_CUDA(ctx, hs_alloc(ctx, &d_run_off, n_runs + 1));
    HS_CUDA(ctx, hs_alloc(ctx, &d_init, n_lab));
    HS_CUDA(ctx, hs_alloc(ctx, &d_out, n_lab));
    HS_CUDA(ctx, hs_alloc(ctx, &d_rank, order_base[g->n_contigs]));
    HS_CUDA(ctx, hs_alloc(ctx, &d_order_base, g->n_contigs + 1));
    HS_CUDA(ctx, hs_alloc(ctx, &d_order, g->total_masked * n_orders));
    HS_CUDA(ctx, hs_h2d(ctx, d_run_window, run_window, n_runs));
    HS_CUDA(ctx, hs_h2d(ctx, d_run_off, run_off.data(), n_runs + 1));
    HS_CUDA(ctx, hs_h2d(ctx, d_init, init_labels, n_lab));
    HS_CUDA(ctx, hs_h2d(ctx, d_rank, order_rank, order_base[g->n_contigs]));
    HS_CUDA(ctx, hs_h2d(ctx, d_order_base, order_base.data(), g->n_contigs + 1));
    const int64_t n_items = (int64_t)g->n_windows * n_orders;
    HS_KERNEL(ctx, "window_order_kernel",
              window_order_kernel<<<(unsigned)((n_items + GR_WARPS - 1) / GR_WARPS), GR_WARPS * 32, 0, ctx->stream>>>(
                  n_items, n_orders, g->d_win_off, g->d_win_reads, g->d_win_contig, g->d_contig_n, d_order_base, d_rank, d_order));
    const size_t smem = (size_t)GR_WARPS * 2 * run_max_m * sizeof(int32_t);
    if (smem > 48 * 1024)
        HS_CUDA(ctx, cudaFuncSetAttribute(whispers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HS_KERNEL(ctx, "whispers_kernel",
              whispers_kernel<<<(unsigned)((n_runs + GR_WARPS - 1) / GR_WARPS), GR_WARPS * 32, smem, ctx->stream>>>(
                  n_runs, d_run_window, d_run_off, d_init, g->d_win_off, g->d_adj_off, g->d_adj, n_orders, d_order, run_max_m, d_out));
    HS_CUDA(ctx, hs_d2h(ctx, labels_out, d_out, n_lab));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    hs_free(ctx, d_run_window); hs_free(ctx, d_run_off); hs_free(ctx, d_init); hs_free(ctx, d_out);
    hs_free(ctx, d_rank); hs_free(ctx, d_order_base); hs_free(ctx, d_order);
    return HSGPU_OK;
}

}  // extern "C"
