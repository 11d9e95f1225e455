// HS_separate_reads -- drop-in for the reference executable of the same name (src/separate_reads.cpp:1398-1790,
// called by hairsplitter.py with 9 positional arguments). The .col parser, the window walk and the sequential
// post-processing of each window (finalize_clustering) run on the host; the read x read agreement counts
// (list_similarities_and_differences_between_reads3), the read graph of every window (create_read_graph_matrix)
// and the chinese-whispers runs started from every SNP (chinese_whispers_high_memory) run on the GPU through the
// C ABI of libhsgpu (hsgpu_pairs_*, hsgpu_graph_*), all windows of all contigs of a shard in one batch. The
// neighbour lists of the low-memory path (create_read_graph_low_memory: -l, amplicons, > 1000x) come from the same
// counts on the GPU; those contigs are clustered on the host like in the reference. Contigs are sharded over the
// visible GPUs (HSGPU_NGPUS, HSGPU_DEVICE), heaviest first.
#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hsgpu.h"
#include "hs_sepreads.h"

using namespace hs;

static double g_t0 = 0;
static bool g_timing = false;
static void phase(const char* name) {
    if (!g_timing) return;
    const double t = omp_get_wtime();
    if (g_t0 > 0) fprintf(stderr, "[hs timing] %-28s %8.3f s\n", name, t - g_t0);
    g_t0 = t;
}

#define GPU_CHECK(ctx, call)                                                                        \
    do {                                                                                            \
        const int _rc = (call);                                                                     \
        if (_rc != HSGPU_OK) {                                                                      \
            std::cout << "ERROR: " #call " failed (" << _rc << "): " << hsgpu_last_error(ctx) << std::endl; \
            std::exit(1);                                                                           \
        }                                                                                           \
    } while (0)

// windows above this size are clustered on the host from the device-built graph (hsgpu_graph_whispers keeps two label
// arrays per run in shared memory)
static const size_t kMaxWhispersWindow = 3200;

// one GPU's share: read x read counts, read graphs and SNP-started clusterings of `jobs`. High-memory contigs get
// create_read_graph_matrix + the clustering runs; low-memory contigs (job->low_now) get the neighbour lists of
// create_read_graph_low_memory and are clustered by the caller.
static void gpu_shard(hsgpu_ctx* ctx, const std::vector<ColContig>& contigs, std::vector<ContigJob*>& jobs, float error_rate,
                      Shuffler& sh, int64_t* stats) {
    if (jobs.empty()) return;
    const int nc = (int)jobs.size();
    // ---- SNP columns of the shard, concatenated ----
    std::vector<int32_t> n_reads(nc);
    std::vector<int64_t> snp_base(nc + 1, 0);
    int64_t n_cells = 0;
    for (int j = 0; j < nc; j++) {
        const ColContig& c = contigs[jobs[j]->n];
        n_reads[j] = (int32_t)c.read_lines.size();
        snp_base[j + 1] = snp_base[j] + (int64_t)c.snps.size();
        for (const Column& s : c.snps) n_cells += (int64_t)s.readIdxs.size();
    }
    const int64_t total_snps = snp_base[nc];
    std::vector<int64_t> snp_off((size_t)total_snps + 1, 0);
    std::vector<uint32_t> read_idx((size_t)n_cells);
    std::vector<uint8_t> code((size_t)n_cells), rb((size_t)total_snps), sb((size_t)total_snps);
    {
        int64_t s = 0;
        for (int j = 0; j < nc; j++)
            for (const Column& col : contigs[jobs[j]->n].snps) {
                snp_off[s + 1] = snp_off[s] + (int64_t)col.readIdxs.size();
                s++;
            }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < nc; j++) {
        int64_t s = snp_base[j];
        for (const Column& col : contigs[jobs[j]->n].snps) {
            std::copy(col.readIdxs.begin(), col.readIdxs.end(), read_idx.begin() + snp_off[s]);
            std::copy(col.content.begin(), col.content.end(), code.begin() + snp_off[s]);
            rb[s] = col.ref_base;
            sb[s] = col.second_base;
            s++;
        }
    }
    hsgpu_pairs* pairs = nullptr;
    GPU_CHECK(ctx, hsgpu_pairs_create(ctx, nc, n_reads.data(), snp_base.data(), snp_off.data(), read_idx.data(), code.data(),
                                      rb.data(), sb.data(), 0, &pairs));
    GPU_CHECK(ctx, hsgpu_pairs_compute(pairs));
    phase("  gpu: read pair counts");

    // ---- windows with SNPs ----
    std::vector<int32_t> win_contig, win_reads;
    std::vector<int64_t> win_off(1, 0);
    std::vector<std::pair<int, int>> win_ref;  // (job, window index)
    std::vector<uint8_t> win_low;
    for (int j = 0; j < nc; j++) {
        for (size_t w = 0; w < jobs[j]->windows.size(); w++) {
            const Window& win = jobs[j]->windows[w];
            if (!win.has_snps) continue;
            win_contig.push_back(j);
            win_low.push_back(jobs[j]->low_now ? 1 : 0);
            win_reads.insert(win_reads.end(), win.masked.begin(), win.masked.end());
            win_off.push_back((int64_t)win_reads.size());
            win_ref.emplace_back(j, (int)w);
        }
    }
    const int n_windows = (int)win_contig.size();
    hsgpu_graph* graph = nullptr;
    GPU_CHECK(ctx, hsgpu_graph_create_ex(pairs, n_windows, win_contig.data(), win_off.data(), win_reads.data(), win_low.data(),
                                         error_rate, &graph));
    int64_t replayed = 0;
    GPU_CHECK(ctx, hsgpu_graph_build(graph, &replayed));
    std::vector<int64_t> adj_off((size_t)win_off.back() + 1, 0);
    int64_t n_adj = 0;
    GPU_CHECK(ctx, hsgpu_graph_adjacency(graph, adj_off.data(), 0, nullptr, &n_adj));
    std::vector<int32_t> adj((size_t)std::max<int64_t>(n_adj, 1));
    GPU_CHECK(ctx, hsgpu_graph_adjacency(graph, nullptr, n_adj, adj.data(), &n_adj));
    phase("  gpu: read graphs");

    // ---- clustering runs: one per (window, restart SNP) ----
    std::vector<int64_t> run_base((size_t)n_windows + 1, 0);  // first run of each window
    auto on_device = [&](int w) {  // does the device cluster this window?
        return !win_low[w] && (size_t)(win_off[w + 1] - win_off[w]) <= kMaxWhispersWindow;
    };
    for (int w = 0; w < n_windows; w++)
        run_base[w + 1] = run_base[w] +
                          (on_device(w) ? (int64_t)jobs[win_ref[w].first]->windows[win_ref[w].second].restart_snps.size() : 0);
    const int64_t n_runs = run_base[n_windows];
    std::vector<int32_t> run_window((size_t)n_runs);
    std::vector<int64_t> run_off((size_t)n_runs + 1, 0);
    for (int w = 0; w < n_windows; w++)
        for (int64_t r = run_base[w]; r < run_base[w + 1]; r++) {
            run_window[r] = w;
            run_off[r + 1] = run_off[r] + (win_off[w + 1] - win_off[w]);
        }
    std::vector<int32_t> init((size_t)run_off[n_runs]), labels((size_t)run_off[n_runs]);
#pragma omp parallel
    {
        std::vector<int> loc, start;
        std::vector<char> mask;
#pragma omp for schedule(dynamic, 8)
        for (int w = 0; w < n_windows; w++) {
            if (!on_device(w)) continue;
            const ContigJob& job = *jobs[win_ref[w].first];
            const Window& win = job.windows[win_ref[w].second];
            const ColContig& c = contigs[job.n];
            const int R = (int)c.read_lines.size();
            loc.assign((size_t)R, -1);
            mask.assign((size_t)R, 0);
            for (size_t i = 0; i < win.masked.size(); i++) {
                loc[win.masked[i]] = (int)i;
                mask[win.masked[i]] = 1;
            }
            for (size_t k = 0; k < win.restart_snps.size(); k++) {
                snp_start_labels(c.snps[win.restart_snps[k]], mask, start);
                int32_t* dst = init.data() + run_off[run_base[w] + (int64_t)k];
                for (size_t i = 0; i < win.masked.size(); i++) dst[i] = loc[start[win.masked[i]]];
            }
        }
    }
    // sweep orders: the reference reshuffles all reads of the contig before every sweep
    const int n_orders = sh.pinned ? 1 : 15;
    std::vector<int32_t> rank;
    for (int j = 0; j < nc; j++) {
        const int R = n_reads[j];
        for (int k = 0; k < n_orders; k++) {
            const std::vector<int>& o = sh.order(R);
            const size_t base = rank.size();
            rank.resize(base + (size_t)R);
            for (int t = 0; t < R; t++) rank[base + (size_t)o[t]] = t;
        }
    }
    GPU_CHECK(ctx, hsgpu_graph_whispers(graph, n_runs, run_window.data(), init.data(), n_orders, rank.data(), labels.data()));
    phase("  gpu: chinese whispers");

    // ---- hand the results to the per-window post-processing ----
#pragma omp parallel for schedule(dynamic, 8)
    for (int w = 0; w < n_windows; w++) {
        ContigJob& job = *jobs[win_ref[w].first];
        const int wi = win_ref[w].second;
        const Window& win = job.windows[wi];
        const int R = (int)contigs[job.n].read_lines.size();
        const int m = (int)win.masked.size();
        ReadGraph& g = job.graphs[wi];
        g.clear(R);
        for (int i = 0; i < m; i++) g.off[win.masked[i] + 1] = (int)(adj_off[win_off[w] + i + 1] - adj_off[win_off[w] + i]);
        for (int r = 0; r < R; r++) g.off[r + 1] += g.off[r];
        g.nbr.resize((size_t)g.off[R]);
        for (int i = 0; i < m; i++) {
            int o = g.off[win.masked[i]];
            for (int64_t e = adj_off[win_off[w] + i]; e < adj_off[win_off[w] + i + 1]; e++) g.nbr[o++] = win.masked[adj[e]];
        }
        if (win_low[w]) {  // neighbour lists of create_read_graph_low_memory: the caller clusters
            g.list_mode = true;
            continue;
        }
        auto& lc = job.local_clusters[wi];
        lc.assign(win.restart_snps.size(), std::vector<int>());
        if (!on_device(w)) {
            // a window too large for the clustering kernel: chinese_whispers_high_memory on the host, on the graph the
            // device built (each thread shuffles with its own generator, like the per-window loop of the pipeline)
            Shuffler local_sh;
            std::vector<char> mask((size_t)R, 0);
            std::vector<int> start;
            for (int r : win.masked) mask[r] = 1;
            for (size_t k = 0; k < win.restart_snps.size(); k++) {
                snp_start_labels(contigs[job.n].snps[win.restart_snps[k]], mask, start);
                lc[k] = chinese_whispers(g, start, mask, local_sh);
            }
            continue;
        }
        for (size_t k = 0; k < win.restart_snps.size(); k++) {
            lc[k].assign((size_t)R, -2);
            const int32_t* src = labels.data() + run_off[run_base[w] + (int64_t)k];
            for (int i = 0; i < m; i++) lc[k][win.masked[i]] = src[i] >= 0 ? win.masked[src[i]] : src[i];
        }
    }
    for (int j = 0; j < nc; j++)
        if (jobs[j]->low_now) jobs[j]->device_lists = true;
    stats[0] += n_windows;
    stats[1] += n_runs;
    stats[2] += win_off.back();
    stats[3] += replayed;
    hsgpu_graph_destroy(graph);
    hsgpu_pairs_destroy(pairs);
}

struct GpuState {
    int n_gpus = 1, first_device = 0;
    std::vector<hsgpu_ctx*> ctxs;
};

static int gpu_prepare(void* user) {
    GpuState& st = *(GpuState*)user;
    for (int g = 0; g < st.n_gpus; g++) {
        if (hsgpu_ctx_create(st.first_device + g, &st.ctxs[g]) != HSGPU_OK) {
            std::cout << "ERROR: no usable GPU " << st.first_device + g << ": " << hsgpu_last_error(nullptr) << std::endl;
            return 1;
        }
    }
    return 0;
}

// shards the contigs over the GPUs, heaviest (reads^2) first onto the least loaded device; a shard is processed in
// batches whose dense count matrices stay under ~12 GB
static void gpu_stages(void* user, const std::vector<ColContig>& contigs, std::vector<ContigJob*>& jobs, float error_rate,
                       Shuffler& sh, int64_t* stats) {
    GpuState& st = *(GpuState*)user;
    std::vector<ContigJob*> order = jobs;
    auto weight = [&](const ContigJob* j) { return (double)contigs[j->n].read_lines.size() * contigs[j->n].read_lines.size(); };
    std::sort(order.begin(), order.end(), [&](const ContigJob* a, const ContigJob* b) {
        return weight(a) != weight(b) ? weight(a) > weight(b) : a->n < b->n;
    });
    std::vector<std::vector<ContigJob*>> shard((size_t)st.n_gpus);
    std::vector<double> load((size_t)st.n_gpus, 0.0);
    for (ContigJob* j : order) {
        const int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        shard[g].push_back(j);
        load[g] += weight(j);
    }
    for (int g = 0; g < st.n_gpus; g++) {
        std::sort(shard[g].begin(), shard[g].end(), [](const ContigJob* a, const ContigJob* b) { return a->n < b->n; });
        size_t i = 0;
        while (i < shard[g].size()) {
            std::vector<ContigJob*> batch;
            double blocks = 0, rows = 0, widest = 0;
            while (i < shard[g].size()) {
                // upper bound of what the batch keeps on the device: the count blocks of every tile pair (8 bytes per
                // read pair; the band of overlapping reads in practice) and the two one-hot operands, whose rows all
                // have the stride of the batch's widest contig (hsgpu_pairs_create: k_ld = the largest SNP count) --
                // so one SNP-rich contig is not batched with hundreds of read-rich ones
                const ColContig& cc = contigs[shard[g][i]->n];
                const double r = (double)((cc.read_lines.size() + 127) / 128 * 128);
                const double k = (double)((cc.snps.size() + 127) / 128 * 128);
                const double need = (blocks + r * r * 8) + (rows + r) * 2.0 * std::max(widest, k);
                if (!batch.empty() && need > 24e9) break;
                blocks += r * r * 8;
                rows += r;
                widest = std::max(widest, k);
                batch.push_back(shard[g][i++]);
            }
            gpu_shard(st.ctxs[g], contigs, batch, error_rate, sh, stats);
        }
    }
}

int main(int argc, char* argv[]) {
    GpuState st;
    if (const char* e = std::getenv("HSGPU_NGPUS")) st.n_gpus = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("HSGPU_DEVICE")) st.first_device = std::atoi(e);
    st.ctxs.assign((size_t)st.n_gpus, nullptr);
    g_timing = std::getenv("HS_TIMING") != nullptr;
    const int rc = separate_reads_pipeline(argc, argv, gpu_prepare, gpu_stages, &st, gpu_stages);
    // The .gro file is written and closed (scoped streams in the pipeline). Skip the unwinding (CUDA context
    // teardown in the runtime's atexit handler, freeing the parsed columns): the operating system reclaims both.
    std::cout.flush();
    std::cerr.flush();
    fflush(nullptr);
    if (!std::getenv("HS_FULL_TEARDOWN")) _exit(rc);
    for (hsgpu_ctx* c : st.ctxs)
        if (c) hsgpu_ctx_destroy(c);
    return rc;
}
