// TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the read-graph and chinese-whispers stages of the
// reference's separate_reads module, used by tests/ to check the CUDA path (hsgpu_graph_*) and, through
// sr_hostcheck (below), the host logic of the HS_separate_reads drop-in without a GPU. Nothing in the product
// path links, loads or executes this file.
//
// C++ rather than C because the reference's results depend on libstdc++'s std::sort (order of equal
// distances, src/separate_reads.cpp:774-776) and std::shuffle + std::mt19937 (sweep order,
// src/cluster_graph.cpp:255-258): the oracle calls the same library functions on the same sequences.
// Parity pinned: tests/test_oracle_sr.py compares both functions with the compiled reference
// (oracle/_ref/libhsref_sr.so, std::random_device pinned by oracle/ref_pin_rng.cpp), and sr_hostcheck's .gro
// with oracle/_ref/HS_separate_reads_pinned byte for byte.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

extern "C" {
void hso_read_pair_counts(int32_t n_reads, int32_t n_snps, const int64_t* snp_off, const uint32_t* read_idx,
                          const uint8_t* code, const uint8_t* ref_base, const uint8_t* second_base, int32_t* sim,
                          int32_t* diff);

// create_read_graph_matrix (src/separate_reads.cpp:706-828) on dense n x n row-major count matrices. masked = the
// ascending read indices with mask == true. Output: CSR over the m masked reads, neighbours as LOCAL indices
// (position in masked), ascending; returns the number of entries (adj may be NULL to size it).
int64_t hso_read_graph(int32_t n_reads, const int32_t* sim, const int32_t* diff, int32_t m, const int32_t* masked,
                       float error_rate, int64_t* adj_off, int32_t* adj) {
    std::vector<char> mask((size_t)n_reads, 0);
    std::vector<int> local((size_t)n_reads, -1);
    for (int i = 0; i < m; i++) {
        mask[masked[i]] = 1;
        local[masked[i]] = i;
    }
    std::vector<std::vector<char>> linked((size_t)m, std::vector<char>((size_t)m, 0));
    std::vector<float> dist((size_t)n_reads);
    std::vector<int> sims((size_t)n_reads), diffs((size_t)n_reads);
    std::vector<std::pair<int, float>> smallest;
    for (int read1 = 0; read1 < n_reads; read1++) {
        if (!mask[read1]) continue;
        std::fill(dist.begin(), dist.end(), 0.f);
        std::fill(sims.begin(), sims.end(), 0);
        std::fill(diffs.begin(), diffs.end(), 0);
        int max_compat = 0;
        // column read1 of the (symmetric) sparse matrices: only masked rows other than read1 are read (:741-752)
        for (int read2 = 0; read2 < n_reads; read2++) {
            if (mask[read2] && read1 != read2) {
                sims[read2] = sim[(size_t)read2 * n_reads + read1];
                diffs[read2] = diff[(size_t)read2 * n_reads + read1];
            }
        }
        for (int r = 0; r < n_reads; r++) {
            if (mask[r] && r != read1 && sims[r] > 0) {
                float d = std::max(0, diffs[r] - 1);
                dist[r] = 1 - d / float(sims[r] + diffs[r]);
                if (sims[r] > max_compat) max_compat = sims[r];
            }
        }
        for (int r = 0; r < n_reads; r++)
            if (mask[r] && r != read1 && sims[r] + diffs[r] < 0.7 * max_compat) dist[r] = 0;
        smallest.clear();
        for (int r = 0; r < n_reads; r++) smallest.push_back(std::make_pair(r, dist[r]));
        std::sort(smallest.begin(), smallest.end(),
                  [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second > b.second; });
        int nb = 0;
        float below = 1 - error_rate * 2;
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
        for (const auto& nbr : smallest) {
            if (nbr.second > below && (nb < 5 || nbr.second == 1 || nbr.second >= above) && mask[nbr.first]) {
                nb++;
                linked[local[read1]][local[nbr.first]] = 1;
                linked[local[nbr.first]][local[read1]] = 1;
            }
        }
    }
    int64_t n = 0;
    for (int i = 0; i < m; i++) {
        adj_off[i] = n;
        for (int j = 0; j < m; j++) {
            if (linked[i][j]) {
                if (adj) adj[n] = j;
                n++;
            }
        }
    }
    adj_off[m] = n;
    return n;
}

// chinese_whispers_high_memory (src/cluster_graph.cpp:240-310) on the graph above. Labels and the graph are in local
// indices (order-preserving: masked is ascending), the sweep order is the shuffle of ALL n_reads reads, seeded before
// sweep s with seeds[min(s, n_seeds-1)], restricted to the masked ones -- what the reference does when
// std::random_device returns that sequence (one constant seed = the pinned reference build).
void hso_chinese_whispers(int32_t n_reads, int32_t m, const int32_t* masked, const int64_t* adj_off, const int32_t* adj,
                          const int32_t* init, int32_t n_seeds, const uint32_t* seeds, int32_t* labels) {
    std::vector<int> local((size_t)n_reads, -1);
    for (int i = 0; i < m; i++) local[masked[i]] = i;
    std::vector<int> clusters(init, init + m);
    int changes = 3, iterations = 0;
    while (changes > 2 && iterations < 15) {
        changes = 0;
        std::vector<int> order((size_t)n_reads);
        std::iota(order.begin(), order.end(), 0);
        std::mt19937 g(seeds[std::min(iterations, n_seeds - 1)]);
        std::shuffle(order.begin(), order.end(), g);
        for (int read : order) {
            const int i = local[read];
            if (i < 0) continue;
            std::vector<int> votes((size_t)m, 0);
            for (int64_t e = adj_off[i]; e < adj_off[i + 1]; e++)
                if (clusters[adj[e]] >= 0) votes[clusters[adj[e]]] += 1;
            int max_index = 0, max_value = 0;
            for (int j = 0; j < m; j++) {
                if (votes[j] > max_value) {
                    max_value = votes[j];
                    max_index = j;
                }
            }
            if (max_value > 0) {
                if (clusters[i] != max_index) changes++;
                clusters[i] = max_index;
            }
        }
        iterations += 1;
    }
    std::memcpy(labels, clusters.data(), sizeof(int32_t) * (size_t)m);
}

// 0..n-1 after std::shuffle with std::mt19937(seed): the node order of one sweep
void hso_shuffled_order(int32_t n, uint32_t seed, int32_t* out) {
    std::vector<int> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::mt19937 g(seed);
    std::shuffle(order.begin(), order.end(), g);
    for (int i = 0; i < n; i++) out[i] = order[i];
}
}  // extern "C"

#ifdef HS_HOSTCHECK_MAIN
// sr_hostcheck: the product's HS_separate_reads pipeline (hairsplitter_b200/host/hs_sepreads*.cpp) with the three GPU
// stages replaced by the oracle functions above. CPU-only check of the host logic against the pinned reference.
#include <cstdlib>

#include "../hairsplitter_b200/host/hs_sepreads.h"

static void oracle_stages(void*, const std::vector<hs::ColContig>& contigs, std::vector<hs::ContigJob*>& jobs, float error_rate,
                          hs::Shuffler& sh, int64_t* stats) {
    for (hs::ContigJob* job : jobs) {
        const hs::ColContig& c = contigs[job->n];
        const int R = (int)c.read_lines.size();
        std::vector<int64_t> snp_off(c.snps.size() + 1, 0);
        std::vector<uint32_t> idx;
        std::vector<uint8_t> code, rb, sb;
        for (size_t s = 0; s < c.snps.size(); s++) {
            idx.insert(idx.end(), c.snps[s].readIdxs.begin(), c.snps[s].readIdxs.end());
            code.insert(code.end(), c.snps[s].content.begin(), c.snps[s].content.end());
            rb.push_back(c.snps[s].ref_base);
            sb.push_back(c.snps[s].second_base);
            snp_off[s + 1] = (int64_t)idx.size();
        }
        std::vector<int32_t> sim((size_t)R * R), diff((size_t)R * R);
        hso_read_pair_counts(R, (int32_t)c.snps.size(), snp_off.data(), idx.data(), code.data(), rb.data(), sb.data(), sim.data(),
                             diff.data());
        for (size_t w = 0; w < job->windows.size(); w++) {
            const hs::Window& win = job->windows[w];
            if (!win.has_snps) continue;
            const int m = (int)win.masked.size();
            std::vector<int64_t> adj_off((size_t)m + 1);
            const int64_t n_adj = hso_read_graph(R, sim.data(), diff.data(), m, win.masked.data(), error_rate, adj_off.data(), nullptr);
            std::vector<int32_t> adj((size_t)std::max<int64_t>(n_adj, 1));
            hso_read_graph(R, sim.data(), diff.data(), m, win.masked.data(), error_rate, adj_off.data(), adj.data());
            hs::ReadGraph& g = job->graphs[w];
            g.clear(R);
            for (int i = 0; i < m; i++) g.off[win.masked[i] + 1] = (int)(adj_off[i + 1] - adj_off[i]);
            for (int r = 0; r < R; r++) g.off[r + 1] += g.off[r];
            g.nbr.resize((size_t)g.off[R]);
            for (int i = 0; i < m; i++) {
                int o = g.off[win.masked[i]];
                for (int64_t e = adj_off[i]; e < adj_off[i + 1]; e++) g.nbr[o++] = win.masked[adj[e]];
            }
            std::vector<char> mask((size_t)R, 0);
            std::vector<int> loc((size_t)R, -1), start;
            for (int i = 0; i < m; i++) {
                mask[win.masked[i]] = 1;
                loc[win.masked[i]] = i;
            }
            auto& lc = job->local_clusters[w];
            lc.clear();
            for (int s : win.restart_snps) {
                hs::snp_start_labels(c.snps[s], mask, start);
                std::vector<int32_t> init((size_t)m), out((size_t)m);
                for (int i = 0; i < m; i++) init[i] = loc[start[win.masked[i]]];
                hso_chinese_whispers(R, m, win.masked.data(), adj_off.data(), adj.data(), init.data(), 1, &sh.pin, out.data());
                std::vector<int> full((size_t)R, -2);
                for (int i = 0; i < m; i++) full[win.masked[i]] = win.masked[out[i]];
                lc.push_back(std::move(full));
                stats[1]++;
            }
            stats[0]++;
            stats[2] += m;
        }
    }
}

int main(int argc, char* argv[]) {
    if (!std::getenv("HS_PIN_SEED")) setenv("HS_PIN_SEED", "20260117", 1);
    return hs::separate_reads_pipeline(argc, argv, nullptr, oracle_stages, nullptr);
}
#endif
