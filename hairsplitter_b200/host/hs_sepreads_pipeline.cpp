// main() of the HS_separate_reads drop-in behind a stage provider; see hs_sepreads.h.
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "hs_sepreads.h"

namespace hs {

static double g_t0 = 0;
static bool g_timing = false;
static void phase(const char* name) {
    if (!g_timing) return;
    const double t = omp_get_wtime();
    if (g_t0 > 0) fprintf(stderr, "[hs timing] %-28s %8.3f s\n", name, t - g_t0);
    g_t0 = t;
}

static const char* kUsage =
    "Usage: ./separate_reads <columns> <num_threads> <error_rate> <ploidy_of_contigs> <low_memory> "
    "<rarest-strain-abundance> <amplicon> <outfile> <DEBUG>";

bool low_memory_counts_are_contig_counts(const ColContig& c) {
    const size_t R = c.read_lines.size();
    std::vector<int> first(R, -1), last(R, -1), count(R, 0);
    for (size_t s = 0; s < c.snps.size(); s++)
        for (int r : c.snps[s].readIdxs) {
            if (r < 0 || (size_t)r >= R) return false;
            if (first[r] < 0) first[r] = (int)s;
            last[r] = (int)s;
            count[r]++;
        }
    for (size_t r = 0; r < R; r++)
        if (count[r] > 0 && count[r] != last[r] - first[r] + 1) return false;
    return true;
}

int separate_reads_pipeline(int argc, char* argv[], int (*prepare)(void* user), SepStages stages, void* user,
                            SepStages low_stages) {
    if (argc != 10) {
        if (argc == 2 && (argv[1] == std::string("-h") || argv[1] == std::string("--help"))) {
            std::cout << kUsage << std::endl;
            return 0;
        }
        std::cout << kUsage << std::endl;
        return 1;
    }
    g_timing = std::getenv("HS_TIMING") != nullptr;
    phase("start");
    const std::string columns_file = argv[1];
    const int num_threads = std::max(1, std::atoi(argv[2]));
    const std::string ploidy_file = argv[4];
    const float error_rate = (float)std::atof(argv[3]);
    const bool amplicon = bool(std::atoi(argv[7]));
    const bool debug = bool(std::atoi(argv[9]));
    const std::string outfile = argv[8];
    const bool low_memory = bool(std::atoi(argv[5]));
    const float rarest = (float)std::atof(argv[6]);
    // the reference's max_coverage is 1e9 when rarest == 0 and an uninitialised (shadowed) int otherwise
    // (src/separate_reads.cpp:1420-1426); both behave as "no limit"
    const int max_coverage = 1000000000;
    omp_set_num_threads(num_threads);

    { std::ofstream out(outfile); }

    std::vector<ColContig> contigs;
    int prepare_rc = 0;
    omp_set_max_active_levels(2);  // the parser's own loops run as a team inside its section
#pragma omp parallel sections num_threads(2)
    {
#pragma omp section
        { prepare_rc = prepare ? prepare(user) : 0; }
#pragma omp section
        { parse_column_file(columns_file, contigs, max_coverage, rarest); }
    }
    phase("parse .col + prepare");
    if (prepare_rc) return 1;

    std::unordered_map<std::string, int> ploidy_of_contigs;
    {
        std::ifstream pf(ploidy_file);
        if (pf) {
            std::string line;
            while (std::getline(pf, line)) {
                std::istringstream iss(line);
                std::string contig;
                int ploidy;
                if (!(iss >> contig >> ploidy)) break;
                ploidy_of_contigs[contig] = ploidy;
            }
        }
    }

    // window size from the read lengths (:1465-1498); the reference sums the lengths in an int
    int n_reads_total = 0, above_4000 = 0;
    uint32_t sum_length = 0;
    std::vector<float> coverages(contigs.size(), 0);
    for (size_t n = 0; n < contigs.size(); n++) {
        for (const auto& r : contigs[n].limits) {
            n_reads_total++;
            sum_length += (uint32_t)(r.second - r.first + 1);
            coverages[n] += r.second - r.first + 1;
            if (r.second - r.first + 1 > 4000) above_4000++;
        }
        coverages[n] /= contigs[n].length;
    }
    const double mean_length = (int)sum_length / double(n_reads_total);
    int size_of_window = 2000;
    if (above_4000 < 20 && mean_length < 4000 && mean_length > 2000) size_of_window = 1000;
    else if (above_4000 < 20 && mean_length < 2000) size_of_window = 500;
    if (amplicon) {
        size_of_window = 0;
        for (const ColContig& c : contigs) size_of_window = std::max(size_of_window, int(c.length));
    }

    // ---- plan: windows, masks and restart SNPs of every contig ----
    std::vector<ContigJob> jobs(contigs.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t n = 0; n < contigs.size(); n++) {
        ContigJob& job = jobs[n];
        job.n = (int)n;
        job.low_now = low_memory || coverages[n] > 1000;
        if (contigs[n].snps.empty()) continue;
        plan_windows(contigs[n], size_of_window, job.windows);
        job.graphs.resize(job.windows.size());
        job.local_clusters.resize(job.windows.size());
        job.haplotypes.resize(job.windows.size());
    }
    phase("plan windows");

    // ---- the data-parallel stages of the high-memory contigs ----
    std::vector<ContigJob*> high;
    for (size_t n = 0; n < contigs.size(); n++)
        if (!contigs[n].snps.empty() && !jobs[n].low_now) high.push_back(&jobs[n]);
    Shuffler master;
    int64_t stats[4] = {0, 0, 0, 0};
    if (!high.empty()) stages(user, contigs, high, error_rate, master, stats);
    phase("graph + clustering stages");
    // ---- the neighbour lists of the low-memory contigs (create_read_graph_low_memory) through the provider, where
    // its pair counts are the reference's (HS_LOWMEM_HOST=1 keeps every such contig on the host loop) ----
    if (low_stages && !std::getenv("HS_LOWMEM_HOST")) {
        std::vector<ContigJob*> low;
        for (size_t n = 0; n < contigs.size(); n++) {
            if (contigs[n].snps.empty() || !jobs[n].low_now) continue;
            size_t max_m = 0;
            for (const Window& w : jobs[n].windows) max_m = std::max(max_m, w.masked.size());
            if (max_m <= 6400 && low_memory_counts_are_contig_counts(contigs[n])) low.push_back(&jobs[n]);
        }
        if (!low.empty()) low_stages(user, contigs, low, error_rate, master, stats);
        phase("low-memory read graphs");
    }

    // ---- low-memory contigs: neighbour lists and clusterings on the host (:1636-1645,1660-1700) ----
    struct Item { int n, w; };
    std::vector<Item> items;
    for (size_t n = 0; n < jobs.size(); n++)
        for (size_t w = 0; w < jobs[n].windows.size(); w++)
            if (jobs[n].windows[w].has_snps) items.push_back({(int)n, (int)w});
#pragma omp parallel
    {
        Shuffler sh;
        std::vector<char> mask;
        std::vector<int> start;
#pragma omp for schedule(dynamic, 4)
        for (size_t it = 0; it < items.size(); it++) {
            ContigJob& job = jobs[items[it].n];
            const int wi = items[it].w;
            const Window& win = job.windows[wi];
            const ColContig& c = contigs[job.n];
            const int R = (int)c.read_lines.size();
            mask.assign((size_t)R, 0);
            for (int r : win.masked) mask[r] = 1;
            ReadGraph lists;  // neighbor_list_low_memory_strengthened
            if (job.low_now) {
                if (job.device_lists) lists = std::move(job.graphs[wi]);
                else create_read_graph_low_memory(c.snps, mask, lists, error_rate);
                auto& lc = job.local_clusters[wi];
                lc.clear();
                for (int s : win.restart_snps) {
                    snp_start_labels(c.snps[s], mask, start);
                    lc.push_back(chinese_whispers(lists, start, mask, sh));
                }
            }
            // finalize_clustering receives `low_memory`, not `low_memory_now` (:1708): a contig pushed to the
            // low-memory path by its coverage alone is post-processed on the (empty) adjacency matrix
            ReadGraph empty;
            const ReadGraph* g = &job.graphs[wi];
            if (low_memory) g = &lists;
            else if (job.low_now) {
                empty.clear(R);
                g = &empty;
            }
            std::vector<int>& hap = job.haplotypes[wi];
            hap.assign((size_t)R, -2);
            finalize_clustering(c.snps, job.local_clusters[wi], *g, low_memory, mask, hap, win.chunk * size_of_window,
                                win.chunk * size_of_window + size_of_window, sh);
            // ploidy limit (:1711-1716)
            const std::string& line = c.line;
            const std::string after = line.substr(line.find("\t") + 1);
            const std::string clipped = after.substr(0, after.find("\t"));
            auto pit = ploidy_of_contigs.find(clipped);
            if (pit != ploidy_of_contigs.end() && pit->second > 0) hap = merge_haplotypes_to_fit_within_limit(pit->second, hap, mask, *g, sh);
            job.graphs[wi] = ReadGraph();
            job.local_clusters[wi].clear();
        }
    }
    phase("finalize windows");

    // ---- output (:1752-1787): contigs in index order ----
    {
        std::ofstream out(outfile, std::ios_base::app);
        std::string buf;
        for (size_t n = 0; n < contigs.size(); n++) {
            if (contigs[n].snps.empty()) continue;
            if (debug) std::cout << "separating reads on contig " << contigs[n].line << "\n";
            buf.clear();
            buf += contigs[n].line;
            buf += '\n';
            for (const std::string& r : contigs[n].read_lines) {
                buf += r;
                buf += '\n';
            }
            for (size_t w = 0; w < jobs[n].windows.size(); w++) {
                const Window& win = jobs[n].windows[w];
                const std::vector<int>& groups = win.has_snps ? jobs[n].haplotypes[w] : win.reads_here;
                buf += "GROUP\t" + std::to_string(win.start) + "\t" + std::to_string(win.end) + "\t";
                for (size_t h = 0; h < groups.size(); h++)
                    if (groups[h] != -2) buf += std::to_string(h) + ",";
                buf += "\t";
                for (size_t h = 0; h < groups.size(); h++)
                    if (groups[h] != -2) buf += std::to_string(groups[h]) + ",";
                buf += "\n";
            }
            out << buf;
        }
    }
    phase("write .gro");
    if (g_timing)
        fprintf(stderr, "[hs timing] windows %lld, clustering runs %lld, masked reads %lld, sort replays %lld\n", (long long)stats[0],
                (long long)stats[1], (long long)stats[2], (long long)stats[3]);
    return 0;
}

}  // namespace hs
