// Host logic of the HS_separate_reads drop-in (reference src/separate_reads.cpp, src/cluster_graph.cpp):
// the .col parser, the window/mask walk of main(), the sequential clustering post-processing
// (finalize_clustering and what it calls) and the low-memory graph path. The data-parallel parts -- the
// read x read counts, the read graph of every window and the chinese-whispers runs started from every SNP --
// run on the GPU through libhsgpu (hsgpu_pairs_*, hsgpu_graph_*); see separate_reads_main.cpp.
#pragma once
#include <cstdint>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "hs_types.h"

namespace hs {

// one CONTIG block of the .col file as parse_column_file leaves it (src/separate_reads.cpp:46-190)
struct ColContig {
    std::string line;                     // the CONTIG line, verbatim (name_of_contigs)
    std::vector<std::string> read_lines;  // the READ lines, verbatim (names_of_reads)
    long length = 0;
    double coverage = 0;
    std::vector<Column> snps;
    std::vector<std::pair<int, int>> limits;  // readLimits
};

// max_coverage: reads beyond this many cells of a SNP are dropped (:157); the reference passes an
// uninitialised int when rarest_strain_abundance != 0 (:1420-1426), which behaves as "no limit".
void parse_column_file(const std::string& path, std::vector<ColContig>& contigs, int max_coverage,
                       float rarest_strain_abundance);
// the text route alone (parse_column_file first tries the writer's binary sidecar, hs_colbin.h)
void parse_column_text(const std::string& path, std::vector<ColContig>& contigs, int max_coverage,
                       float rarest_strain_abundance);

// Where the seeds of the per-sweep shuffles come from (cluster_graph.cpp:255-258,429-432: a fresh
// std::random_device-seeded mt19937 per sweep). HS_PIN_SEED=<n> in the environment replaces random_device by the
// constant n -- the same pin the test build of the reference uses -- which makes the output reproducible.
struct Shuffler {
    bool pinned = false;
    uint32_t pin = 0;
    std::random_device rd;
    std::unordered_map<int, std::vector<int>> cache;  // pinned: one order per n
    std::vector<int> scratch;
    Shuffler();
    uint32_t seed() { return pinned ? pin : (uint32_t)rd(); }
    // 0..n-1 shuffled exactly as the reference does it (std::shuffle with a std::mt19937)
    const std::vector<int>& order(int n);
};

// symmetric 0/1 graph over the reads of a contig, neighbours ascending (the column iteration of the
// reference's Eigen matrix, or its sorted neighbour lists)
struct ReadGraph {
    int n = 0;
    std::vector<int> off;  // n+1
    std::vector<int> nbr;
    bool list_mode = false;  // built by create_read_graph_low_memory (neighbour lists)
    void clear(int n_reads) {
        n = n_reads;
        off.assign((size_t)n_reads + 1, 0);
        nbr.clear();
    }
};

// one window of main()'s walk (:1523-1622)
struct Window {
    int chunk = 0;
    int start = 0, end = 0;       // the GROUP coordinates
    bool has_snps = false;
    std::vector<int> reads_here;  // no-SNP windows: the finished haplotype vector (0 / -2)
    std::vector<int> masked;      // reads spanning the window, ascending
    std::vector<int> restart_snps;  // indices in snps of the SNPs a clustering run starts from (:1677)
};

void plan_windows(const ColContig& c, int size_of_window, std::vector<Window>& out);

// starting labels of the run launched from one SNP (:1680-1693), as read indices; n = number of reads
void snp_start_labels(const Column& snp, const std::vector<char>& mask, std::vector<int>& labels);

// chinese_whispers / chinese_whispers_high_memory (cluster_graph.cpp:152-310)
std::vector<int> chinese_whispers(const ReadGraph& g, const std::vector<int>& initial, const std::vector<char>& mask,
                                  Shuffler& sh);

// create_read_graph_low_memory (:538-693)
void create_read_graph_low_memory(const std::vector<Column>& snps, const std::vector<char>& mask, ReadGraph& g,
                                  float error_rate);

// finalize_clustering (:897-993) with everything below it; graph = what the reference hands over for the value of
// its `low_memory` argument (neighbour lists, or the Eigen adjacency matrix)
void finalize_clustering(const std::vector<Column>& snps, const std::vector<std::vector<int>>& local_clusters,
                         const ReadGraph& g, bool low_memory, const std::vector<char>& mask, std::vector<int>& haplotypes,
                         int posstart, int posend, Shuffler& sh);

// merge_haplotypes_to_fit_within_limit (:1342-1396)
std::vector<int> merge_haplotypes_to_fit_within_limit(int max_haplotypes, const std::vector<int>& clusters,
                                                      const std::vector<char>& mask, const ReadGraph& g, Shuffler& sh);

// per-contig state of the pipeline
struct ContigJob {
    int n = 0;             // index in contigs
    bool low_now = false;  // coverage > 1000 or -l: the neighbour-list path (create_read_graph_low_memory)
    bool device_lists = false;  // low_now: the stage provider built the neighbour lists (graphs[w], list mode)
    std::vector<Window> windows;
    // per window with SNPs: the read graph and the clusterings started from its SNPs (filled by the stage provider
    // for the high-memory contigs)
    std::vector<ReadGraph> graphs;                              // indexed like windows
    std::vector<std::vector<std::vector<int>>> local_clusters;  // [window][run][read]
    std::vector<std::vector<int>> haplotypes;                   // [window][read]
};

// The data-parallel stages of a batch of high-memory contigs: list_similarities_and_differences_between_reads3,
// create_read_graph_matrix for every window with SNPs, and one chinese_whispers_high_memory run per restart SNP.
// Fills job.graphs[w] and job.local_clusters[w]. The product passes the libhsgpu implementation
// (separate_reads_main.cpp); stats = {windows, runs, masked reads, sort replays}.
typedef void (*SepStages)(void* user, const std::vector<ColContig>& contigs, std::vector<ContigJob*>& jobs, float error_rate,
                          Shuffler& sh, int64_t* stats);

// main() of HS_separate_reads (src/separate_reads.cpp:1398-1790) behind the stage provider. `prepare` runs
// concurrently with the .col parser (the product creates its GPU contexts there) and returns non-zero on failure.
// `low_stages` (may be null): the same provider interface for the low-memory contigs whose reads all cover runs of
// consecutive SNP columns (see low_memory_counts_are_contig_counts): it fills job.graphs[w] with the neighbour lists of
// create_read_graph_low_memory and sets job.device_lists; the clusterings of those contigs stay on the host.
int separate_reads_pipeline(int argc, char* argv[], int (*prepare)(void* user), SepStages stages, void* user,
                            SepStages low_stages = nullptr);

// create_read_graph_low_memory addresses a read's alleles by offset from its first SNP (:597-605). When every read
// appears in a run of consecutive SNP columns that is the contig-wide similarity / difference count of the pair.
bool low_memory_counts_are_contig_counts(const ColContig& c);

}  // namespace hs
