// TEST INFRASTRUCTURE ONLY (oracle/): thin C-ABI shim around the UNMODIFIED reference's separate_reads
// module (RolandFaure/Hairsplitter, sources compiled where they lie under /root/reference/src), so the
// Python tests and bench.py's CPU baseline can call the reference's own functions on flat arrays.
// Nothing in the product path links or loads this file.
//
// Reference entry points wrapped here:
//   list_similarities_and_differences_between_reads3   src/separate_reads.cpp:374  (Eigen sparse products)
//   list_similarities_and_differences_between_reads2   src/separate_reads.cpp:323  (its dense restatement)
//   create_read_graph_matrix                           src/separate_reads.cpp:706
//   create_read_graph_low_memory                       src/separate_reads.cpp:538  (the -l / amplicon / > 1000x path)
//   chinese_whispers_high_memory                       src/cluster_graph.cpp:240  (std::random_device pinned by
//                                                      ref_pin_rng.cpp, linked into this library with -Bsymbolic)
#include <cstdint>
#include <cstring>
#include <vector>

#include "cluster_graph.h"
#include "separate_reads.h"

// the definition in src/separate_reads.cpp:538-544 (the header declares another parameter type for the lists)
void create_read_graph_low_memory(std::vector<Column>& snps, std::vector<bool>& mask, int chunk, int sizeOfWindow,
                                  std::vector<std::vector<int>>& neighbor_list_low_memory, float& errorRate);

static std::vector<Column> make_columns(int n_snps, const int64_t* snp_off, const uint32_t* idx, const uint8_t* code,
                                        const uint8_t* rb, const uint8_t* sb) {
    std::vector<Column> snps(n_snps);
    for (int s = 0; s < n_snps; s++) {
        Column& c = snps[s];
        c.pos = s;
        c.ref_base = rb[s];
        c.second_base = sb[s];
        c.readIdxs.assign(idx + snp_off[s], idx + snp_off[s + 1]);
        c.content.assign(code + snp_off[s], code + snp_off[s + 1]);
    }
    return snps;
}

static void densify(const Eigen::SparseMatrix<int>& m, int n, int32_t* out) {
    std::memset(out, 0, sizeof(int32_t) * (size_t)n * n);
    for (int k = 0; k < m.outerSize(); ++k)
        for (Eigen::SparseMatrix<int>::InnerIterator it(m, k); it; ++it) out[(size_t)it.row() * n + it.col()] = it.value();
}

extern "C" {

// sim/diff: dense n x n row-major (may be null to time the reference call alone)
int hsref_read_pair_counts(int n_reads, int n_snps, const int64_t* snp_off, const uint32_t* idx, const uint8_t* code,
                           const uint8_t* rb, const uint8_t* sb, int32_t* sim, int32_t* diff) {
    std::vector<Column> snps = make_columns(n_snps, snp_off, idx, code, rb, sb);
    Eigen::SparseMatrix<int> similarity(n_reads, n_reads), difference(n_reads, n_reads);
    list_similarities_and_differences_between_reads3(snps, similarity, difference);
    if (sim) densify(similarity, n_reads, sim);
    if (diff) densify(difference, n_reads, diff);
    return 0;
}

// create_read_graph_matrix on the reference's own sparse count matrices. masked = ascending read indices with
// mask == true. Output: CSR over the masked reads, neighbours as local indices (Eigen's column iteration order).
// Returns the number of entries (adj may be NULL).
int64_t hsref_read_graph(int n_reads, int n_snps, const int64_t* snp_off, const uint32_t* idx, const uint8_t* code,
                         const uint8_t* rb, const uint8_t* sb, int m, const int32_t* masked, float error_rate,
                         int64_t* adj_off, int32_t* adj) {
    std::vector<Column> snps = make_columns(n_snps, snp_off, idx, code, rb, sb);
    Eigen::SparseMatrix<int> similarity(n_reads, n_reads), difference(n_reads, n_reads), adjacency(n_reads, n_reads);
    list_similarities_and_differences_between_reads3(snps, similarity, difference);
    std::vector<bool> mask(n_reads, false);
    std::vector<int> local(n_reads, -1);
    for (int i = 0; i < m; i++) {
        mask[masked[i]] = true;
        local[masked[i]] = i;
    }
    create_read_graph_matrix(mask, 0, 2000, similarity, difference, adjacency, error_rate);
    int64_t n = 0;
    for (int i = 0; i < m; i++) {
        adj_off[i] = n;
        for (Eigen::SparseMatrix<int>::InnerIterator it(adjacency, masked[i]); it; ++it) {
            if (it.value() == 0) continue;
            if (adj) adj[n] = local[it.row()];
            n++;
        }
    }
    adj_off[m] = n;
    return n;
}

// create_read_graph_low_memory over the SNP columns of a contig: the sorted neighbour lists of the masked reads, as
// a CSR over the masked reads with local indices (same output form as hsref_read_graph)
int64_t hsref_read_graph_low_memory(int n_reads, int n_snps, const int64_t* snp_off, const uint32_t* idx, const uint8_t* code,
                                    const uint8_t* rb, const uint8_t* sb, int m, const int32_t* masked, float error_rate,
                                    int64_t* adj_off, int32_t* adj) {
    std::vector<Column> snps = make_columns(n_snps, snp_off, idx, code, rb, sb);
    std::vector<bool> mask(n_reads, false);
    std::vector<int> local(n_reads, -1);
    for (int i = 0; i < m; i++) {
        mask[masked[i]] = true;
        local[masked[i]] = i;
    }
    std::vector<std::vector<int>> lists(n_reads);
    create_read_graph_low_memory(snps, mask, 0, 2000, lists, error_rate);
    int64_t n = 0;
    for (int i = 0; i < m; i++) {
        adj_off[i] = n;
        for (int r : lists[masked[i]]) {
            if (adj) adj[n] = local[r];
            n++;
        }
    }
    adj_off[m] = n;
    return n;
}

// chinese_whispers_high_memory on a graph given as the CSR above; init/labels are per masked read, as READ indices
void hsref_chinese_whispers(int n_reads, int m, const int32_t* masked, const int64_t* adj_off, const int32_t* adj,
                            const int32_t* init, int32_t* labels) {
    std::vector<Eigen::Triplet<int>> trip;
    for (int i = 0; i < m; i++)
        for (int64_t e = adj_off[i]; e < adj_off[i + 1]; e++) trip.push_back(Eigen::Triplet<int>(masked[adj[e]], masked[i], 1));
    Eigen::SparseMatrix<int> adjacency(n_reads, n_reads);
    adjacency.setFromTriplets(trip.begin(), trip.end());
    std::vector<bool> mask(n_reads, false);
    std::vector<int> start(n_reads);
    for (int r = 0; r < n_reads; r++) start[r] = r;
    for (int i = 0; i < m; i++) {
        mask[masked[i]] = true;
        start[masked[i]] = init[i];
    }
    std::vector<int> out = chinese_whispers_high_memory(adjacency, start, mask);
    for (int i = 0; i < m; i++) labels[i] = out[masked[i]];
}
}
