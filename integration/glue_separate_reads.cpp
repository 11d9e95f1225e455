// Reference-side glue for HS_separate_reads, compiled against the reference's own headers
// and run under the reference's own, unmodified main().
//
// Defines list_similarities_and_differences_between_reads3 with the reference's signature
// (src/separate_reads.h:34-37): the two Eigen sparse products over the one-hot SNP matrices
// (src/separate_reads.cpp:374-433) become one call of hsgpu_read_pair_counts -- the tcgen05 int8 kernel of
// csrc/pairs.cu -- and the dense int32 counts go back into the two Eigen::SparseMatrix<int> the rest of the reference
// reads (create_read_graph_matrix walks their columns with InnerIterator and only looks at values, :737-748).
// oracle/Makefile builds oracle/_ref/HS_separate_reads_glued from it: the reference's objects sit in a shared
// library (libhsref_sr_open.so: main() renamed at compile time, nothing else touched), whose call to the function
// goes through the PLT and binds to the definition below. The reference keeps doing everything else: argv,
// parse_column_file, the window walk, create_read_graph_matrix, chinese_whispers, finalize_clustering, the GROUP writer.
// Linked with oracle/ref_pin_rng.cpp like HS_separate_reads_pinned, so that the .gro can be compared byte for byte.
#include <cstdlib>
#include <iostream>
#include <vector>

#include "separate_reads.h"  // the reference's (src/), found through -I

#include "hsgpu.h"

namespace {
thread_local hsgpu_ctx* g_ctx = nullptr;

hsgpu_ctx* context() {
    if (!g_ctx) {
        int device = 0;
        if (const char* e = std::getenv("HSGPU_DEVICE")) device = std::atoi(e);
        const int rc = hsgpu_ctx_create(device, &g_ctx);
        if (rc != HSGPU_OK) {
            std::cout << "ERROR: hsgpu_ctx_create failed (" << rc << "): " << hsgpu_last_error(nullptr) << std::endl;
            std::exit(1);  // no CPU fallback
        }
    }
    return g_ctx;
}
}  // namespace

// src/separate_reads.cpp:374-433
void list_similarities_and_differences_between_reads3(std::vector<Column>& snps, Eigen::SparseMatrix<int>& similarity,
                                                      Eigen::SparseMatrix<int>& difference) {
    hsgpu_ctx* ctx = context();
    const int32_t n_reads = (int32_t)similarity.rows();
    const int32_t n_snps = (int32_t)snps.size();
    // the SNP columns as CSR, as parse_column_file left them
    std::vector<int64_t> snp_off((size_t)n_snps + 1, 0);
    for (int32_t s = 0; s < n_snps; s++) snp_off[(size_t)s + 1] = snp_off[(size_t)s] + (int64_t)snps[(size_t)s].readIdxs.size();
    std::vector<uint32_t> read_idx((size_t)std::max<int64_t>(snp_off[(size_t)n_snps], 1));
    std::vector<uint8_t> code((size_t)std::max<int64_t>(snp_off[(size_t)n_snps], 1));
    std::vector<uint8_t> ref_base((size_t)std::max(n_snps, 1)), second_base((size_t)std::max(n_snps, 1));
    for (int32_t s = 0; s < n_snps; s++) {
        const Column& c = snps[(size_t)s];
        ref_base[(size_t)s] = c.ref_base;
        second_base[(size_t)s] = c.second_base;
        for (size_t r = 0; r < c.readIdxs.size(); r++) {
            read_idx[(size_t)snp_off[(size_t)s] + r] = c.readIdxs[r];
            code[(size_t)snp_off[(size_t)s] + r] = c.content[r];
        }
    }
    std::vector<int32_t> sim((size_t)n_reads * (size_t)n_reads + 1), diff((size_t)n_reads * (size_t)n_reads + 1);
    const int rc = hsgpu_read_pair_counts(ctx, n_reads, n_snps, snp_off.data(), read_idx.data(), code.data(), ref_base.data(),
                                          second_base.data(), sim.data(), diff.data());
    if (rc != HSGPU_OK) {
        std::cout << "ERROR: hsgpu_read_pair_counts failed (" << rc << "): " << hsgpu_last_error(ctx) << std::endl;
        std::exit(1);
    }
    // back into the reference's containers: one entry per non-zero count, column by column
    std::vector<Eigen::Triplet<int>> sim_entries, diff_entries;
    for (int32_t j = 0; j < n_reads; j++)
        for (int32_t i = 0; i < n_reads; i++) {
            const size_t at = (size_t)i * (size_t)n_reads + (size_t)j;
            if (sim[at]) sim_entries.push_back(Eigen::Triplet<int>(i, j, sim[at]));
            if (diff[at]) diff_entries.push_back(Eigen::Triplet<int>(i, j, diff[at]));
        }
    similarity = Eigen::SparseMatrix<int>(n_reads, n_reads);
    difference = Eigen::SparseMatrix<int>(n_reads, n_reads);
    similarity.setFromTriplets(sim_entries.begin(), sim_entries.end());
    difference.setFromTriplets(diff_entries.begin(), diff_entries.end());
}

int hs_ref_separate_reads_main(int argc, char* argv[]);  // the reference's main(), renamed when the library is built

int main(int argc, char* argv[]) { return hs_ref_separate_reads_main(argc, argv); }
