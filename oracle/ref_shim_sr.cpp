// TEST INFRASTRUCTURE ONLY (oracle/): thin C-ABI shim around the UNMODIFIED reference's separate_reads
// module (RolandFaure/Hairsplitter, sources compiled where they lie under /root/reference/src), so the
// Python tests and bench.py's CPU baseline can call the reference's own functions on flat arrays.
// Nothing in the product path links or loads this file.
//
// Reference entry points wrapped here:
//   list_similarities_and_differences_between_reads3   src/separate_reads.cpp:374  (Eigen sparse products)
//   list_similarities_and_differences_between_reads2   src/separate_reads.cpp:323  (its dense restatement)
#include <cstdint>
#include <cstring>
#include <vector>

#include "separate_reads.h"

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
}
