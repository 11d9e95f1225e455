// Device-resident result of hsgpu_pairs_compute as the later stages read it (graph.cu, the dense fetch of pairs.cu).
//
// The reference keeps similarity / difference as two R x R Eigen::SparseMatrix<int> (src/separate_reads.cpp:
// 414-432). Here only the tile pairs that share SNP blocks exist: reads are ordered by their first SNP, every
// scheduled pair of 128-read tiles (ti <= tj) owns one 128 x 128 int32 block per matrix (rows = reads of tile ti,
// columns = reads of tile tj), and a per-contig nt x nt map gives the block of a tile pair (-1: the reads of the
// two tiles share no SNP, all counts are zero). Memory follows the band of overlapping reads, not R^2.
#pragma once
#include <stdint.h>

#define HS_PV_TILE 128
#define HS_PV_BLOCK (HS_PV_TILE * HS_PV_TILE)

struct HsPairView {
    const int32_t* row_of;     // read (contig-local) -> row in first-SNP order (contig-local)
    const int32_t* tilemap;    // [nt * nt], valid for ti <= tj
    const int32_t* sim;        // blocks of all contigs of the batch
    const int32_t* diff;
    const uint8_t* has_cells;  // read (contig-local) -> 1 when it appears in at least one SNP column
    int32_t nt, n;
};

#ifdef __CUDACC__
__device__ __forceinline__ void hs_pair_get(const HsPairView& v, int r, int k, int& s, int& d) {
    int a = __ldg(v.row_of + r), b = __ldg(v.row_of + k);
    if ((a >> 7) > (b >> 7)) {  // both matrices are symmetric
        const int t = a;
        a = b;
        b = t;
    }
    const int slot = __ldg(v.tilemap + (a >> 7) * v.nt + (b >> 7));
    if (slot < 0) {
        s = 0;
        d = 0;
        return;
    }
    const int64_t off = (int64_t)slot * HS_PV_BLOCK + (a & 127) * HS_PV_TILE + (b & 127);
    s = __ldg(v.sim + off);
    d = __ldg(v.diff + off);
}
#endif
