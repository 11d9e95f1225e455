// Read x read SNP agreement counts: list_similarities_and_differences_between_reads3 (reference
// src/separate_reads.cpp:374-433), the one dense contraction of the pipeline.
//
//   A[r,s] = 1 iff read r carries second_base at SNP s, R[r,s] = 1 iff it carries ref_base (:384-394)
//   similarity = 3*A*At + R*Rt,   difference = A*Rt + R*At,   diagonals zeroed (:414-432)
//
// B200 formulation. A and R are two K-major u8 matrices in HBM (one row per read, one byte per SNP), all
// contigs of a batch stacked, every contig padded to whole 128-row tiles. One work item = one pair of
// row tiles (i <= j) of one contig and the range of 128-SNP blocks in which both tiles have cells. A
// persistent, warp-specialised CTA per SM walks its share of the work list:
//   warp 0   TMA producer: four 128 x 128-byte boxes per stage (A_i, R_i, A_j, R_j; 128-byte swizzle)
//            through a 3-stage mbarrier ring -- the minimum bytes per SNP block, each tile loaded once;
//   warp 1   tcgen05.mma.kind::i8 issuer (one lane). Accumulators in TMEM, 384 columns:
//              [0,128)   = A_i * A_j^T      [128,256) = A_i * R_j^T + R_i * A_j^T      [256,384) = R_i * R_j^T
//            two M128 x N256 x K32 instructions per 32 SNPs: A_i x [A_j;R_j] lands on columns 0..255,
//            R_i x [A_j;R_j] on columns 128..383, so the middle block accumulates both cross products;
//   warps 2-5 epilogue: tcgen05.ld the three blocks, sim = 3*acc0 + acc2, diff = acc1, zero the
//            diagonal, store the 128 x 128 block of the tile pair row-wise (16-byte stores).
// Reads are ordered by their first SNP before tiling, so tiles far from the diagonal share no SNP block and
// are never scheduled: work AND output are the band of overlapping reads, not R^2 (pairs_view.cuh): every
// scheduled tile pair owns one block per matrix, a per-contig tile map finds it.
#include <cuda.h>  // CUtensorMap and its enums only; the encoder is fetched from the driver at run time

#include <algorithm>
#include <numeric>
#include <vector>

#include <atomic>

#include "common.cuh"
#include "pairs_view.cuh"

#define PG_TILE 128                     // rows per tile (UMMA M, and N per operand half)
#define PG_BK 128                       // SNPs (= bytes of K) per pipeline stage: one 128-byte swizzle row
#define PG_SUB (PG_TILE * PG_BK)        // one operand sub-tile: 16 KB
#define PG_STAGE_BYTES (4 * PG_SUB)     // A_i, R_i, A_j, R_j
#define PG_STAGES 3
#define PG_THREADS 192
#define PG_TMEM_COLS 512
#define PG_SMEM (PG_STAGES * PG_STAGE_BYTES + 1024 + 256)

struct PairWork {
    int32_t contig;
    int32_t ti, tj;    // row tiles within the contig, ti <= tj
    int32_t kb0, kb1;  // SNP blocks [kb0, kb1)
};

struct PairContig {
    int64_t row0;      // first row of the contig in the stacked operand matrices
    int64_t map_off;   // offset of the contig's nt x nt tile map
    int64_t read0;     // first read of the contig in the per-read arrays (row_of, has_cells)
    int32_t n, n_pad;
};

struct hsgpu_pairs {
    hsgpu_ctx* ctx = nullptr;
    int32_t n_contigs = 0;
    int32_t flags = 0;
    int64_t total_rows = 0, k_ld = 0, n_cells = 0, n_work = 0;
    int64_t kblocks_listed = 0, kblocks_dense = 0, tiles_dense = 0;
    std::vector<PairContig> h_contigs;
    uint8_t *d_A = nullptr, *d_R = nullptr;
    int32_t* d_rowmap = nullptr;   // stacked row -> read index in the contig (-1 = padding)
    int32_t* d_rowof = nullptr;    // per read: stacked row (inverse map, for the one-hot scatter)
    int32_t* d_rowloc = nullptr;   // per read: row inside its contig (what HsPairView.row_of points into)
    uint8_t* d_has_cells = nullptr;  // per read: appears in at least one SNP column
    PairContig* d_contigs = nullptr;
    PairWork* d_work = nullptr;
    int32_t *d_sim = nullptr, *d_diff = nullptr;  // one 128 x 128 block per scheduled tile pair (work item)
    int32_t* d_tilemap = nullptr;                 // per contig nt x nt: block of the tile pair (ti <= tj), -1 = none
    bool all_identity = true;
    int64_t* d_read_base = nullptr;
    int32_t* d_err = nullptr;
    // the SNP columns (inputs), kept until the operands are built
    int64_t *d_snp_off = nullptr, *d_snp_base = nullptr;
    uint32_t* d_read_idx = nullptr;
    uint8_t *d_code = nullptr, *d_rb = nullptr, *d_sb = nullptr;
    int32_t* d_snp_contig = nullptr;
    int64_t total_snps = 0;
    CUtensorMap tmapA, tmapR;
    bool computed = false;
};

// ---- operand construction ---------------------------------------------------------------------------------
__global__ void onehot_kernel(int64_t total_snps, const int32_t* __restrict__ snp_contig,
                              const int64_t* __restrict__ snp_base, const int64_t* __restrict__ snp_off,
                              const uint32_t* __restrict__ read_idx, const uint8_t* __restrict__ code,
                              const uint8_t* __restrict__ ref_base, const uint8_t* __restrict__ second_base,
                              const int64_t* __restrict__ read_base, const int32_t* __restrict__ rowof, int64_t ld,
                              uint8_t* __restrict__ A, uint8_t* __restrict__ R) {
    const int64_t s = blockIdx.x;
    if (s >= total_snps) return;
    const int c = snp_contig[s];
    const int64_t k = s - snp_base[c];  // SNP index within its contig = byte of K
    const int rb = ref_base[s], sb = second_base[s];
    const int64_t rb0 = read_base[c];
    for (int64_t i = snp_off[s] + threadIdx.x; i < snp_off[s + 1]; i += blockDim.x) {
        const int64_t row = rowof[rb0 + read_idx[i]];
        const int cd = code[i];
        if (cd == rb) R[row * ld + k] = 1;  // ref is tested first (:386)
        else if (cd == sb) A[row * ld + k] = 1;
    }
}

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pg_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pg_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pg_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pg_smem(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(pg_smem(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// a wait that cannot hang the device: after ~4 s of spinning the kernel reports and traps
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int32_t* err, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000ll) {
            atomicExch(err, code);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(pg_smem(dst)), "l"(tmap), "r"(pg_smem(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)1 << 16;                    // leading byte offset (not used by swizzled K-major layouts)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version of sm_100
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::i8: u8 x u8 -> s32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(PG_TILE >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pg_smem(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// ---- the contraction ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PG_THREADS, 1)
pair_umma_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapR,
                 const PairWork* __restrict__ work, int n_work, const PairContig* __restrict__ contigs,
                 int32_t* __restrict__ sim, int32_t* __restrict__ diff, int32_t* __restrict__ err) {
    // sim / diff: one 128 x 128 block per work item (pairs_view.cuh)
    extern __shared__ unsigned char pg_raw[];
    // 1024-byte alignment: the swizzle pattern is a function of the shared-memory address
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pg_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PG_STAGES * PG_STAGE_BYTES);
    uint64_t* full = bars;                   // [PG_STAGES] TMA -> MMA
    uint64_t* empty = bars + PG_STAGES;      // [PG_STAGES] MMA -> TMA
    uint64_t* acc_full = bars + 2 * PG_STAGES;   // MMA -> epilogue
    uint64_t* acc_empty = acc_full + 1;          // epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < PG_STAGES; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pg_smem(tmem_slot)),
                     "r"(PG_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const PairWork wk = work[w];
                const int row_i = (int)(contigs[wk.contig].row0) + wk.ti * PG_TILE;
                const int row_j = (int)(contigs[wk.contig].row0) + wk.tj * PG_TILE;
                for (int kb = wk.kb0; kb < wk.kb1; kb++) {
                    mbar_wait(empty + stage, phase ^ 1, err, 1);
                    mbar_arrive_expect_tx(full + stage, PG_STAGE_BYTES);
                    unsigned char* st = smem + stage * PG_STAGE_BYTES;
                    tma_load_2d(st + 0 * PG_SUB, &tmapA, full + stage, kb * PG_BK, row_i);
                    tma_load_2d(st + 1 * PG_SUB, &tmapR, full + stage, kb * PG_BK, row_i);
                    tma_load_2d(st + 2 * PG_SUB, &tmapA, full + stage, kb * PG_BK, row_j);
                    tma_load_2d(st + 3 * PG_SUB, &tmapR, full + stage, kb * PG_BK, row_j);
                    if (++stage == PG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            constexpr uint32_t idesc256 = umma_idesc(256), idesc128 = umma_idesc(128);
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const PairWork wk = work[w];
                mbar_wait(acc_empty, acc_phase ^ 1, err, 2);  // the epilogue has drained the previous tile
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = wk.kb0; kb < wk.kb1; kb++) {
                    mbar_wait(full + stage, phase, err, 3);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = pg_smem(smem + stage * PG_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < PG_BK / 32; k++) {
                        const uint64_t dAi = umma_desc(st + 0 * PG_SUB + 32 * k);
                        const uint64_t dRi = umma_desc(st + 1 * PG_SUB + 32 * k);
                        const uint64_t dAj = umma_desc(st + 2 * PG_SUB + 32 * k);  // N = 256 runs on into R_j
                        const uint64_t dRj = umma_desc(st + 3 * PG_SUB + 32 * k);
                        if (kb == wk.kb0 && k == 0) {
                            umma_i8(tmem + 0, dAi, dAj, idesc256, 0);    // acc0 = Ai Aj^T, acc1 = Ai Rj^T
                            umma_i8(tmem + 128, dRi, dAj, idesc128, 1);  // acc1 += Ri Aj^T
                            umma_i8(tmem + 256, dRi, dRj, idesc128, 0);  // acc2 = Ri Rj^T
                        } else {
                            umma_i8(tmem + 0, dAi, dAj, idesc256, 1);
                            umma_i8(tmem + 128, dRi, dAj, idesc256, 1);
                        }
                    }
                    umma_commit(empty + stage);  // frees the stage once these MMAs have read it
                    if (++stage == PG_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(acc_full);
                acc_phase ^= 1;
            }
        }
    } else {
        // ===== epilogue: 4 warps, warp (w % 4) owns TMEM lanes 32*(w % 4) .. +31 =====
        // Output: the 128 x 128 block of work item w (pairs_view.cuh); thread t owns row t of the tile and stores
        // it as 16-byte pieces. The mirror tile is not materialised: readers swap the indices.
        const int quarter = warp & 3;
        const int t = quarter * 32 + lane;  // row of the tile
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const PairWork wk = work[w];
            const int orow = wk.ti * PG_TILE + t;
            const int ocol0 = wk.tj * PG_TILE;
            mbar_wait(acc_full, acc_phase, err, 4);
            acc_phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            int32_t* const sim_b = sim + (int64_t)w * HS_PV_BLOCK + t * PG_TILE;
            int32_t* const diff_b = diff + (int64_t)w * HS_PV_BLOCK + t * PG_TILE;
            const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int cb = 0; cb < 4; cb++) {
                uint32_t v0[32], v1[32], v2[32];
                tmem_ld32(lane_addr + cb * 32, v0);
                tmem_ld32(lane_addr + 128 + cb * 32, v1);
                tmem_ld32(lane_addr + 256 + cb * 32, v2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb == 3) {
                    // the accumulators are in registers: hand TMEM back before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty);
                }
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const bool dead = (ocol0 + cb * 32 + c == orow);  // :417-432
                    v0[c] = dead ? 0u : 3u * v0[c] + v2[c];
                    v1[c] = dead ? 0u : v1[c];
                }
                int4* ps = reinterpret_cast<int4*>(sim_b + cb * 32);
                int4* pd = reinterpret_cast<int4*>(diff_b + cb * 32);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    ps[c] = make_int4((int)v0[4 * c], (int)v0[4 * c + 1], (int)v0[4 * c + 2], (int)v0[4 * c + 3]);
                    pd[c] = make_int4((int)v1[4 * c], (int)v1[4 * c + 1], (int)v1[4 * c + 2], (int)v1[4 * c + 3]);
                }
            }
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(PG_TMEM_COLS) : "memory");
    }
}

// The dense n x n matrices of one contig in the caller's read order (hsgpu_pairs_fetch: tests and measurements).
__global__ void __launch_bounds__(256) pair_dense_kernel(HsPairView v, int32_t* __restrict__ sim, int32_t* __restrict__ diff) {
    const int r = blockIdx.x;
    for (int k = threadIdx.x; k < v.n; k += blockDim.x) {
        int s, d;
        hs_pair_get(v, r, k, s, d);
        sim[(int64_t)r * v.n + k] = s;
        diff[(int64_t)r * v.n + k] = d;
    }
}

// Plain SIMT statement of the same contraction over the same operands and the same work list (dp4a, one CTA per
// tile pair). Only reachable with HSGPU_PAIRS_SIMT: the A/B check of the tensor-core kernel in the tests.
__global__ void __launch_bounds__(256) pair_simt_kernel(const PairWork* __restrict__ work, const PairContig* __restrict__ contigs,
                                                        int64_t ld_k, const uint8_t* __restrict__ A, const uint8_t* __restrict__ R,
                                                        int32_t* __restrict__ sim, int32_t* __restrict__ diff) {
    const PairWork wk = work[blockIdx.x];
    const PairContig pc = contigs[wk.contig];
    for (int e = threadIdx.x; e < HS_PV_BLOCK; e += blockDim.x) {
        const int i = wk.ti * PG_TILE + (e >> 7), j = wk.tj * PG_TILE + (e & 127);
        const uint32_t* ai = reinterpret_cast<const uint32_t*>(A + (pc.row0 + i) * ld_k);
        const uint32_t* ri = reinterpret_cast<const uint32_t*>(R + (pc.row0 + i) * ld_k);
        const uint32_t* aj = reinterpret_cast<const uint32_t*>(A + (pc.row0 + j) * ld_k);
        const uint32_t* rj = reinterpret_cast<const uint32_t*>(R + (pc.row0 + j) * ld_k);
        unsigned aa = 0, rr = 0, x = 0;
        for (int64_t k = (int64_t)wk.kb0 * (PG_BK / 4); k < (int64_t)wk.kb1 * (PG_BK / 4); k++) {
            const uint32_t a0 = ai[k], r0 = ri[k], a1 = aj[k], r1 = rj[k];
            aa = __dp4a(a0, a1, aa);
            rr = __dp4a(r0, r1, rr);
            x = __dp4a(a0, r1, x);
            x = __dp4a(r0, a1, x);
        }
        const bool dead = i == j;
        sim[(int64_t)blockIdx.x * HS_PV_BLOCK + e] = dead ? 0 : (int)(3 * aa + rr);
        diff[(int64_t)blockIdx.x * HS_PV_BLOCK + e] = dead ? 0 : (int)x;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*PgEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int pg_make_tmap(hsgpu_ctx* ctx, CUtensorMap* tm, void* base, int64_t k_ld, int64_t rows) {
    static std::atomic<PgEncodeTiled> enc_cache{nullptr};  // the same pointer whoever stores it first
    PgEncodeTiled enc = enc_cache.load();
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        HS_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess)
            HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu_pairs: the driver does not export cuTensorMapEncodeTiled");
        enc = (PgEncodeTiled)fn;
        enc_cache.store(enc);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)k_ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k_ld};
    const cuuint32_t box[2] = {PG_BK, PG_TILE};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "hsgpu_pairs: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        HS_FAIL(ctx, HSGPU_ERR_CUDA, buf);
    }
    return HSGPU_OK;
}

extern "C" {

void hsgpu_pairs_destroy(hsgpu_pairs* h) {
    if (!h) return;
    hsgpu_ctx* ctx = h->ctx;
    cudaSetDevice(ctx->device);
    hs_free(ctx, h->d_A); hs_free(ctx, h->d_R); hs_free(ctx, h->d_rowmap); hs_free(ctx, h->d_rowof);
    hs_free(ctx, h->d_rowloc); hs_free(ctx, h->d_has_cells);
    hs_free(ctx, h->d_contigs); hs_free(ctx, h->d_work);
    hs_free(ctx, h->d_sim); hs_free(ctx, h->d_diff); hs_free(ctx, h->d_tilemap); hs_free(ctx, h->d_read_base);
    hs_free(ctx, h->d_err); hs_free(ctx, h->d_snp_off); hs_free(ctx, h->d_snp_base); hs_free(ctx, h->d_read_idx);
    hs_free(ctx, h->d_code); hs_free(ctx, h->d_rb); hs_free(ctx, h->d_sb); hs_free(ctx, h->d_snp_contig);
    hs_stream_sync(ctx);
    delete h;
}

int hsgpu_pairs_create(hsgpu_ctx* ctx, int32_t n_contigs, const int32_t* n_reads, const int64_t* snp_base,
                       const int64_t* snp_off, const uint32_t* read_idx, const uint8_t* code, const uint8_t* ref_base,
                       const uint8_t* second_base, int32_t flags, hsgpu_pairs** out) {
    if (!ctx || !out || n_contigs < 0 || (n_contigs > 0 && (!n_reads || !snp_base))) return HSGPU_ERR_ARG;
    *out = nullptr;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t total_snps = n_contigs > 0 ? snp_base[n_contigs] : 0;
    if (total_snps > 0 && (!snp_off || !ref_base || !second_base)) return HSGPU_ERR_ARG;
    const int64_t n_cells = total_snps > 0 ? snp_off[total_snps] : 0;
    if (n_cells > 0 && (!read_idx || !code)) return HSGPU_ERR_ARG;
    const bool dense = (flags & HSGPU_PAIRS_DENSE) != 0, keep_order = (flags & HSGPU_PAIRS_KEEP_ORDER) != 0;

    hsgpu_pairs* h = new hsgpu_pairs();
    h->ctx = ctx;
    h->n_contigs = n_contigs;
    h->flags = flags;
    h->total_snps = total_snps;
    h->n_cells = n_cells;
    h->h_contigs.resize(n_contigs);
    std::vector<int64_t> read_base(n_contigs + 1, 0);
    int64_t max_snps = 0, rows = 0, map_elems = 0;
    for (int c = 0; c < n_contigs; c++) {
        if (n_reads[c] < 0 || snp_base[c + 1] < snp_base[c]) {
            delete h;
            HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pairs_create: negative read or SNP count");
        }
        read_base[c + 1] = read_base[c] + n_reads[c];
        max_snps = std::max(max_snps, snp_base[c + 1] - snp_base[c]);
        PairContig& pc = h->h_contigs[c];
        pc.n = n_reads[c];
        pc.n_pad = (n_reads[c] + PG_TILE - 1) / PG_TILE * PG_TILE;
        pc.row0 = rows;
        pc.map_off = map_elems;
        pc.read0 = read_base[c];
        rows += pc.n_pad;
        const int64_t nt = pc.n_pad / PG_TILE;
        map_elems += nt * nt;
    }
    if (rows >= (int64_t)1 << 31) {
        delete h;
        HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_pairs_create: more than 2^31 operand rows in one batch");
    }
    h->total_rows = rows;
    h->k_ld = std::max<int64_t>(PG_BK, (max_snps + PG_BK - 1) / PG_BK * PG_BK);
    const int64_t total_reads = read_base[n_contigs];

    // first / last SNP of every read (contig-local SNP indices), then the row order: by first SNP
    std::vector<int32_t> first(total_reads, INT32_MAX), last(total_reads, -1), snp_contig(total_snps);
    for (int c = 0; c < n_contigs; c++) {
        for (int64_t s = snp_base[c]; s < snp_base[c + 1]; s++) {
            snp_contig[s] = c;
            const int32_t k = (int32_t)(s - snp_base[c]);
            if (snp_off[s + 1] < snp_off[s]) {
                delete h;
                HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pairs_create: snp_off is not ascending");
            }
            for (int64_t i = snp_off[s]; i < snp_off[s + 1]; i++) {
                if (read_idx[i] >= (uint32_t)n_reads[c]) {
                    delete h;
                    HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_pairs_create: read index out of range");
                }
                const int64_t r = read_base[c] + read_idx[i];
                if (first[r] == INT32_MAX) first[r] = k;
                last[r] = k;  // SNPs are visited in ascending order
            }
        }
    }
    std::vector<int32_t> rowmap(rows, -1), rowof(std::max<int64_t>(total_reads, 1), 0), rowloc(std::max<int64_t>(total_reads, 1), 0);
    std::vector<uint8_t> has_cells(std::max<int64_t>(total_reads, 1), 0);
    for (int64_t r = 0; r < total_reads; r++) has_cells[r] = last[r] >= 0;
    std::vector<PairWork> work;
    int64_t kb_listed = 0, kb_dense = 0, tiles_dense = 0;
    for (int c = 0; c < n_contigs; c++) {
        PairContig& pc = h->h_contigs[c];
        const int64_t rb0 = read_base[c];
        std::vector<int32_t> order(pc.n);
        std::iota(order.begin(), order.end(), 0);
        if (!keep_order) {
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return first[rb0 + a] < first[rb0 + b]; });
            for (int32_t r = 0; r < pc.n; r++)
                if (order[r] != r) {
                    h->all_identity = false;
                    break;
                }
        }
        for (int32_t r = 0; r < pc.n; r++) {
            rowmap[pc.row0 + r] = order[r];
            rowof[rb0 + order[r]] = (int32_t)(pc.row0 + r);
            rowloc[rb0 + order[r]] = r;
        }
        const int nt = pc.n_pad / PG_TILE;
        const int nkb = (int)((snp_base[c + 1] - snp_base[c] + PG_BK - 1) / PG_BK);
        std::vector<int32_t> klo(nt, INT32_MAX), khi(nt, 0);
        for (int32_t r = 0; r < pc.n; r++) {
            const int64_t g = rb0 + order[r];
            if (last[g] < 0) continue;
            klo[r / PG_TILE] = std::min(klo[r / PG_TILE], first[g] / PG_BK);
            khi[r / PG_TILE] = std::max(khi[r / PG_TILE], last[g] / PG_BK + 1);
        }
        for (int i = 0; i < nt; i++)
            for (int j = i; j < nt; j++) {
                tiles_dense++;
                kb_dense += nkb;
                int k0 = dense ? 0 : std::max(klo[i], klo[j]);
                int k1 = dense ? nkb : std::min(khi[i], khi[j]);
                if (k0 >= k1) continue;
                work.push_back(PairWork{c, i, j, k0, k1});
                kb_listed += k1 - k0;
            }
    }
    // longest items first: the static round-robin over the persistent CTAs then ends evenly
    std::stable_sort(work.begin(), work.end(), [](const PairWork& a, const PairWork& b) { return a.kb1 - a.kb0 > b.kb1 - b.kb0; });
    h->n_work = (int64_t)work.size();
    h->kblocks_listed = kb_listed;
    h->kblocks_dense = kb_dense;
    h->tiles_dense = tiles_dense;
    if ((int64_t)work.size() >= (int64_t)1 << 31) {
        delete h;
        HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_pairs_create: more than 2^31 tile pairs in one batch");
    }
    // the block of every scheduled tile pair = its place in the (sorted) work list
    std::vector<int32_t> tilemap((size_t)std::max<int64_t>(map_elems, 1), -1);
    for (size_t w = 0; w < work.size(); w++) {
        const PairContig& pc = h->h_contigs[work[w].contig];
        tilemap[(size_t)(pc.map_off + (int64_t)work[w].ti * (pc.n_pad / PG_TILE) + work[w].tj)] = (int32_t)w;
    }

#define PG_TRY(call)                                                               \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) {                                                   \
            hsgpu_pairs_destroy(h);                                                \
            return hs_cuda_fail(ctx, _e, #call, __FILE__, __LINE__);               \
        }                                                                          \
    } while (0)
    PG_TRY(hs_alloc(ctx, &h->d_A, rows * h->k_ld));
    PG_TRY(hs_alloc(ctx, &h->d_R, rows * h->k_ld));
    PG_TRY(hs_alloc(ctx, &h->d_rowmap, rows));
    PG_TRY(hs_alloc(ctx, &h->d_rowof, total_reads));
    PG_TRY(hs_alloc(ctx, &h->d_contigs, n_contigs));
    PG_TRY(hs_alloc(ctx, &h->d_work, h->n_work));
    PG_TRY(hs_alloc(ctx, &h->d_sim, h->n_work * HS_PV_BLOCK));
    PG_TRY(hs_alloc(ctx, &h->d_diff, h->n_work * HS_PV_BLOCK));
    PG_TRY(hs_alloc(ctx, &h->d_tilemap, (int64_t)tilemap.size()));
    PG_TRY(hs_alloc(ctx, &h->d_rowloc, total_reads));
    PG_TRY(hs_alloc(ctx, &h->d_has_cells, total_reads));
    PG_TRY(hs_alloc(ctx, &h->d_read_base, n_contigs + 1));
    PG_TRY(hs_h2d(ctx, h->d_tilemap, tilemap.data(), (int64_t)tilemap.size()));
    PG_TRY(hs_h2d(ctx, h->d_rowloc, rowloc.data(), total_reads));
    PG_TRY(hs_h2d(ctx, h->d_has_cells, has_cells.data(), total_reads));
    PG_TRY(hs_h2d(ctx, h->d_read_base, read_base.data(), n_contigs + 1));
    PG_TRY(hs_alloc(ctx, &h->d_err, 1));
    PG_TRY(hs_alloc(ctx, &h->d_snp_off, total_snps + 1));
    PG_TRY(hs_alloc(ctx, &h->d_snp_base, n_contigs + 1));
    PG_TRY(hs_alloc(ctx, &h->d_read_idx, n_cells));
    PG_TRY(hs_alloc(ctx, &h->d_code, n_cells));
    PG_TRY(hs_alloc(ctx, &h->d_rb, total_snps));
    PG_TRY(hs_alloc(ctx, &h->d_sb, total_snps));
    PG_TRY(hs_alloc(ctx, &h->d_snp_contig, total_snps));
    PG_TRY(hs_h2d(ctx, h->d_rowmap, rowmap.data(), rows));
    PG_TRY(hs_h2d(ctx, h->d_rowof, rowof.data(), total_reads));
    PG_TRY(hs_h2d(ctx, h->d_contigs, h->h_contigs.data(), n_contigs));
    PG_TRY(hs_h2d(ctx, h->d_work, work.data(), h->n_work));
    if (total_snps > 0) PG_TRY(hs_h2d(ctx, h->d_snp_off, snp_off, total_snps + 1));
    if (n_contigs > 0) PG_TRY(hs_h2d(ctx, h->d_snp_base, snp_base, n_contigs + 1));
    PG_TRY(hs_h2d(ctx, h->d_read_idx, read_idx, n_cells));
    PG_TRY(hs_h2d(ctx, h->d_code, code, n_cells));
    PG_TRY(hs_h2d(ctx, h->d_rb, ref_base, total_snps));
    PG_TRY(hs_h2d(ctx, h->d_sb, second_base, total_snps));
    PG_TRY(hs_h2d(ctx, h->d_snp_contig, snp_contig.data(), total_snps));
    PG_TRY(cudaMemsetAsync(h->d_A, 0, (size_t)std::max<int64_t>(rows * h->k_ld, 1), ctx->stream));
    PG_TRY(cudaMemsetAsync(h->d_R, 0, (size_t)std::max<int64_t>(rows * h->k_ld, 1), ctx->stream));
    PG_TRY(cudaMemsetAsync(h->d_err, 0, sizeof(int32_t), ctx->stream));
    if (total_snps > 0) {
        if (ctx->profiling) hs_prof_begin(ctx, "onehot_kernel");
        onehot_kernel<<<(unsigned)total_snps, 128, 0, ctx->stream>>>(total_snps, h->d_snp_contig, h->d_snp_base, h->d_snp_off,
                                                                      h->d_read_idx, h->d_code, h->d_rb, h->d_sb, h->d_read_base,
                                                                      h->d_rowof, h->k_ld, h->d_A, h->d_R);
        if (ctx->profiling) hs_prof_end(ctx);
        ctx->launches++;
        PG_TRY(cudaGetLastError());
    }
    // host vectors above are pageable: the copies must have left them before we return
    PG_TRY(hs_stream_sync(ctx));
#undef PG_TRY
    if (rows > 0) {
        int rc = pg_make_tmap(ctx, &h->tmapA, h->d_A, h->k_ld, rows);
        if (!rc) rc = pg_make_tmap(ctx, &h->tmapR, h->d_R, h->k_ld, rows);
        if (rc) {
            hsgpu_pairs_destroy(h);
            return rc;
        }
    }
    *out = h;
    return HSGPU_OK;
}

int hsgpu_pairs_compute(hsgpu_pairs* h) {
    if (!h) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = h->ctx;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    h->computed = false;
    if (h->n_work == 0) {
        h->computed = true;
        return HSGPU_OK;
    }
    if (h->flags & HSGPU_PAIRS_SIMT) {
        HS_KERNEL(ctx, "pair_simt_kernel", pair_simt_kernel<<<(unsigned)h->n_work, 256, 0, ctx->stream>>>(
            h->d_work, h->d_contigs, h->k_ld, h->d_A, h->d_R, h->d_sim, h->d_diff));
        h->computed = true;
        return HSGPU_OK;
    }
    static std::atomic<bool> attr_set[64];  // per device: function attributes belong to the device's context
    if (!attr_set[ctx->device & 63]) {
        HS_CUDA(ctx, cudaFuncSetAttribute(pair_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PG_SMEM));
        attr_set[ctx->device & 63] = true;
    }
    const int grid = (int)std::min<int64_t>(h->n_work, ctx->sm_count);
    HS_KERNEL(ctx, "pair_umma_kernel", pair_umma_kernel<<<grid, PG_THREADS, PG_SMEM, ctx->stream>>>(
        h->tmapA, h->tmapR, h->d_work, (int)h->n_work, h->d_contigs, h->d_sim, h->d_diff, h->d_err));
    h->computed = true;
    return HSGPU_OK;
}

static HsPairView pg_view(const hsgpu_pairs* h, int contig) {
    const PairContig& pc = h->h_contigs[contig];
    HsPairView v;
    v.row_of = h->d_rowloc + pc.read0;
    v.tilemap = h->d_tilemap + pc.map_off;
    v.sim = h->d_sim;
    v.diff = h->d_diff;
    v.has_cells = h->d_has_cells + pc.read0;
    v.nt = pc.n_pad / PG_TILE;
    v.n = pc.n;
    return v;
}

int hsgpu_pairs_fetch(hsgpu_pairs* h, int32_t contig, int32_t* sim, int32_t* diff) {
    if (!h || contig < 0 || contig >= h->n_contigs) return HSGPU_ERR_ARG;
    hsgpu_ctx* ctx = h->ctx;
    if (!h->computed) HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_pairs_fetch: call hsgpu_pairs_compute first");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const PairContig& pc = h->h_contigs[contig];
    int32_t *d_s = nullptr, *d_d = nullptr;
    HsTemps temps(ctx);
    temps.own(d_s, d_d);
    if (pc.n > 0 && (sim || diff)) {
        // the dense n x n form exists only here, for the caller who asks for it (the later stages of the library read
        // the blocks): it is built on the device and copied out
        if (pc.n > 46000) HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_pairs_fetch: dense n x n output limited to 46000 reads per contig");
        const int64_t nn = (int64_t)pc.n * pc.n;
        HS_CUDA(ctx, hs_alloc(ctx, &d_s, nn));
        HS_CUDA(ctx, hs_alloc(ctx, &d_d, nn));
        HS_KERNEL(ctx, "pair_dense_kernel", pair_dense_kernel<<<pc.n, 256, 0, ctx->stream>>>(pg_view(h, contig), d_s, d_d));
        if (sim) HS_CUDA(ctx, hs_d2h(ctx, sim, d_s, nn));
        if (diff) HS_CUDA(ctx, hs_d2h(ctx, diff, d_d, nn));
    }
    int32_t err = 0;
    HS_CUDA(ctx, hs_d2h(ctx, &err, h->d_err, 1));
    HS_CUDA(ctx, hs_stream_sync(ctx));
    hs_free(ctx, d_s);
    hs_free(ctx, d_d);
    if (err) HS_FAIL(ctx, HSGPU_ERR_CUDA, "hsgpu_pairs: the tensor-core kernel timed out on a barrier");
    return HSGPU_OK;
}

int hsgpu_pairs_info(hsgpu_pairs* h, int64_t* info) {
    if (!h || !info) return HSGPU_ERR_ARG;
    info[0] = h->n_work;          // tile pairs scheduled
    info[1] = h->tiles_dense;     // tile pairs of the full upper triangles
    info[2] = h->kblocks_listed;  // 128-SNP blocks executed (each = 2 x 4 MMAs of 128x256x32)
    info[3] = h->kblocks_dense;   // the same for the full upper triangles over all SNPs
    info[4] = h->total_rows;
    info[5] = h->k_ld;
    info[6] = h->n_work * HS_PV_BLOCK;  // elements stored per matrix: one block per scheduled tile pair
    info[7] = h->all_identity ? 1 : 0;
    return HSGPU_OK;
}

int hsgpu_read_pair_counts(hsgpu_ctx* ctx, int32_t n_reads, int32_t n_snps, const int64_t* snp_off,
                           const uint32_t* read_idx, const uint8_t* code, const uint8_t* ref_base,
                           const uint8_t* second_base, int32_t* sim, int32_t* diff) {
    if (!ctx || n_reads < 0 || n_snps < 0 || !sim || !diff) return HSGPU_ERR_ARG;
    if (n_reads == 0) return HSGPU_OK;
    const int64_t base[2] = {0, n_snps};
    hsgpu_pairs* h = nullptr;
    int rc = hsgpu_pairs_create(ctx, 1, &n_reads, base, snp_off, read_idx, code, ref_base, second_base, 0, &h);
    if (rc) return rc;
    rc = hsgpu_pairs_compute(h);
    if (!rc) rc = hsgpu_pairs_fetch(h, 0, sim, diff);
    hsgpu_pairs_destroy(h);
    return rc;
}

}  // extern "C"

// internal view for the read-graph stage (graph.cu): the device-resident result blocks of one contig
int hs_pairs_contigs(hsgpu_pairs* h) { return h ? h->n_contigs : 0; }
int hs_pairs_view(hsgpu_pairs* h, int32_t contig, hsgpu_ctx** ctx, HsPairView* view) {
    if (!h || contig < 0 || contig >= h->n_contigs) return HSGPU_ERR_ARG;
    *ctx = h->ctx;
    if (!h->computed) return HSGPU_ERR_STATE;
    *view = pg_view(h, contig);
    return HSGPU_OK;
}
