// Read x read SNP agreement counts: list_similarities_and_differences_between_reads3 (reference
// src/separate_reads.cpp:374-433), the one dense contraction of the pipeline.
//
//   A[r,s] = 1 iff read r carries second_base at SNP s, R[r,s] = 1 iff it carries ref_base (:384-394)
//   similarity = 3*A*At + R*Rt,   difference = A*Rt + R*At,   diagonals zeroed (:414-432)
//
// Both products share the left operand U = [A | R] (n x 2S, int8, K-contiguous):
//   similarity = U * [3A | R]^T,   difference = U * [R | A]^T
// so one kernel streams U once and feeds two int8 IMMA accumulators (mma.sync.m16n8k32.s8, exact
// int32 accumulation), one 128x128 output tile of each matrix per CTA.
#include "common.cuh"

#define PG_BM 128
#define PG_BN 128
#define PG_BK 64
#define PG_LD 80  // padded row stride in bytes (20 words: conflict-free fragment loads)

__global__ void onehot_kernel(int n_snps, const int64_t* __restrict__ snp_off, const uint32_t* __restrict__ read_idx,
                              const uint8_t* __restrict__ code, const uint8_t* __restrict__ ref_base,
                              const uint8_t* __restrict__ second_base, int s_pad, int64_t ld, int8_t* __restrict__ U,
                              int8_t* __restrict__ Vs, int8_t* __restrict__ Vd) {
    const int s = blockIdx.x;
    if (s >= n_snps) return;
    const int rb = ref_base[s], sb = second_base[s];
    for (int64_t i = snp_off[s] + threadIdx.x; i < snp_off[s + 1]; i += blockDim.x) {
        const int64_t r = read_idx[i];
        const int c = code[i];
        if (c == rb) {  // ref is tested first (:386)
            U[r * ld + s_pad + s] = 1;
            Vs[r * ld + s_pad + s] = 1;
            Vd[r * ld + s] = 1;
        } else if (c == sb) {
            U[r * ld + s] = 1;
            Vs[r * ld + s] = 3;
            Vd[r * ld + s_pad + s] = 1;
        }
    }
}

__device__ __forceinline__ void mma_s8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C tiles: sim[bm.., bn..] and diff[bm.., bn..]; 8 warps as 4 (m) x 2 (n), warp tile 32 x 64
__global__ void __launch_bounds__(256) pair_gemm_kernel(int n, int64_t ld, int K, const int8_t* __restrict__ U,
                                                        const int8_t* __restrict__ Vs, const int8_t* __restrict__ Vd,
                                                        int32_t* __restrict__ sim, int32_t* __restrict__ diff) {
    __shared__ __align__(16) unsigned char sU[PG_BM * PG_LD];
    __shared__ __align__(16) unsigned char sS[PG_BN * PG_LD];
    __shared__ __align__(16) unsigned char sD[PG_BN * PG_LD];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid & 3, wn = wid >> 2;
    const int g = lane >> 2, t = lane & 3;
    const int bm = blockIdx.y * PG_BM, bn = blockIdx.x * PG_BN;
    int accS[2][8][4], accD[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int k = 0; k < 4; k++) accS[i][j][k] = accD[i][j][k] = 0;

    for (int k0 = 0; k0 < K; k0 += PG_BK) {
        __syncthreads();
        // 128 rows x 64 bytes = 512 uint4 per operand; 256 threads -> 2 each
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int v = tid + 256 * i;
            const int row = v >> 2, part = v & 3;
            const uint4 u = *reinterpret_cast<const uint4*>(U + (int64_t)(bm + row) * ld + k0 + 16 * part);
            const uint4 s = *reinterpret_cast<const uint4*>(Vs + (int64_t)(bn + row) * ld + k0 + 16 * part);
            const uint4 d = *reinterpret_cast<const uint4*>(Vd + (int64_t)(bn + row) * ld + k0 + 16 * part);
            *reinterpret_cast<uint4*>(sU + row * PG_LD + 16 * part) = u;
            *reinterpret_cast<uint4*>(sS + row * PG_LD + 16 * part) = s;
            *reinterpret_cast<uint4*>(sD + row * PG_LD + 16 * part) = d;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PG_BK; kk += 32) {
            uint32_t af[2][4];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const unsigned char* base = sU + (wm * 32 + i * 16 + g) * PG_LD + kk + 4 * t;
                af[i][0] = *reinterpret_cast<const uint32_t*>(base);
                af[i][1] = *reinterpret_cast<const uint32_t*>(base + 8 * PG_LD);
                af[i][2] = *reinterpret_cast<const uint32_t*>(base + 16);
                af[i][3] = *reinterpret_cast<const uint32_t*>(base + 8 * PG_LD + 16);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int off = (wn * 64 + j * 8 + g) * PG_LD + kk + 4 * t;
                uint32_t bs[2], bd[2];
                bs[0] = *reinterpret_cast<const uint32_t*>(sS + off);
                bs[1] = *reinterpret_cast<const uint32_t*>(sS + off + 16);
                bd[0] = *reinterpret_cast<const uint32_t*>(sD + off);
                bd[1] = *reinterpret_cast<const uint32_t*>(sD + off + 16);
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    mma_s8(accS[i][j], af[i], bs);
                    mma_s8(accD[i][j], af[i], bd);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = bm + wm * 32 + i * 16 + g + 8 * h;
                const int col = bn + wn * 64 + j * 8 + 2 * t;
                if (row < n) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        if (col + e < n) {
                            const bool diag = row == col + e;  // :417-432
                            sim[(int64_t)row * n + col + e] = diag ? 0 : accS[i][j][2 * h + e];
                            diff[(int64_t)row * n + col + e] = diag ? 0 : accD[i][j][2 * h + e];
                        }
                    }
                }
            }
}

extern "C" int hsgpu_read_pair_counts(hsgpu_ctx* ctx, int32_t n_reads, int32_t n_snps, const int64_t* snp_off,
                                      const uint32_t* read_idx, const uint8_t* code, const uint8_t* ref_base,
                                      const uint8_t* second_base, int32_t* sim, int32_t* diff) {
    if (!ctx || n_reads < 0 || n_snps < 0 || !sim || !diff) return HSGPU_ERR_ARG;
    if (n_reads == 0) return HSGPU_OK;
    if (n_reads > 46000) HS_FAIL(ctx, HSGPU_ERR_LIMIT, "hsgpu_read_pair_counts: dense n x n output limited to 46000 reads");
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n_cells = n_snps > 0 ? snp_off[n_snps] : 0;
    for (int64_t i = 0; i < n_cells; i++)
        if (read_idx[i] >= (uint32_t)n_reads) HS_FAIL(ctx, HSGPU_ERR_ARG, "hsgpu_read_pair_counts: read index out of range");
    const int n_pad = (n_reads + PG_BM - 1) / PG_BM * PG_BM;
    const int s_pad = (n_snps + 31) / 32 * 32;
    const int K = 2 * s_pad;
    const int64_t ld = K > 0 ? K : 64;
    int8_t *U = nullptr, *Vs = nullptr, *Vd = nullptr;
    int32_t *d_sim = nullptr, *d_diff = nullptr;
    int64_t* d_off = nullptr;
    uint32_t* d_idx = nullptr;
    uint8_t *d_code = nullptr, *d_rb = nullptr, *d_sb = nullptr;
    const int64_t opbytes = (int64_t)n_pad * ld;
    HS_CUDA(ctx, hs_alloc(ctx, &U, opbytes));
    HS_CUDA(ctx, hs_alloc(ctx, &Vs, opbytes));
    HS_CUDA(ctx, hs_alloc(ctx, &Vd, opbytes));
    HS_CUDA(ctx, hs_alloc(ctx, &d_sim, (int64_t)n_reads * n_reads));
    HS_CUDA(ctx, hs_alloc(ctx, &d_diff, (int64_t)n_reads * n_reads));
    HS_CUDA(ctx, cudaMemsetAsync(U, 0, opbytes, ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(Vs, 0, opbytes, ctx->stream));
    HS_CUDA(ctx, cudaMemsetAsync(Vd, 0, opbytes, ctx->stream));
    if (n_snps > 0) {
        HS_CUDA(ctx, hs_alloc(ctx, &d_off, n_snps + 1));
        HS_CUDA(ctx, hs_alloc(ctx, &d_idx, n_cells));
        HS_CUDA(ctx, hs_alloc(ctx, &d_code, n_cells));
        HS_CUDA(ctx, hs_alloc(ctx, &d_rb, n_snps));
        HS_CUDA(ctx, hs_alloc(ctx, &d_sb, n_snps));
        HS_CUDA(ctx, hs_h2d(ctx, d_off, snp_off, n_snps + 1));
        HS_CUDA(ctx, hs_h2d(ctx, d_idx, read_idx, n_cells));
        HS_CUDA(ctx, hs_h2d(ctx, d_code, code, n_cells));
        HS_CUDA(ctx, hs_h2d(ctx, d_rb, ref_base, n_snps));
        HS_CUDA(ctx, hs_h2d(ctx, d_sb, second_base, n_snps));
        HS_KERNEL(ctx, "onehot_kernel", onehot_kernel<<<n_snps, 128, 0, ctx->stream>>>(n_snps, d_off, d_idx, d_code, d_rb, d_sb, s_pad, ld, U, Vs, Vd));
    }
    dim3 grid(n_pad / PG_BN, n_pad / PG_BM);
    HS_KERNEL(ctx, "pair_gemm_kernel", pair_gemm_kernel<<<grid, 256, 0, ctx->stream>>>(n_reads, ld, K, U, Vs, Vd, d_sim, d_diff));
    HS_CUDA(ctx, hs_d2h(ctx, sim, d_sim, (int64_t)n_reads * n_reads));
    HS_CUDA(ctx, hs_d2h(ctx, diff, d_diff, (int64_t)n_reads * n_reads));
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    hs_free(ctx, U); hs_free(ctx, Vs); hs_free(ctx, Vd); hs_free(ctx, d_sim); hs_free(ctx, d_diff);
    hs_free(ctx, d_off); hs_free(ctx, d_idx); hs_free(ctx, d_code); hs_free(ctx, d_rb); hs_free(ctx, d_sb);
    return HSGPU_OK;
}
