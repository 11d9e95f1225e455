// Integer-pipe micro-benchmark: the denominators of the realignment roofline (SURVEY.md 8d asks for a measured
// INT32 / logic peak; MEASURED_PEAKS.json only has HBM and bf16). Not part of include/hsgpu.h; called by
// scripts/peaks_int.py, whose output is committed as profiles/peaks_int.json.
//
// Every thread runs CHAINS independent dependency chains of one instruction kind for `iters` trips; the result
// is thread-instructions per second over the whole chip:
//   kind 0  lop3.b32          (ALU pipe)      -- what the Myers recurrences are made of
//   kind 1  add.u32 (IADD3)   (ALU pipe)
//   kind 2  mad.lo.u32 (IMAD) (FMA pipe)
//   kind 3  lop3 + mad.lo alternating (both pipes: the most a warp scheduler can issue, 1 instruction / clock)
//   kind 4  shf.l.wrap (funnel shift, ALU pipe)
#include "common.cuh"

#define PK_CHAINS 8

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t seed, int iters, uint32_t* __restrict__ out) {
    uint32_t v[PK_CHAINS];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) v[c] = seed * (2 * c + 1) + t;
    uint32_t a = seed | 1u, b = ~seed;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < PK_CHAINS; c++) {
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(a), "r"(b));
                if (KIND == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(a));
                if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(a), "r"(b));
                if (KIND == 3) {
                    if (c & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(a), "r"(b));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(a), "r"(b));
                }
                if (KIND == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(a), "r"(b));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < PK_CHAINS; c++) r ^= v[c];
    if (r == 0x12345678u) out[t & 1023] = r;  // keeps the chains alive
}

extern "C" int hsgpu_debug_int_peak(hsgpu_ctx* ctx, int kind, int iters, double* thread_ops_per_s, double* ms_out) {
    if (!ctx || !thread_ops_per_s) return HSGPU_ERR_ARG;
    HS_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* d_out = nullptr;
    HS_CUDA(ctx, hs_alloc(ctx, &d_out, 1024));
    const int ctas = ctx->sm_count * 8, threads = 256;  // 2048 threads per SM: every scheduler has 16 warps to pick from
    cudaEvent_t e0, e1;
    HS_CUDA(ctx, cudaEventCreate(&e0));
    HS_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {  // first trip warms the clocks up
        HS_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        switch (kind) {
            case 0: int_peak_kernel<0><<<ctas, threads, 0, ctx->stream>>>(12345u + rep, iters, d_out); break;
            case 1: int_peak_kernel<1><<<ctas, threads, 0, ctx->stream>>>(12345u + rep, iters, d_out); break;
            case 2: int_peak_kernel<2><<<ctas, threads, 0, ctx->stream>>>(12345u + rep, iters, d_out); break;
            case 3: int_peak_kernel<3><<<ctas, threads, 0, ctx->stream>>>(12345u + rep, iters, d_out); break;
            case 4: int_peak_kernel<4><<<ctas, threads, 0, ctx->stream>>>(12345u + rep, iters, d_out); break;
            default: return HSGPU_ERR_ARG;
        }
        HS_LAUNCH_CHECK(ctx);
        HS_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        HS_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        HS_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    hs_free(ctx, d_out);
    const double ops = (double)ctas * threads * (double)iters * 8.0 * PK_CHAINS;
    *thread_ops_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return HSGPU_OK;
}
