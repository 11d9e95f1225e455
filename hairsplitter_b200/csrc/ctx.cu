// Context management, host-side packing helpers and device prefix sums.
#include <cstdio>
#include <cstring>

#include <chrono>
#include <mutex>

#include "common.cuh"
#include "rank.cuh"

static thread_local std::string g_tls_error;

void hs_set_error(hsgpu_ctx* ctx, const std::string& msg) {
    g_tls_error = msg;
    if (ctx) ctx->err = msg;
}

int hs_cuda_fail(hsgpu_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    hs_set_error(ctx, buf);
    return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? HSGPU_ERR_NO_DEVICE : HSGPU_ERR_CUDA;
}

void hs_prof_begin(hsgpu_ctx* ctx, const char* name) {
    HsProfEntry e;
    e.name = name;
    cudaEventCreate(&e.begin);
    cudaEventCreate(&e.end);
    cudaEventRecord(e.begin, ctx->stream);
    ctx->prof.push_back(e);
}
void hs_prof_end(hsgpu_ctx* ctx) { cudaEventRecord(ctx->prof.back().end, ctx->stream); }

extern "C" {

int hsgpu_profile_enable(hsgpu_ctx* ctx, int on) {
    if (!ctx) return HSGPU_ERR_ARG;
    ctx->profiling = on != 0;
    return HSGPU_OK;
}

// "name\tlaunches\ttotal_ms\n" per kernel, accumulated since the last report; clears the log
const char* hsgpu_profile_report(hsgpu_ctx* ctx) {
    if (!ctx) return "";
    cudaStreamSynchronize(ctx->stream);
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<int64_t> cnt;
    for (auto& e : ctx->prof) {
        float t = 0;
        cudaEventElapsedTime(&t, e.begin, e.end);
        cudaEventDestroy(e.begin);
        cudaEventDestroy(e.end);
        size_t i = 0;
        for (; i < names.size(); i++)
            if (names[i] == e.name) break;
        if (i == names.size()) {
            names.push_back(e.name);
            ms.push_back(0);
            cnt.push_back(0);
        }
        ms[i] += t;
        cnt[i]++;
    }
    ctx->prof.clear();
    ctx->prof_report.clear();
    for (size_t i = 0; i < names.size(); i++) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s\t%lld\t%.6f\n", names[i].c_str(), (long long)cnt[i], ms[i]);
        ctx->prof_report += buf;
    }
    return ctx->prof_report.c_str();
}

// HSGPU_TIMING=1: where the time of hsgpu_ctx_create goes (almost all of it is the driver's)
static void ctx_lap(const char* what, std::chrono::steady_clock::time_point& t0) {
    static const bool on = getenv("HSGPU_TIMING") != nullptr;
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[hsgpu timing] ctx_create: %-34s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
}

int hsgpu_ctx_create(int device, hsgpu_ctx** out) {
    if (!out) return HSGPU_ERR_ARG;
    *out = nullptr;
    auto lap = std::chrono::steady_clock::now();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    ctx_lap("cudaGetDeviceCount (cuInit)", lap);
    if (e != cudaSuccess || n == 0) {
        hs_set_error(nullptr, "libhsgpu: no CUDA device available (this library has no CPU fallback)");
        return HSGPU_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        hs_set_error(nullptr, "libhsgpu: device index out of range");
        return HSGPU_ERR_ARG;
    }
    cudaDeviceProp prop;
    HS_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    ctx_lap("cudaGetDeviceProperties", lap);
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof(buf), "libhsgpu: device %d (%.64s) is sm_%d%d; kernels are built for sm_100a only", device,
                 prop.name, prop.major, prop.minor);
        hs_set_error(nullptr, buf);
        return HSGPU_ERR_NO_DEVICE;
    }
    HS_CUDA(nullptr, cudaSetDevice(device));
    hsgpu_ctx* ctx = new hsgpu_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    ctx_lap("cudaSetDevice + stream (primary context)", lap);
    if (e != cudaSuccess) {
        delete ctx;
        return hs_cuda_fail(nullptr, e, "cudaStreamCreate", __FILE__, __LINE__);
    }
    // the context's own stream-ordered pool; freed blocks stay in it instead of going back to the driver
    // (HSGPU_SHARED_POOL=1: the device's default pool, for A/B measurements)
    {
        uint64_t thr = UINT64_MAX;
        static const bool shared_pool = getenv("HSGPU_SHARED_POOL") && atoi(getenv("HSGPU_SHARED_POOL")) != 0;
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (!shared_pool && cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
            cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thr);
        } else {
            ctx->pool = nullptr;
            cudaGetLastError();
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    {
        static HsRankLut lut;  // identical for every context; contexts may be created from several threads at once
        static std::once_flag lut_once;
        std::call_once(lut_once, []() { hs_build_rank_lut(lut); });
        ctx_lap("memory pool + rank table (host)", lap);
        e = cudaMalloc(&ctx->d_rank_lut, sizeof(HsRankLut));
        if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rank_lut, &lut, sizeof(HsRankLut), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaStreamDestroy(ctx->stream);
            delete ctx;
            return hs_cuda_fail(nullptr, e, "rank table upload", __FILE__, __LINE__);
        }
    }
    ctx_lap("cudaMalloc + cudaMemcpy (rank table)", lap);
    e = cudaHostAlloc((void**)&ctx->h_scratch, 64, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->scratch_event, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        hsgpu_ctx_destroy(ctx);
        return hs_cuda_fail(nullptr, e, "pinned scratch", __FILE__, __LINE__);
    }
    {
        const char* w = getenv("HSGPU_WAIT");
        if (w && !strcmp(w, "block")) {
            if (cudaEventCreateWithFlags(&ctx->wait_event, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) {
                ctx->wait_event = nullptr;
                cudaGetLastError();
            }
        }
    }
    ctx_lap("cudaHostAlloc + event", lap);
    *out = ctx;
    return HSGPU_OK;
}

void hsgpu_ctx_destroy(hsgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch_event) cudaEventDestroy(ctx->scratch_event);
    if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaStreamDestroy(ctx->stream);
    if (ctx->d_rank_lut) cudaFree(ctx->d_rank_lut);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

const char* hsgpu_last_error(hsgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_tls_error.c_str(); }

int hsgpu_sync(hsgpu_ctx* ctx) {
    if (!ctx) return HSGPU_ERR_ARG;
    HS_CUDA(ctx, hs_stream_sync(ctx));
    return HSGPU_OK;
}

int64_t hsgpu_launch_count(hsgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* hsgpu_stream(hsgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int hsgpu_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return HSGPU_ERR_ARG;
    cudaError_t e = cudaHostAlloc(out, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault);
    if (e != cudaSuccess) return hs_cuda_fail(nullptr, e, "cudaHostAlloc", __FILE__, __LINE__);
    return HSGPU_OK;
}
void hsgpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void hsgpu_pack_bases_ascii(const char* seq, int64_t n, uint32_t* out) {
    int64_t nw = (n + 15) / 16;
    for (int64_t w = 0; w < nw; w++) {
        uint32_t v = 0;
        int64_t lim = n - w * 16 < 16 ? n - w * 16 : 16;
        for (int64_t j = 0; j < lim; j++) {
            uint32_t b;
            switch (seq[w * 16 + j]) {  // src/sequence.cpp:16-21: anything but A, C, G packs as T
                case 'A': b = 0; break;
                case 'C': b = 1; break;
                case 'G': b = 2; break;
                default: b = 3; break;
            }
            v |= b << (2 * j);
        }
        out[w] = v;
    }
}

void hsgpu_pack_bases_codes(const uint8_t* codes, int64_t n, uint32_t* out) {
    int64_t nw = (n + 15) / 16;
    for (int64_t w = 0; w < nw; w++) {
        uint32_t v = 0;
        int64_t lim = n - w * 16 < 16 ? n - w * 16 : 16;
        for (int64_t j = 0; j < lim; j++) v |= (uint32_t)(codes[w * 16 + j] & 3) << (2 * j);
        out[w] = v;
    }
}

int64_t hsgpu_parse_cigar(const char* cigar, uint32_t* out, int64_t capacity) {
    if (!cigar) return HSGPU_ERR_ARG;
    if (cigar[0] == '*' && cigar[1] == 0) return 0;  // src/tools.cpp:29-31
    static const char* letters = "MIDNSHP=X";
    int64_t n = 0;
    uint64_t num = 0;
    bool have = false;
    for (const char* c = cigar; *c; c++) {
        if (*c >= '0' && *c <= '9') {
            num = num * 10 + (uint64_t)(*c - '0');
            have = true;
            if (num >= (1ull << 28)) return HSGPU_ERR_LIMIT;
        } else {
            const char* l = strchr(letters, *c);
            if (!l || !have) return HSGPU_ERR_ARG;
            if (n >= capacity) return HSGPU_ERR_CAPACITY;
            out[n++] = (uint32_t)(num << 4) | (uint32_t)(l - letters);
            num = 0;
            have = false;
        }
    }
    return n;
}

}  // extern "C"

void* hs_host_stage(hsgpu_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_stage_bytes) return ctx->h_stage;
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    size_t want = 1 << 16;
    while (want < bytes) want <<= 1;
    if (cudaHostAlloc(&ctx->h_stage, want, cudaHostAllocDefault) != cudaSuccess) {
        ctx->h_stage = nullptr;
        return nullptr;
    }
    ctx->h_stage_bytes = want;
    return ctx->h_stage;
}

extern "C" {

int64_t hsgpu_pack_cigar8(const uint32_t* ops, int64_t n_ops, uint8_t* out, int64_t capacity) {
    if (!ops && n_ops > 0) return HSGPU_ERR_ARG;
    // BAM op -> kind: M I D N S H P = X
    static const int8_t kind_of[16] = {0, 1, 2, -1, 3, 3, -1, 0, 0, -1, -1, -1, -1, -1, -1, -1};
    int64_t n = 0;
    for (int64_t i = 0; i < n_ops; i++) {
        const int kind = kind_of[ops[i] & 15u];
        if (kind < 0) return HSGPU_ERR_ARG;
        uint32_t len = ops[i] >> 4;
        do {  // a zero-length op stays one (empty) op
            const uint32_t piece = len > 63u ? 63u : len;
            if (out) {
                if (n >= capacity) return HSGPU_ERR_CAPACITY;
                out[n] = (uint8_t)((piece << 2) | (uint32_t)kind);
            }
            n++;
            len -= piece;
        } while (len > 0);
    }
    return n;
}

}  // extern "C"

// ---- device exclusive scan ----------------------------------------------------------------------
// Three-phase scan with 1024-thread blocks handling 4096 elements each; recursion over block sums.
#define SCAN_THREADS 1024
#define SCAN_ITEMS 4
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_kernel(const TIn* __restrict__ in, int64_t* __restrict__ out,
                                                                  int64_t n, int64_t* __restrict__ block_sums) {
    __shared__ long long warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)tid * SCAN_ITEMS;
    long long v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? (long long)in[base + i] : 0;
        s += v[i];
    }
    long long incl = hs_warp_incl_scan64(s, lane);
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        long long w = warp_tot[lane];
        long long wi = hs_warp_incl_scan64(w, lane);
        warp_tot[lane] = wi - w;
        if (lane == 31 && block_sums) block_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    long long run = warp_tot[wid] + incl - s;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void scan_add_kernel(int64_t* __restrict__ out, int64_t n, const int64_t* __restrict__ block_off) {
    int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    long long add = block_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++, i += SCAN_THREADS)
        if (i < n) out[i] += add;
}

__global__ void scan_total_kernel(const int64_t* block_off, const int64_t* block_sums, int64_t nb, int64_t* total) {
    *total = block_off[nb - 1] + block_sums[nb - 1];
}

// Short inputs (every per-read / per-tile table of a batch of contig chunks): ONE launch of one CTA -- 1 driver call
// instead of the 13 of the recursive scan. Each of the 32 warps owns a contiguous segment: it sums it (loads fully
// pipelined, no barrier), the 32 totals are scanned once, then the warp walks its segment again 32 elements at a
// time (warp scan + running carry, next row already requested). Two barriers in all; out may alias in.
#define SCAN_SINGLE_MAX (64 * 1024)
template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_single_kernel(const TIn* in, int64_t* out, int64_t n,
                                                                   int64_t* __restrict__ total) {
    __shared__ long long s_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t per = (((n + 31) / 32) + 31) & ~(int64_t)31;  // elements per warp, a multiple of 32
    const int64_t b = (int64_t)wid * per;
    const int64_t e = b + per < n ? b + per : n;
    long long acc = 0;
#pragma unroll 8
    for (int64_t k = b + lane; k < e; k += 32) acc += (long long)in[k];
    acc = hs_warp_sum64(acc);
    if (lane == 0) s_tot[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        const long long w = s_tot[lane];
        const long long wi = hs_warp_incl_scan64(w, lane);
        s_tot[lane] = wi - w;
        if (lane == 31 && total) *total = wi;
    }
    __syncthreads();
    long long carry = s_tot[wid];
    // eight rows in flight: the loop is a chain of global-load latencies otherwise (one CTA, nothing else to run)
    constexpr int RB = 8;
    for (int64_t k0 = b; k0 < e; k0 += 32 * RB) {
        long long v[RB];
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const int64_t k = k0 + 32 * r + lane;
            v[r] = k < e ? (long long)in[k] : 0;  // all read before any of these rows is written (out may alias in)
        }
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const int64_t k = k0 + 32 * r + lane;
            const long long incl = hs_warp_incl_scan64(v[r], lane);
            if (k < e) out[k] = carry + incl - v[r];
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

template <typename TIn>
static int scan_impl(hsgpu_ctx* ctx, const TIn* in, int64_t* out, int64_t n, int64_t* total) {
    if (n <= 0) {
        if (total) HS_CUDA(ctx, cudaMemsetAsync(total, 0, sizeof(int64_t), ctx->stream));
        return HSGPU_OK;
    }
    if (n <= SCAN_SINGLE_MAX) {
        HS_KERNEL(ctx, "scan_single_kernel<TIn>", scan_single_kernel<TIn><<<1, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, total));
        return HSGPU_OK;
    }
    int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    int64_t* sums = nullptr;
    int64_t* offs = nullptr;
    HS_CUDA(ctx, hs_alloc(ctx, &sums, nb));
    HS_CUDA(ctx, hs_alloc(ctx, &offs, nb));
    HS_KERNEL(ctx, "scan_block_kernel<TIn>", scan_block_kernel<TIn><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, sums));
    if (nb > 1) {
        int rc = scan_impl<int64_t>(ctx, sums, offs, nb, nullptr);
        if (rc) return rc;
        HS_KERNEL(ctx, "scan_add_kernel", scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(out, n, offs));
    } else {
        HS_CUDA(ctx, cudaMemsetAsync(offs, 0, sizeof(int64_t), ctx->stream));
    }
    if (total) {
        HS_KERNEL(ctx, "scan_total_kernel", scan_total_kernel<<<1, 1, 0, ctx->stream>>>(offs, sums, nb, total));
    }
    hs_free(ctx, sums);
    hs_free(ctx, offs);
    return HSGPU_OK;
}

int hs_exclusive_scan_i64(hsgpu_ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total) {
    return scan_impl<int64_t>(ctx, in, out, n, total);
}
int hs_exclusive_scan_u32_to_i64(hsgpu_ctx* ctx, const uint32_t* in, int64_t* out, int64_t n, int64_t* total) {
    return scan_impl<uint32_t>(ctx, in, out, n, total);
}
