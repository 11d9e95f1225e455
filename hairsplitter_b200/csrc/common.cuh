// Shared host/device plumbing for libhsgpu (sm_100a only; no CPU fallback anywhere in this library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/hsgpu.h"

#define HS_TILE 128        // columns per pileup tile (one CTA of the column kernels)
#define HS_ALIGN 16        // every pileup row is padded to 16-byte boundaries in column space
#define HS_NCODES 125      // 3-mer codes '!'..'!'+124
#define HS_CODE0 33

struct HsProfEntry {
    const char* name;
    cudaEvent_t begin, end;
};

struct hsgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 0;
    // optional per-kernel timing (hsgpu_profile_enable): one event pair per launch
    bool profiling = false;
    std::vector<HsProfEntry> prof;
    std::string prof_report;
    void* d_rank_lut = nullptr;  // HsRankLut (rank.cuh), built once per context
    // one pinned word + event for a value read back without a stream stall (hs_resolve_max_tile_reads)
    int64_t* h_scratch = nullptr;
    cudaEvent_t scratch_event = nullptr;
    struct hsgpu_pileup* scratch_owner = nullptr;
    // growable pinned staging area for results that are read back in one piece (hs_host_stage)
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    // every context allocates from its own stream-ordered pool: what a context frees is what its next batch takes,
    // whatever the other contexts of the device are doing (with the device's shared default pool, contexts on
    // different streams kept making the pool grow -- physical allocations under the driver's global lock, 10-20 ms
    // stalls of every thread)
    cudaMemPool_t pool = nullptr;
    // how the host waits for the stream: spinning (the driver's default, lowest latency) or sleeping on an event
    // created with cudaEventBlockingSync (HSGPU_WAIT=block: several contexts per core, e.g. 8 ranks x 3 host threads
    // on a 16-core box, where spinning waiters take the cores from the threads that have work)
    cudaEvent_t wait_event = nullptr;
};

static inline cudaError_t hs_stream_sync(hsgpu_ctx* ctx) {
    if (!ctx->wait_event) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->wait_event, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->wait_event);
}

static inline cudaError_t hs_malloc_async(hsgpu_ctx* ctx, void** p, size_t bytes) {
    if (ctx->pool) return cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream);
    return cudaMallocAsync(p, bytes, ctx->stream);
}

// pinned staging area of at least `bytes` (contents are not kept when it grows); null on failure
void* hs_host_stage(hsgpu_ctx* ctx, size_t bytes);

void hs_prof_begin(hsgpu_ctx* ctx, const char* name);
void hs_prof_end(hsgpu_ctx* ctx);

void hs_set_error(hsgpu_ctx* ctx, const std::string& msg);
int hs_cuda_fail(hsgpu_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define HS_CUDA(ctx, call)                                                        \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return hs_cuda_fail((ctx), _e, #call, __FILE__, __LINE__); \
    } while (0)

#define HS_LAUNCH_CHECK(ctx)                                                      \
    do {                                                                          \
        (ctx)->launches++;                                                        \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess) return hs_cuda_fail((ctx), _e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// launch a kernel: counts it, checks the launch, and brackets it with events when profiling is on
#define HS_KERNEL(ctx, name, ...)                  \
    do {                                           \
        if ((ctx)->profiling) hs_prof_begin((ctx), (name)); \
        __VA_ARGS__;                               \
        if ((ctx)->profiling) hs_prof_end((ctx));  \
        HS_LAUNCH_CHECK(ctx);                      \
    } while (0)

#define HS_FAIL(ctx, code, msg)      \
    do {                             \
        hs_set_error((ctx), (msg));  \
        return (code);               \
    } while (0)

// stream-ordered device allocation helpers
template <typename T>
static inline cudaError_t hs_alloc(hsgpu_ctx* ctx, T** p, int64_t n) {
    if (n <= 0) n = 1;
    return hs_malloc_async(ctx, (void**)p, (size_t)n * sizeof(T));
}
template <typename T>
static inline void hs_free(hsgpu_ctx* ctx, T*& p) {
    if (p) cudaFreeAsync((void*)p, ctx->stream);
    p = nullptr;
}
// The stream-ordered temporaries of one call. A call frees them itself on its way out (hs_free nulls the pointer);
// whatever is still allocated when the call returns early -- HS_CUDA / HS_KERNEL / HS_FAIL on an error -- goes back to
// the context's pool here instead of staying allocated until the context is destroyed.
struct HsTemps {
    hsgpu_ctx* ctx;
    std::vector<void**> slots;
    explicit HsTemps(hsgpu_ctx* c) : ctx(c) {}
    template <typename T>
    void own(T*& p) { slots.push_back(reinterpret_cast<void**>(&p)); }
    template <typename T, typename... Rest>
    void own(T*& p, Rest&... rest) {
        own(p);
        own(rest...);
    }
    ~HsTemps() {
        for (void** s : slots)
            if (*s) {
                cudaFreeAsync(*s, ctx->stream);
                *s = nullptr;
            }
    }
};
// several arrays out of ONE stream-ordered allocation (a batch object used to cost ~40 cudaMallocAsync /
// cudaFreeAsync pairs; the host side of the e2e path is bound by driver calls)
struct HsCarve {
    struct Piece { void** slot; size_t bytes; };
    std::vector<Piece> pieces;
    template <typename T>
    void add(T** slot, int64_t n) {
        if (n <= 0) n = 1;
        pieces.push_back({reinterpret_cast<void**>(slot), ((size_t)n * sizeof(T) + 255) & ~(size_t)255});
    }
    size_t total() const {
        size_t t = 0;
        for (const Piece& q : pieces) t += q.bytes;
        return t ? t : 256;
    }
    void place(void* base) {  // the pieces inside a block the caller owns (256-byte aligned)
        size_t off = 0;
        for (const Piece& q : pieces) {
            *q.slot = static_cast<char*>(base) + off;
            off += q.bytes;
        }
    }
    cudaError_t alloc(hsgpu_ctx* ctx, void** base) {
        size_t total = 0;
        for (const Piece& q : pieces) total += q.bytes;
        cudaError_t e = hs_malloc_async(ctx, base, total ? total : 256);
        if (e != cudaSuccess) return e;
        size_t off = 0;
        for (const Piece& q : pieces) {
            *q.slot = static_cast<char*>(*base) + off;
            off += q.bytes;
        }
        return cudaSuccess;
    }
};

template <typename T>
static inline cudaError_t hs_h2d(hsgpu_ctx* ctx, T* dst, const T* src, int64_t n) {
    if (n <= 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
}
template <typename T>
static inline cudaError_t hs_d2h(hsgpu_ctx* ctx, T* dst, const T* src, int64_t n) {
    if (n <= 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
}

// exclusive prefix sums on the device (scan.cu). out may alias in. total (device pointer, may be
// null) receives the grand total.
int hs_exclusive_scan_i64(hsgpu_ctx* ctx, const int64_t* in, int64_t* out, int64_t n, int64_t* total);
int hs_exclusive_scan_u32_to_i64(hsgpu_ctx* ctx, const uint32_t* in, int64_t* out, int64_t n, int64_t* total);

// ---- device helpers ----------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ int hs_base2(const uint32_t* __restrict__ words, int64_t i) {
    return (int)((__ldg(words + (i >> 4)) >> (2 * (int)(i & 15))) & 3u);
}
__device__ __forceinline__ int hs_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ long long hs_warp_incl_scan64(long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ long long hs_warp_sum64(long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
#endif

// the device-resident pileup of a batch of contigs
struct hsgpu_pileup {
    hsgpu_ctx* ctx = nullptr;
    int32_t n_contigs = 0;
    int64_t n_reads = 0;
    int64_t n_cols = 0;    // sum of contig lengths
    int64_t n_tiles = 0;
    int64_t n_cigar = 0;
    bool built = false, ranked = false;
    float auto_threshold = 0.33f;

    // host mirrors of the small per-contig arrays
    std::vector<int32_t> h_contig_len;
    std::vector<int64_t> h_contig_read_off, h_col_base, h_tile_base;
    std::vector<int64_t> h_stats;  // 3 per contig: distance, aligned, cells

    // carved allocations (HsCarve): everything hsgpu_pileup_create / the first hsgpu_column_rank allocate
    void* d_create_block = nullptr;
    void* d_rank_block = nullptr;
    std::vector<int64_t> h_tables;  // col_base | tile_base | suspect_base | super_base, uploaded in one copy

    // inputs on the device
    int32_t* d_contig_len = nullptr;
    uint32_t* d_contig_bases = nullptr;
    int64_t* d_contig_word_off = nullptr;
    int64_t* d_contig_read_off = nullptr;
    int64_t* d_col_base = nullptr;
    int64_t* d_tile_base = nullptr;
    int32_t* d_tile_contig = nullptr;
    int32_t* d_read_contig = nullptr;
    uint32_t* d_read_bases = nullptr;
    int64_t* d_read_word_off = nullptr;
    int32_t* d_read_len = nullptr;
    uint8_t* d_cigar = nullptr;   // one byte per op: len << 2 | kind (pileup.cu)
    int64_t* d_cigar_off = nullptr;
    int32_t* d_read_start = nullptr;
    uint8_t* d_read_strand = nullptr;

    // pileup proper
    int32_t* d_read_end = nullptr;    // positionOfReads[n].second
    int32_t* d_read_tlead = nullptr;  // read offset of the first aligned base (leading clips)
    uint8_t* d_read_flags = nullptr;  // HS_READ_IRREGULAR: clips between aligned parts
    unsigned int* d_next_read = nullptr;  // work counter of pileup_kernel's persistent warps
    int64_t n_irregular = 0;
    int64_t max_tile_reads = 0;       // most reads over one 128-column tile (picks the histogram width)
    int64_t* d_row_alloc = nullptr;   // bytes reserved per read (multiple of 16), then its exclusive scan
    int64_t* d_row_base = nullptr;    // codes[row_base[r] + q] = cell of read r at column q (row_base % 16 == 0)
    uint8_t* d_codes = nullptr;
    int64_t codes_bytes = 0;
    unsigned long long* d_stats = nullptr;  // 3 per contig

    // tile index: reads overlapping each 128-column tile, ascending
    int64_t* d_tile_off = nullptr;
    int32_t* d_tile_reads = nullptr;
    int64_t tile_entries = 0;
    int64_t n_super = 0;              // super-tiles of 4 tiles (first level of the index build)
    int64_t* d_super_base = nullptr;  // first super-tile of every contig
    int64_t* d_super_off = nullptr;

    // column summaries (call_variants)
    uint8_t* d_k0 = nullptr;
    uint8_t* d_k1 = nullptr;
    uint8_t* d_flags = nullptr;
    uint32_t* d_counts = nullptr;  // c0,c1,c2 per column
    uint32_t* d_depth = nullptr;   // cells per column
    int32_t* d_min_reads = nullptr;
    int32_t* d_suspect_pos = nullptr;   // per contig region [col_base[c]/6 + 8c ...)
    uint8_t* d_suspect_auto = nullptr;
    int32_t* d_n_suspects = nullptr;
    unsigned long long* d_depth_sum = nullptr;
    std::vector<int64_t> h_suspect_base;
    int64_t* d_suspect_base = nullptr;
    int64_t* d_col_off = nullptr;  // exclusive scan of d_depth (lazy, for export)
    int32_t* d_work = nullptr;     // [0..3] counters, then column ids for the re-read literal replay
    uint32_t* d_arena = nullptr;   // (code, count) lists of the columns whose ties are resolved by the deferred kernel
    uint32_t* d_item_off = nullptr;
    unsigned int arena_words = 0;
    int64_t* d_tile_sus = nullptr; // accepted suspects per tile, then its exclusive scan
    bool have_col_off = false;

    // partitions of the batch (hsgpu_partitions_set) and the work arrays of the robust filter (contingency.cu)
    void* d_filter_block = nullptr;
    void* d_fdesc = nullptr;
    uint32_t* d_frows = nullptr;  // 2-bit partition states per read (contingency.cu)
    bool have_parts = false;
    void* d_filter_work = nullptr;
    uint32_t* d_factive = nullptr;
    uint32_t* d_fkept = nullptr;  // bitmap of the kept columns
    int32_t* d_fkept_list = nullptr;
    unsigned int* d_fcounters = nullptr;
    uint32_t* d_foverflow = nullptr;  // columns handed from robust_filter_lanes_kernel to robust_filter_kernel
    int64_t* d_fhdr = nullptr;
};

int hs_resolve_max_tile_reads(hsgpu_pileup* p);

// column flags
#define HS_FLAG_CANDIDATE 1  // passes :525-528 (everything but the spacing rule)
#define HS_FLAG_AUTO 2       // c1 > u*c0 (:531)
#define HS_FLAG_RESCUE 4     // rescue pre-filter of loop 4 (:751-752)
#define HS_FLAG_SUSPECT 8    // candidate that also passed the spacing rule (:529)
#define HS_FLAG_INLIST 16    // member of the caller's snps_in of hsgpu_robust_filter (set and cleared by that call)
#define HS_FLAG_ACTIVE 32    // rescue candidate whose second code has more than 4 carriers: loop 4 can keep it (:745-764)
