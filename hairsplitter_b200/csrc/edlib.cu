// placeholder until the Myers kernel lands (next commit): fails loudly, never falls back to a CPU
#include "common.cuh"
extern "C" int hsgpu_edlib_align_batch(hsgpu_ctx* ctx, int32_t, const char*, const int64_t*, const char*, const int64_t*,
                                       int32_t, int32_t, int32_t, hsgpu_edlib_result*, int32_t*, int32_t*, int64_t,
                                       uint8_t*, int64_t) {
    HS_FAIL(ctx, HSGPU_ERR_STATE, "hsgpu_edlib_align_batch: not implemented in this build");
}
