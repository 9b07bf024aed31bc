// lk_sw_api.cu — Smith-Waterman entry points of the C ABI (placeholder until sw_kernels.cu lands).
#include "lk_api_internal.h"

extern "C" lk_status lk_sw_batch_host(lk_ctx *ctx, int64_t, const int64_t *, const uint8_t *, const int64_t *, const uint8_t *,
                                      const lk_sw_params *, lk_sw_overhang, int32_t, uint32_t *, int32_t *, int32_t *) {
    return lk_fail(ctx, LK_ERR_INVALID, "SW not built yet");
}
extern "C" lk_status lk_sw_upload(lk_ctx *ctx, int64_t, const int64_t *, const uint8_t *, const int64_t *, const uint8_t *,
                                  const lk_sw_params *, lk_sw_overhang, int32_t, lk_resident **) {
    return lk_fail(ctx, LK_ERR_INVALID, "SW not built yet");
}
extern "C" lk_status lk_sw_run_resident(lk_ctx *ctx, lk_resident *, float *) { return lk_fail(ctx, LK_ERR_INVALID, "SW not built yet"); }
extern "C" lk_status lk_sw_download(lk_ctx *ctx, lk_resident *, uint32_t *, int32_t *, int32_t *) { return lk_fail(ctx, LK_ERR_INVALID, "SW not built yet"); }
extern "C" lk_status lk_sw_align(lk_ctx *ctx, const uint8_t *, int64_t, const uint8_t *, int64_t, const lk_sw_params *,
                                 lk_sw_overhang, char *, size_t, int32_t *) { return lk_fail(ctx, LK_ERR_INVALID, "SW not built yet"); }
