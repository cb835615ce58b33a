// crops.cu - A8 / N3: crop + resize + open_clip preprocess on device (placeholder until the
// bit-exact cv2/PIL resamplers land; the entry point reports that clearly).
#include "common.cuh"

extern "C" int32_t hmsg_make_crops(hmsg_ctx* ctx, int64_t, int32_t, int32_t, const int32_t*, int32_t, int32_t, float**) {
  if (!ctx) return HMSG_ERR_ARG;
  return ctx->fail(HMSG_ERR_STATE, "hmsg_make_crops: device-side crop/resize is not built yet; pass preprocessed crops to hmsg_encode_images");
}
