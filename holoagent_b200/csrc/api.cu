// api.cu - context management of libhmsg_b200.so (the C-ABI declared in include/hmsg_b200.h).
#include "common.cuh"

static std::string g_create_error;

extern "C" int32_t hmsg_version(void) { return 200; }

extern "C" int32_t hmsg_ctx_create(int32_t device, hmsg_ctx** out) {
  if (!out) return HMSG_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_error = std::string("hmsg_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
    return HMSG_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    g_create_error = "hmsg_ctx_create: device index out of range";
    return HMSG_ERR_ARG;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return HMSG_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "hmsg_ctx_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; this library is built for sm_100a (B200) only";
    return HMSG_ERR_CUDA;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return HMSG_ERR_CUDA; }
  hmsg_ctx* ctx = new hmsg_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return HMSG_ERR_CUDA; }
  e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->sync_event, cudaEventDisableTiming);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return HMSG_ERR_CUDA; }
  *out = ctx;
  return HMSG_OK;
}

extern "C" int32_t hmsg_ctx_destroy(hmsg_ctx* ctx) {
  if (!ctx) return HMSG_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->stream);
  vit_destroy(ctx);
  knn_destroy(ctx);
  objects_destroy(ctx);
  masks3d_destroy(ctx);
  comm_destroy(ctx);
  crops_destroy(ctx);
  free_dev(ctx->depth); free_dev(ctx->rgb); free_dev(ctx->poses); free_dev(ctx->frameK); free_dev(ctx->d_bounds);
  free_dev(ctx->bitmap); free_dev(ctx->prefix); free_dev(ctx->blocksums); free_dev(ctx->vox_acc); free_dev(ctx->vox_cnt);
  free_dev(ctx->vox_ijk); free_dev(ctx->rad_cnt); free_dev(ctx->nbitmap); free_dev(ctx->nprefix); free_dev(ctx->node_xyz);
  free_dev(ctx->node_rgb); free_dev(ctx->node_ijk); free_dev(ctx->node_vox); free_dev(ctx->sum_feats); free_dev(ctx->counter);
  free_dev(ctx->maskbits); free_dev(ctx->pix_idx); free_dev(ctx->win); free_dev(ctx->Fp); free_dev(ctx->feats_stage);
  free_dev(ctx->boxes_stage); free_dev(ctx->seg_stage); free_dev(ctx->cbitmap); free_dev(ctx->far_list); free_dev(ctx->far_count); free_dev(ctx->mask_cnt); free_dev(ctx->mask_rect);
  if (ctx->scratch) cudaFree(ctx->scratch);
  for (auto& pc : ctx->prof) for (auto e : pc.ev) cudaEventDestroy(e);
  for (auto e : ctx->upload_events) cudaEventDestroy(e);
  if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return HMSG_OK;
}

extern "C" const char* hmsg_last_error(const hmsg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int32_t hmsg_sync(hmsg_ctx* ctx) {
  if (!ctx) return HMSG_ERR_ARG;
  HMSG_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" void* hmsg_stream(hmsg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int64_t hmsg_launch_count(const hmsg_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int32_t hmsg_prof_enable(hmsg_ctx* ctx, uint32_t class_mask) {
  if (!ctx) return HMSG_ERR_ARG;
  ctx->prof_mask = class_mask;
  return HMSG_OK;
}

extern "C" int32_t hmsg_prof_read(hmsg_ctx* ctx, int32_t cls, double* ms, int64_t* launches, double* work) {
  if (!ctx) return HMSG_ERR_ARG;
  if (cls < 0 || cls >= PROF_NCLASS) return ctx->fail(HMSG_ERR_ARG, "hmsg_prof_read: bad class");
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  ProfClass& p = ctx->prof[cls];
  double tot = 0.0;
  for (size_t i = 0; i + 1 < p.used; i += 2) {
    float t = 0.f;
    cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]);
    tot += t;
  }
  if (ms) *ms = tot;
  if (launches) *launches = (int64_t)(p.used / 2);
  if (work) *work = p.work;
  p.used = 0;
  p.work = 0.0;
  return HMSG_OK;
}

int32_t vit_set_option(hmsg_ctx* ctx, const char* key, int value);
int32_t knn_set_option(hmsg_ctx* ctx, const char* key, int value);
int32_t crops_set_option(hmsg_ctx* ctx, const char* key, int value);
int32_t masks3d_set_option(hmsg_ctx* ctx, const char* key, int value);

extern "C" int32_t hmsg_set_option(hmsg_ctx* ctx, const char* key, int32_t value) {
  if (!ctx || !key) return HMSG_ERR_ARG;
  int32_t rc = vit_set_option(ctx, key, value);
  if (rc == -1) rc = knn_set_option(ctx, key, value);
  if (rc == -1) rc = crops_set_option(ctx, key, value);
  if (rc == -1) rc = masks3d_set_option(ctx, key, value);
  if (rc == -1) return ctx->fail(HMSG_ERR_ARG, std::string("hmsg_set_option: unknown key ") + key);
  return rc;
}
