// common.cuh - context, error handling and exact-rounding device helpers shared by the
// sm_100a kernels of libhmsg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include "../../include/hmsg_b200.h"

#define HMSG_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      return ctx->fail(HMSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                     \
  } while (0)

#define HMSG_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    ctx->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return ctx->fail(HMSG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// ---- optional per-kernel-class CUDA-event profiling (bench.py roofline numbers)
enum { PROF_GEMM = 0, PROF_ATTN, PROF_ELTWISE, PROF_KNN, PROF_NN, PROF_SCATTER, PROF_GEOM, PROF_CROPS, PROF_MASK3D, PROF_COMM, PROF_NCLASS };
struct ProfClass {
  std::vector<cudaEvent_t> ev;   // pairs (begin, end)
  size_t used = 0;
  double work = 0.0;             // algorithmic flops or bytes accumulated
};

struct VitState;   // encoder.cu
struct KnnState;   // knn.cu
struct ObjState;   // objects.cu
struct M3dState;   // masks3d.cu
struct CommState;  // comm.cu

struct GridDesc {
  double vmin[3];     // min_bound - vs/2  (Open3D voxel_min_bound)
  double vs;
  double inv_vs;      // rn(1/vs): used only for the fast floor below and for pruning bounds
  int nx, ny, nz, nzp;   // nzp = nz rounded up to a multiple of 32: a bitmap word never spans columns
  long long nwords;
  // coarse level for the far search: one bit per 4x4x4 block of cells (node occupancy)
  int cnx, cny, cnz, cnzp;
  long long cnwords;
};

struct CamDesc {
  double fx, fy, cx, cy;
  float scale;
  int H, W;
};

// host -> device frame uploads run on their own stream; consumers wait for exactly the uploads they read
struct UploadRec {
  int64_t f0, n;
  cudaEvent_t ev;
  bool waited;
};

struct hmsg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // hmsg_scene_put_frames from host memory (H2D overlaps compute on `stream`)
  std::vector<UploadRec> uploads;
  std::vector<cudaEvent_t> upload_events;  // pool (reused across scenes)
  size_t upload_events_used = 0;
  cudaEvent_t sync_event = nullptr;
  // make `stream` wait for every pending upload that overlaps frames [f0, f0 + n)
  void wait_frames(int64_t f0, int64_t n) {
    for (auto& u : uploads)
      if (!u.waited && u.f0 < f0 + n && f0 < u.f0 + u.n) { cudaStreamWaitEvent(stream, u.ev, 0); u.waited = true; }
  }
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;

  // ---- scene
  CamDesc cam{};
  double vs = 0.05;
  int64_t cap = 0, nframes = 0;
  uint16_t* depth = nullptr;   // [cap,H,W]
  uint8_t* rgb = nullptr;      // [cap,H,W,3]
  double* poses = nullptr;     // [cap,16]
  double* frameK = nullptr;    // [cap,4] per-frame (fx, fy, cx, cy) or nullptr (one K per scene)

  // ---- voxel table (A2)
  long long* d_bounds = nullptr;   // 6 ordered-int doubles: min xyz, max xyz
  double min_bound[3]{}, max_bound[3]{};
  GridDesc grid{};
  uint32_t* bitmap = nullptr;      // occupancy, 1 bit per cell, k fastest
  uint32_t* prefix = nullptr;      // exclusive popcount prefix per word
  uint32_t* blocksums = nullptr;
  size_t bitmap_cap_words = 0;     // grow-only capacities (a re-build of the same scene does not touch the allocator)
  size_t voxel_cap = 0, node_cap = 0, cbitmap_cap_words = 0;
  int64_t n_voxels = 0;
  double* vox_acc = nullptr;       // [n_voxels,6] sum -> mean of xyz, rgb
  uint32_t* vox_cnt = nullptr;
  int32_t* vox_ijk = nullptr;
  uint32_t* rad_cnt = nullptr;
  bool voxels_built = false;
  bool grid_set = false;

  // ---- nodes (A3)
  int64_t n_nodes = 0;
  uint32_t* nbitmap = nullptr;     // occupancy of kept voxels
  uint32_t* nprefix = nullptr;
  uint32_t* cbitmap = nullptr;     // coarse (4x4x4 blocks) occupancy of nodes
  int* far_list = nullptr;         // worklist of (frame-in-batch * HW + pixel) needing the far search
  size_t far_list_bytes = 0;
  int* far_count = nullptr;
  double* node_xyz = nullptr;      // [n_nodes,3]
  double* node_rgb = nullptr;
  int32_t* node_ijk = nullptr;
  int64_t* node_vox = nullptr;
  bool nodes_built = false;

  // ---- features (A5/A6)
  int d = 0;
  float* sum_feats = nullptr;      // [n_nodes,d]
  float* counter = nullptr;        // [n_nodes]
  size_t feat_cap = 0;
  // batch scratch
  int batch_cap = 0;
  int batch_M = 0, batch_MW = 0;
  int64_t batch_begin = -1;
  int batch_n = 0;
  std::vector<int> batch_counts;   // real masks per frame of the batch (ragged SAM output; <= batch_M)
  int32_t* mask_cnt = nullptr;     // device copy [batch_cap]
  size_t mask_cnt_bytes = 0;
  int32_t* mask_rect = nullptr;    // [batch_cap, M, 4] pixel bounding box (x0, y0, x1, y1) exclusive of every mask (A7 scans it)
  size_t mask_rect_bytes = 0;
  uint32_t* maskbits = nullptr;    // [batch_cap, H*W, MW]
  size_t maskbits_bytes = 0;
  int32_t* pix_idx = nullptr;      // [batch_cap, H*W]
  int64_t pix_idx_for = -1;        // batch_begin the pixel->node map was last computed for (-1: stale)
  size_t pix_idx_bytes = 0;
  unsigned long long* win = nullptr;   // [batch_cap, n_nodes]
  size_t win_bytes = 0;
  uint32_t epoch = 0;
  float* Fp = nullptr;             // [batch_cap, M, d]
  size_t Fp_bytes = 0;
  float* feats_stage = nullptr;    // staging for host-provided encoder outputs
  size_t feats_stage_bytes = 0;
  int32_t* boxes_stage = nullptr;
  size_t boxes_stage_bytes = 0;
  uint8_t* seg_stage = nullptr;
  size_t seg_stage_bytes = 0;
  // generic scratch
  void* scratch = nullptr;
  size_t scratch_bytes = 0;

  VitState* vit = nullptr;
  KnnState* knn = nullptr;
  ObjState* obj = nullptr;
  M3dState* m3d = nullptr;
  CommState* comm = nullptr;

  uint32_t prof_mask = 0;
  ProfClass prof[PROF_NCLASS];
  void prof_begin(int cls) {
    if (!(prof_mask & (1u << cls))) return;
    ProfClass& p = prof[cls];
    if (p.used + 2 > p.ev.size()) {
      for (int i = 0; i < 2; i++) { cudaEvent_t e; cudaEventCreate(&e); p.ev.push_back(e); }
    }
    cudaEventRecord(p.ev[p.used], stream);
  }
  void prof_end(int cls, double work) {
    if (!(prof_mask & (1u << cls))) return;
    ProfClass& p = prof[cls];
    cudaEventRecord(p.ev[p.used + 1], stream);
    p.used += 2;
    p.work += work;
  }

  int32_t fail(int32_t code, const std::string& msg) {
    err = msg;
    return code;
  }
  // grow-only device buffer
  template <typename T>
  int32_t reserve(T** p, size_t* cur_bytes, size_t need_bytes) {
    if (*cur_bytes >= need_bytes && *p) return HMSG_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cur_bytes = 0;
    cudaError_t e = cudaMalloc((void**)p, need_bytes);
    if (e != cudaSuccess) return fail(HMSG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cur_bytes = need_bytes;
    return HMSG_OK;
  }
};

template <typename T>
static inline void free_dev(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// ---------------------------------------------------------------------------------------
// exact-rounding helpers: every operation is an individually rounded IEEE op (no FMA
// contraction) so that voxel keys / node indices match the float64 NumPy/Open3D arithmetic
// of the reference bit for bit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long d2ord(double v) {
  long long b = __double_as_longlong(v);
  return b >= 0 ? b : (b ^ 0x7FFFFFFFFFFFFFFFLL);
}
static inline double ord2d_host(long long o) {
  long long b = o >= 0 ? o : (o ^ 0x7FFFFFFFFFFFFFFFLL);
  double v;
  memcpy(&v, &b, 8);
  return v;
}

struct Pose {
  double t[16];
};

// generic.py:111,122-124,137: depth_f32 = u16/scale (float32 division); X=(x-cx)*d/fx ...;
// Open3D transform: T*[X Y Z 1]^T then divide by w.
__device__ __forceinline__ void unproject_px(unsigned short dep, int x, int y, const CamDesc& c,
                                             const double* T, double& wx, double& wy, double& wz) {
  float df = __fdiv_rn((float)dep, c.scale);
  double dd = (double)df;
  double X = __ddiv_rn(__dmul_rn(__dsub_rn((double)x, c.cx), dd), c.fx);
  double Y = __ddiv_rn(__dmul_rn(__dsub_rn((double)y, c.cy), dd), c.fy);
  double Z = dd;
  double w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[12], X), __dmul_rn(T[13], Y)), __dmul_rn(T[14], Z)), T[15]);
  double ax = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], X), __dmul_rn(T[1], Y)), __dmul_rn(T[2], Z)), T[3]);
  double ay = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], X), __dmul_rn(T[5], Y)), __dmul_rn(T[6], Z)), T[7]);
  double az = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], X), __dmul_rn(T[9], Y)), __dmul_rn(T[10], Z)), T[11]);
  if (w == 1.0) {   // rigid pose (bottom row 0 0 0 1): x / 1.0 == x exactly, skip three fp64 divisions
    wx = ax; wy = ay; wz = az;
  } else {
    wx = __ddiv_rn(ax, w);
    wy = __ddiv_rn(ay, w);
    wz = __ddiv_rn(az, w);
  }
}

// Open3D VoxelDownSample: ref = (p - voxel_min_bound) / voxel_size; floor
__device__ __forceinline__ double cell_coord(double p, double vmin, double vs) {
  return __ddiv_rn(__dsub_rn(p, vmin), vs);
}

// floor((p - vmin) / vs) exactly as the reference computes it, without paying an fp64 division per
// coordinate: q = (p - vmin) * rn(1/vs) differs from rn((p - vmin)/vs) by < 7e-10 for |q| < 2^21, so
// the floors agree unless q is within 1e-8 of an integer - only then is the exact division evaluated.
__device__ __forceinline__ int cell_index(double p, double vmin, double vs, double inv_vs) {
  double t = __dsub_rn(p, vmin);
  double q = __dmul_rn(t, inv_vs);
  double f = floor(q);
  double frac = q - f;
  if (frac < 1e-8 || frac > 1.0 - 1e-8) f = floor(__ddiv_rn(t, vs));
  return (int)f;
}

__device__ __forceinline__ double sqdist3(double ax, double ay, double az, double bx, double by, double bz) {
  double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// sub-module entry points used by api.cu
int32_t vit_destroy(hmsg_ctx* ctx);
int32_t knn_destroy(hmsg_ctx* ctx);
int32_t crops_destroy(hmsg_ctx* ctx);
int32_t objects_destroy(hmsg_ctx* ctx);
int32_t masks3d_destroy(hmsg_ctx* ctx);
void masks3d_invalidate_scratch(hmsg_ctx* ctx);
int32_t comm_destroy(hmsg_ctx* ctx);
int32_t features_ensure_pix_idx(hmsg_ctx* ctx);
int32_t geometry_radius_count(hmsg_ctx* ctx, double radius, int64_t v_begin, int64_t n);
int32_t geometry_radius_finish(hmsg_ctx* ctx, int32_t nb_points, int64_t* n_nodes);
int32_t geometry_points_to_node_dev(hmsg_ctx* ctx, const double* d_pts, long long n, int64_t* d_idx, double* d_dist);
int32_t vit_encode_device(hmsg_ctx* ctx, const float* dx, int B, float* dout, int normalize);
