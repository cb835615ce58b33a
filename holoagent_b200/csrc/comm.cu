// comm.cu - multi-GPU ingest through the C-ABI: NCCL over NVLink 5 / NVSwitch, one ctx per GPU.
//
// The reference has no multi-GPU path; north_star shards the frame batches across the GPUs of one box.  A C / C++
// caller of libhmsg_b200.so gets the whole sharded build without torch:
//   hmsg_comm_unique_id / hmsg_comm_init   (or hmsg_comm_attach with the caller's ncclComm_t)
//   hmsg_voxel_build_sharded    local bounds / occupancy / accumulation over this rank's frame ranges merged by
//                               an all-reduce(min) of 6 doubles, an all-gather + OR of the occupancy bitmap and an
//                               all-reduce(sum) of the f64 voxel accumulators  -> identical voxel table on every rank
//   hmsg_radius_filter_sharded  each rank counts neighbours for its slice of the voxel table, one exchange of the
//                               uint32 counts, identical node table on every rank
//   hmsg_allgather_nodes        the node-embedding merge: every rank holds a dense partial [n_nodes, d] (+ counter).
//                               Row slices are exchanged all-to-all (grouped ncclSend/ncclRecv: NVSwitch gives every
//                               pair full bandwidth), the owner sums the `world` partials of its slice IN RANK ORDER
//                               (deterministic, independent of NCCL's algorithm choice) and the finished slices are
//                               all-gathered.  Per-rank traffic 2 * (world-1)/world * n*(d+1)*4 B instead of the
//                               (world-1) * n*(d+1)*4 B of gathering dense partials.  Optionally the per-frame mask
//                               embeddings F_p (the reference's frames_feats) ride in the same group.
// NCCL is resolved at run time with dlopen (the copy torch already loaded when there is one, so a caller-supplied
// communicator and the library agree on the NCCL build); the handful of prototypes used here are declared locally.
#include "common.cuh"
#include <dlfcn.h>
#include <algorithm>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
static NcclApi g_nccl;

static bool nccl_load() {
  if (g_nccl.lib) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process (torch's bundled NCCL)
  if (!h) { const char* e = getenv("HMSG_NCCL_LIB"); if (e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_nccl.error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
  bool ok = true;
  auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) { ok = false; g_nccl.error = std::string("dlsym ") + n; } return p; };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.CommCount = (decltype(g_nccl.CommCount))sym("ncclCommCount");
  g_nccl.CommUserRank = (decltype(g_nccl.CommUserRank))sym("ncclCommUserRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
  g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!ok) return false;
  g_nccl.lib = h;
  return true;
}

#define HMSG_NCCL(call)                                                                                       \
  do {                                                                                                        \
    ncclResult_t r__ = (call);                                                                                \
    if (r__ != 0) return ctx->fail(HMSG_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__));    \
  } while (0)

struct CommState {
  ncclComm_t comm = nullptr;
  bool owned = false;
  int rank = 0, world = 1;
  double* d_mm = nullptr;                                   // 6 doubles: (min xyz, -max xyz)
  uint32_t* gather = nullptr; size_t gather_bytes = 0;      // bitmap all-gather
  float* recv = nullptr; size_t recv_bytes = 0;             // node partial slices [world][rows_per*(d+1)]
  // last exchange (bench "scaling report"): bytes this rank sent + received
  double last_bytes = 0.0;
};

int32_t comm_destroy(hmsg_ctx* ctx) {
  CommState* cs = ctx->comm;
  if (!cs) return HMSG_OK;
  if (cs->owned && cs->comm && g_nccl.lib) g_nccl.CommDestroy(cs->comm);
  free_dev(cs->d_mm); free_dev(cs->gather); free_dev(cs->recv);
  delete cs;
  ctx->comm = nullptr;
  return HMSG_OK;
}

extern "C" int32_t hmsg_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return HMSG_ERR_ARG;
  if (!nccl_load()) return HMSG_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return HMSG_ERR_NCCL;
  memcpy(id_out, id.internal, 128);
  return HMSG_OK;
}

extern "C" int32_t hmsg_comm_init(hmsg_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!id || world < 1 || rank < 0 || rank >= world) return ctx->fail(HMSG_ERR_ARG, "hmsg_comm_init: bad argument");
  if (!nccl_load()) return ctx->fail(HMSG_ERR_NCCL, "hmsg_comm_init: " + g_nccl.error);
  comm_destroy(ctx);
  HMSG_CUDA(cudaSetDevice(ctx->device));
  CommState* cs = new CommState();
  ctx->comm = cs;
  ncclUniqueId uid;
  memcpy(uid.internal, id, 128);
  HMSG_NCCL(g_nccl.CommInitRank(&cs->comm, world, uid, rank));
  cs->owned = true; cs->rank = rank; cs->world = world;
  return HMSG_OK;
}

extern "C" int32_t hmsg_comm_attach(hmsg_ctx* ctx, void* nccl_comm) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!nccl_comm) return ctx->fail(HMSG_ERR_ARG, "hmsg_comm_attach: null communicator");
  if (!nccl_load()) return ctx->fail(HMSG_ERR_NCCL, "hmsg_comm_attach: " + g_nccl.error);
  comm_destroy(ctx);
  CommState* cs = new CommState();
  ctx->comm = cs;
  cs->comm = (ncclComm_t)nccl_comm;
  cs->owned = false;
  HMSG_NCCL(g_nccl.CommCount(cs->comm, &cs->world));
  HMSG_NCCL(g_nccl.CommUserRank(cs->comm, &cs->rank));
  return HMSG_OK;
}

extern "C" int32_t hmsg_comm_info(hmsg_ctx* ctx, int32_t* rank, int32_t* world, double* last_exchange_bytes) {
  if (!ctx) return HMSG_ERR_ARG;
  CommState* cs = ctx->comm;
  if (rank) *rank = cs ? cs->rank : 0;
  if (world) *world = cs ? cs->world : 1;
  if (last_exchange_bytes) *last_exchange_bytes = cs ? cs->last_bytes : 0.0;
  return HMSG_OK;
}

// resolves the communicator of a call: the caller's (nccl_comm != NULL; attached on first use) or the ctx's own
static int32_t comm_of(hmsg_ctx* ctx, void* nccl_comm, CommState** out) {
  if (nccl_comm && (!ctx->comm || ctx->comm->comm != (ncclComm_t)nccl_comm)) {
    int32_t rc = hmsg_comm_attach(ctx, nccl_comm);
    if (rc) return rc;
  }
  if (!ctx->comm) return ctx->fail(HMSG_ERR_STATE, "no communicator: call hmsg_comm_init / hmsg_comm_attach or pass an ncclComm_t");
  *out = ctx->comm;
  return HMSG_OK;
}

__global__ void k_pack_minmax(const long long* __restrict__ ord, double* __restrict__ mm) {
  int k = threadIdx.x;
  if (k >= 6) return;
  long long o = ord[k];
  long long b = o >= 0 ? o : (o ^ 0x7FFFFFFFFFFFFFFFLL);
  double v = __longlong_as_double(b);
  mm[k] = k < 3 ? v : -v;                                   // one MIN all-reduce covers both bounds
}

extern "C" int32_t hmsg_voxel_build_sharded(hmsg_ctx* ctx, void* nccl_comm, const int64_t* ranges, int32_t n_ranges, int64_t* n_voxels,
                                            double* min_bound_out) {
  if (!ctx) return HMSG_ERR_ARG;
  CommState* cs;
  int32_t rc;
  if ((rc = comm_of(ctx, nccl_comm, &cs))) return rc;
  if (n_ranges < 0 || (n_ranges > 0 && !ranges)) return ctx->fail(HMSG_ERR_ARG, "hmsg_voxel_build_sharded: bad ranges");
  // ---- bounds: local min/max over this rank's frames, then all-reduce
  double mm[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n_ranges; i++) {
    double t[6];
    if ((rc = hmsg_voxel_bounds(ctx, ranges[2 * i], ranges[2 * i + 1], t))) return rc;
    for (int k = 0; k < 3; k++) { mm[k] = std::min(mm[k], t[k]); mm[3 + k] = std::max(mm[3 + k], t[3 + k]); }
  }
  if (!cs->d_mm) HMSG_CUDA(cudaMalloc((void**)&cs->d_mm, 64));
  double pk[6] = {mm[0], mm[1], mm[2], -mm[3], -mm[4], -mm[5]};
  ctx->prof_begin(PROF_COMM);
  HMSG_CUDA(cudaMemcpyAsync(cs->d_mm, pk, 48, cudaMemcpyHostToDevice, ctx->stream));
  HMSG_NCCL(g_nccl.AllReduce(cs->d_mm, cs->d_mm, 6, NCCL_FLOAT64, NCCL_MIN, cs->comm, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(pk, cs->d_mm, 48, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->prof_end(PROF_COMM, 48.0);
  for (int k = 0; k < 3; k++) { mm[k] = pk[k]; mm[3 + k] = -pk[3 + k]; }
  if ((rc = hmsg_voxel_grid_set(ctx, mm))) return rc;
  // ---- occupancy: local marks, all-gather, OR
  for (int i = 0; i < n_ranges; i++)
    if ((rc = hmsg_voxel_mark(ctx, ranges[2 * i], ranges[2 * i + 1]))) return rc;
  const long long nw = ctx->grid.nwords;
  if ((rc = ctx->reserve(&cs->gather, &cs->gather_bytes, (size_t)nw * 4 * cs->world))) return rc;
  ctx->prof_begin(PROF_COMM);
  HMSG_NCCL(g_nccl.AllGather(ctx->bitmap, cs->gather, (size_t)nw * 4, NCCL_UINT8, cs->comm, ctx->stream));
  ctx->prof_end(PROF_COMM, (double)nw * 4 * (cs->world - 1) * 2);
  if ((rc = hmsg_voxel_bitmap_or(ctx, cs->gather, cs->world))) return rc;
  int64_t nv = 0;
  if ((rc = hmsg_voxel_scan(ctx, &nv))) return rc;
  // ---- accumulators: local sums, all-reduce
  for (int i = 0; i < n_ranges; i++)
    if ((rc = hmsg_voxel_accumulate(ctx, ranges[2 * i], ranges[2 * i + 1]))) return rc;
  ctx->prof_begin(PROF_COMM);
  if (nv > 0) {
    HMSG_NCCL(g_nccl.AllReduce(ctx->vox_acc, ctx->vox_acc, (size_t)nv * 6, NCCL_FLOAT64, NCCL_SUM, cs->comm, ctx->stream));
    HMSG_NCCL(g_nccl.AllReduce(ctx->vox_cnt, ctx->vox_cnt, (size_t)nv, NCCL_UINT32, NCCL_SUM, cs->comm, ctx->stream));
  }
  ctx->prof_end(PROF_COMM, (double)nv * 52 * 2.0 * (cs->world - 1) / cs->world);
  if ((rc = hmsg_voxel_finalize(ctx))) return rc;
  if (n_voxels) *n_voxels = nv;
  if (min_bound_out) for (int k = 0; k < 3; k++) min_bound_out[k] = ctx->min_bound[k];
  return HMSG_OK;
}

static inline void slice_of(long long n, int world, int r, long long* begin, long long* cnt) {
  const long long per = (n + world - 1) / world;
  const long long b = std::min(n, per * r), e = std::min(n, per * (r + 1));
  *begin = b; *cnt = e - b;
}

extern "C" int32_t hmsg_radius_filter_sharded(hmsg_ctx* ctx, void* nccl_comm, int32_t nb_points, double radius, int64_t* n_nodes) {
  if (!ctx) return HMSG_ERR_ARG;
  CommState* cs;
  int32_t rc;
  if ((rc = comm_of(ctx, nccl_comm, &cs))) return rc;
  const long long nv = ctx->n_voxels;
  long long b, c;
  slice_of(nv, cs->world, cs->rank, &b, &c);
  if ((rc = geometry_radius_count(ctx, radius, b, c))) return rc;
  ctx->prof_begin(PROF_COMM);
  HMSG_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < cs->world; p++) {
    long long pb, pc;
    slice_of(nv, cs->world, p, &pb, &pc);
    if (p == cs->rank) continue;
    if (c > 0) HMSG_NCCL(g_nccl.Send(ctx->rad_cnt + b, (size_t)c, NCCL_UINT32, p, cs->comm, ctx->stream));
    if (pc > 0) HMSG_NCCL(g_nccl.Recv(ctx->rad_cnt + pb, (size_t)pc, NCCL_UINT32, p, cs->comm, ctx->stream));
  }
  HMSG_NCCL(g_nccl.GroupEnd());
  ctx->prof_end(PROF_COMM, (double)nv * 4 * 2.0 * (cs->world - 1) / cs->world);
  return geometry_radius_finish(ctx, nb_points, n_nodes);
}

// out[i] = sum over ranks r = 0..world-1 (in that order) of recv[r][i]
__global__ void __launch_bounds__(256) k_sum_slices(const float* __restrict__ recv, int world, long long stride, long long count, float* __restrict__ out_sum,
                                                    long long nd, float* __restrict__ out_cnt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float a = 0.f;
  for (int r = 0; r < world; r++) a += recv[(long long)r * stride + i];
  if (i < nd) out_sum[i] = a; else out_cnt[i - nd] = a;
}

extern "C" int32_t hmsg_allgather_nodes(hmsg_ctx* ctx, void* nccl_comm, const float* Fp_local, int64_t fp_floats, float* Fp_all,
                                        int64_t fp_stride_floats) {
  if (!ctx) return HMSG_ERR_ARG;
  CommState* cs;
  int32_t rc;
  if ((rc = comm_of(ctx, nccl_comm, &cs))) return rc;
  if (!ctx->sum_feats || ctx->d == 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_allgather_nodes: call hmsg_features_begin first");
  if (Fp_all && (!Fp_local || fp_floats < 0 || fp_stride_floats < fp_floats))
    return ctx->fail(HMSG_ERR_ARG, "hmsg_allgather_nodes: bad F_p arguments (Fp_local must hold fp_stride_floats floats)");
  const int W = cs->world, me = cs->rank, d = ctx->d;
  const long long n = ctx->n_nodes;
  const long long per = (n + W - 1) / W;
  long long mb, mc;
  slice_of(n, W, me, &mb, &mc);
  const long long stride = per * (d + 1);
  if ((rc = ctx->reserve(&cs->recv, &cs->recv_bytes, (size_t)std::max<long long>(stride, 1) * W * 4))) return rc;
  double bytes = 0.0;
  ctx->prof_begin(PROF_COMM);
  // ---- phase 1: all-to-all of the partial row slices (+ F_p rows when asked)
  HMSG_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < W; p++) {
    long long pb, pc;
    slice_of(n, W, p, &pb, &pc);
    if (pc > 0) {
      HMSG_NCCL(g_nccl.Send(ctx->sum_feats + pb * d, (size_t)pc * d, NCCL_FLOAT32, p, cs->comm, ctx->stream));
      HMSG_NCCL(g_nccl.Send(ctx->counter + pb, (size_t)pc, NCCL_FLOAT32, p, cs->comm, ctx->stream));
    }
    if (mc > 0) {
      HMSG_NCCL(g_nccl.Recv(cs->recv + (long long)p * stride, (size_t)mc * d, NCCL_FLOAT32, p, cs->comm, ctx->stream));
      HMSG_NCCL(g_nccl.Recv(cs->recv + (long long)p * stride + mc * d, (size_t)mc, NCCL_FLOAT32, p, cs->comm, ctx->stream));
    }
    if (p != me) bytes += (double)(pc + mc) * (d + 1) * 4;
    if (Fp_all && fp_stride_floats > 0) {
      HMSG_NCCL(g_nccl.Send(Fp_local, (size_t)fp_stride_floats, NCCL_FLOAT32, p, cs->comm, ctx->stream));
      HMSG_NCCL(g_nccl.Recv(Fp_all + (long long)p * fp_stride_floats, (size_t)fp_stride_floats, NCCL_FLOAT32, p, cs->comm, ctx->stream));
      if (p != me) bytes += 2.0 * fp_stride_floats * 4;
    }
  }
  HMSG_NCCL(g_nccl.GroupEnd());
  // ---- owner sums its slice in rank order
  if (mc > 0) {
    const long long count = mc * (d + 1);
    k_sum_slices<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(cs->recv, W, stride, count, ctx->sum_feats + mb * d, mc * d, ctx->counter + mb);
    HMSG_LAUNCH_CHECK();
  }
  // ---- phase 2: all-gather of the finished slices (in place: disjoint row ranges)
  HMSG_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < W; p++) {
    if (p == me) continue;
    long long pb, pc;
    slice_of(n, W, p, &pb, &pc);
    if (mc > 0) {
      HMSG_NCCL(g_nccl.Send(ctx->sum_feats + mb * d, (size_t)mc * d, NCCL_FLOAT32, p, cs->comm, ctx->stream));
      HMSG_NCCL(g_nccl.Send(ctx->counter + mb, (size_t)mc, NCCL_FLOAT32, p, cs->comm, ctx->stream));
    }
    if (pc > 0) {
      HMSG_NCCL(g_nccl.Recv(ctx->sum_feats + pb * d, (size_t)pc * d, NCCL_FLOAT32, p, cs->comm, ctx->stream));
      HMSG_NCCL(g_nccl.Recv(ctx->counter + pb, (size_t)pc, NCCL_FLOAT32, p, cs->comm, ctx->stream));
    }
    bytes += (double)(pc + mc) * (d + 1) * 4;
  }
  HMSG_NCCL(g_nccl.GroupEnd());
  ctx->prof_end(PROF_COMM, bytes);
  cs->last_bytes = bytes;
  return HMSG_OK;
}
