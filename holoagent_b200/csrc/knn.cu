// knn.cu - A11: node-embedding x language-query similarity + top-k, one fused HBM-bound pass.
// Reference: fsr_vln/memory/hmsg/graph/graph.py:3126-3151 (query_hmsg_object core), :2196-2200
// (query_graph), :1452-1454 (identify_object), :2888-2897 (global view retrieval).
//   sim = np.dot(q, E.T) (raw dot product, nothing is normalised); argsort desc.
//
// Layout: E [N,d] float32 row-major in HBM (2.048 GB at N=1M, d=512: never fits the 126 MB
// L2, every pass streams from HBM).  One pass serves a batch of BQ query rows held in shared
// memory.  A warp owns RW consecutive rows per step: lane l loads float4 columns l+32j
// (fully coalesced 512 B per request), accumulates RW*BQ partial dot products in fp32 FMAs,
// and the warp reduces them with a transposing butterfly (RW*BQ-1 shuffles instead of
// 5*RW*BQ).  Scores never go to HBM: each warp keeps a sorted top-K list per query in
// shared memory, CTAs merge their warps' lists, and a tiny second kernel merges the per-CTA
// lists.  Ties are broken towards the lower row index.
#include "common.cuh"
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

#define KNN_TPB 256
#define KNN_WARPS (KNN_TPB / 32)
#define KMAX 32

struct KnnState {
  const float* E = nullptr;
  float* E_owned = nullptr;
  int64_t N = 0;
  int d = 0;
  float* q_dev = nullptr;      size_t q_bytes = 0;
  uint8_t* mask_dev = nullptr; size_t mask_bytes = 0;
  float* part_s = nullptr;     size_t part_s_bytes = 0;
  int* part_i = nullptr;       size_t part_i_bytes = 0;
  float* out_s = nullptr;      size_t out_s_bytes = 0;
  long long* out_i = nullptr;  size_t out_i_bytes = 0;
  int* out_n = nullptr;        size_t out_n_bytes = 0;
  float2* rowstate = nullptr;  size_t rowstate_bytes = 0;
  // k > KMAX ("return everything" calls, e.g. top_k = len(objects)): dense scores + 64-bit key sort
  float* dense = nullptr;                size_t dense_bytes = 0;
  unsigned long long* keys = nullptr;    size_t keys_bytes = 0;
  unsigned char* sort_tmp = nullptr;     size_t sort_tmp_bytes = 0;
  int grid = 0;
};

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ bool better(float s, int i, float s2, int i2) { return s > s2 || (s == s2 && i < i2); }

// transposing butterfly: v[0..NV) per lane -> lane L ends with the warp total of index L >> (5 - log2 NV)
template <int NV>
__device__ __forceinline__ float warp_transpose_reduce(float* v, int lane) {
  int off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1) {
    const int half = n >> 1;
    bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; i++) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  float r = v[0];
  for (; off >= 1; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

template <int NV> struct Log2 { static constexpr int v = 1 + Log2<NV / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// mode 0: plain top-K per query row.
// mode 1: negative-prompt selection (graph.py:3134-3151): rows whose argmax over the Qp query rows is
//         `query_id`, ranked by that max; list 0 holds the result.  q_base/rowstate allow Qp > BQ
//         (multi-pass running max/argmax per row).
template <int DV, int BQ, int RW>
__global__ void __launch_bounds__(KNN_TPB, 2)
k_sim_topk(const float* __restrict__ E, long long N, const float* __restrict__ Q, int nq, int K, const uint8_t* __restrict__ row_mask, int mode,
           int query_id, int q_base, int last_pass, float2* __restrict__ rowstate, float* __restrict__ part_s, int* __restrict__ part_i) {
  constexpr int d = 128 * DV;
  constexpr int NV = BQ * RW;
  constexpr int SH = 5 - Log2<NV>::v;
  extern __shared__ __align__(16) unsigned char smraw[];
  float4* sq = reinterpret_cast<float4*>(smraw);                                   // [BQ][d/4]
  float* tk_s = reinterpret_cast<float*>(smraw + (size_t)BQ * d * 4);             // [WARPS][BQ][KMAX]
  int* tk_i = reinterpret_cast<int*>(tk_s + KNN_WARPS * BQ * KMAX);               // [WARPS][BQ][KMAX]
  int* tk_n = tk_i + KNN_WARPS * BQ * KMAX;                                        // [WARPS][BQ]
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < BQ * (d / 4); i += blockDim.x) {
    int q = i / (d / 4);
    sq[i] = (q < nq) ? reinterpret_cast<const float4*>(Q)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < KNN_WARPS * BQ; i += blockDim.x) tk_n[i] = 0;
  __syncthreads();
  const int nlists = (mode == 0) ? BQ : 1;
  float* my_s = tk_s + wid * BQ * KMAX;
  int* my_i = tk_i + wid * BQ * KMAX;
  int* my_n = tk_n + wid * BQ;

  long long gw = (long long)blockIdx.x * KNN_WARPS + wid;
  long long tw = (long long)gridDim.x * KNN_WARPS;
  for (long long row0 = gw * RW; row0 < N; row0 += tw * RW) {
    float4 e[RW][DV];
#pragma unroll
    for (int r = 0; r < RW; r++) {
      long long row = row0 + r;
      if (row < N) {
        const float4* src = reinterpret_cast<const float4*>(E + row * d);
#pragma unroll
        for (int j = 0; j < DV; j++) e[r][j] = ld_stream(src + lane + 32 * j);
      } else {
#pragma unroll
        for (int j = 0; j < DV; j++) e[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DV; j++) {
#pragma unroll
      for (int q = 0; q < BQ; q++) {
        float4 qv = sq[q * (d / 4) + lane + 32 * j];
#pragma unroll
        for (int r = 0; r < RW; r++) {
          float a = acc[r * BQ + q];
          a = fmaf(e[r][j].x, qv.x, a); a = fmaf(e[r][j].y, qv.y, a); a = fmaf(e[r][j].z, qv.z, a); a = fmaf(e[r][j].w, qv.w, a);
          acc[r * BQ + q] = a;
        }
      }
    }
    float score = warp_transpose_reduce<NV>(acc, lane);
    int vi = lane >> SH;
    int r = vi / BQ, q = vi % BQ;
    long long row = row0 + r;
    bool leader = (lane & ((1 << SH) - 1)) == 0;
    bool rowok = row < N && (row_mask == nullptr || row_mask[row] != 0);
    bool cand;
    int list;
    if (mode == 0) {
      cand = leader && rowok && q < nq;
      list = q;
    } else {
      // argmax over the BQ query lanes of this row (first max wins, like np.argmax)
      float bs = (q < nq) ? score : -INFINITY;
      int bq = q + q_base;
#pragma unroll
      for (int o = 1; o < BQ; o <<= 1) {
        float os = __shfl_xor_sync(0xffffffffu, bs, o << SH);
        int oq = __shfl_xor_sync(0xffffffffu, bq, o << SH);
        if (os > bs || (os == bs && oq < bq)) { bs = os; bq = oq; }
      }
      bool rl = leader && q == 0 && row < N;
      if (rl && rowstate != nullptr) {
        if (q_base > 0) {
          float2 st = rowstate[row];
          if (!(bs > st.x)) { bs = st.x; bq = __float_as_int(st.y); }   // earlier query rows win ties
        }
        if (!last_pass) rowstate[row] = make_float2(bs, __int_as_float(bq));
      }
      cand = rl && rowok && last_pass && bq == query_id;
      score = bs;
      list = 0;
    }
    if (cand) {
      int n = my_n[list];
      if (n >= K) {
        float ls = my_s[list * KMAX + K - 1];
        int li = my_i[list * KMAX + K - 1];
        cand = better(score, (int)row, ls, li);
      }
    }
    unsigned need = __ballot_sync(0xffffffffu, cand);
    while (need) {
      int src = __ffs(need) - 1;
      need &= need - 1;
      float s = __shfl_sync(0xffffffffu, score, src);
      int ri = (int)__shfl_sync(0xffffffffu, (int)row, src);
      int l = __shfl_sync(0xffffffffu, list, src);
      if (lane == 0) {
        float* ls = my_s + l * KMAX;
        int* li = my_i + l * KMAX;
        int n = my_n[l];
        int pos = (n < K) ? n : K - 1;
        if (n < K || better(s, ri, ls[K - 1], li[K - 1])) {
          while (pos > 0 && better(s, ri, ls[pos - 1], li[pos - 1])) { ls[pos] = ls[pos - 1]; li[pos] = li[pos - 1]; pos--; }
          ls[pos] = s; li[pos] = ri;
          if (n < K) my_n[l] = n + 1;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // CTA merge: warp w merges list q = w, w+WARPS, ... of the 8 warps (8-way merge of sorted lists)
  for (int l = wid; l < nlists; l += KNN_WARPS) {
    int head = 0;
    int src = lane;   // lanes 0..7 own warp `lane`'s list
    int n = (src < KNN_WARPS) ? tk_n[src * BQ + l] : 0;
    for (int kk = 0; kk < K; kk++) {
      float s = -INFINITY; int i = 0x7fffffff;
      if (src < KNN_WARPS && head < n) { s = tk_s[(src * BQ + l) * KMAX + head]; i = tk_i[(src * BQ + l) * KMAX + head]; }
      float bs = s; int bi = i; int bl = lane;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        float os = __shfl_xor_sync(0xffffffffu, bs, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (better(os, oi, bs, bi)) { bs = os; bi = oi; bl = ol; }
      }
      bs = __shfl_sync(0xffffffffu, bs, 0); bi = __shfl_sync(0xffffffffu, bi, 0); bl = __shfl_sync(0xffffffffu, bl, 0);
      if (lane == bl && bi != 0x7fffffff) head++;
      if (lane == 0) {
        part_s[((long long)blockIdx.x * BQ + l) * KMAX + kk] = (bi == 0x7fffffff) ? -INFINITY : bs;
        part_i[((long long)blockIdx.x * BQ + l) * KMAX + kk] = (bi == 0x7fffffff) ? -1 : bi;
      }
    }
  }
}

// final merge: one block per list; 1024 threads each own up to ceil(G/1024) per-CTA sorted lists.
__global__ void __launch_bounds__(1024) k_topk_merge(const float* __restrict__ part_s, const int* __restrict__ part_i, int G, int BQ, int K,
                                                     long long* __restrict__ out_i, float* __restrict__ out_s, int* __restrict__ out_n,
                                                     int out_stride) {
  __shared__ float ss[32];
  __shared__ int si[32];
  __shared__ int sl[32];
  int l = blockIdx.x;
  int t = threadIdx.x;
  // each thread merges its own lists lazily: keep one head per owned CTA list (G <= 4096 assumed => <= 4 per thread)
  int heads[4] = {0, 0, 0, 0};
  int found = 0;
  for (int kk = 0; kk < K; kk++) {
    float bs = -INFINITY; int bi = 0x7fffffff; int bo = -1;
#pragma unroll
    for (int o = 0; o < 4; o++) {
      int g = t + o * 1024;
      if (g < G && heads[o] < K) {
        float s = part_s[((long long)g * BQ + l) * KMAX + heads[o]];
        int i = part_i[((long long)g * BQ + l) * KMAX + heads[o]];
        if (i >= 0 && better(s, i, bs, bi)) { bs = s; bi = i; bo = o; }
      }
    }
    float ws = bs; int wi = bi; int wl = t;
    for (int o = 16; o > 0; o >>= 1) {
      float os = __shfl_xor_sync(0xffffffffu, ws, o);
      int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      if (better(os, oi, ws, wi)) { ws = os; wi = oi; wl = ol; }
    }
    if ((t & 31) == 0) { ss[t >> 5] = ws; si[t >> 5] = wi; sl[t >> 5] = wl; }
    __syncthreads();
    if (t < 32) {
      float s2 = ss[t]; int i2 = si[t]; int l2 = sl[t];
      for (int o = 16; o > 0; o >>= 1) {
        float os = __shfl_xor_sync(0xffffffffu, s2, o);
        int oi = __shfl_xor_sync(0xffffffffu, i2, o);
        int ol = __shfl_xor_sync(0xffffffffu, l2, o);
        if (better(os, oi, s2, i2)) { s2 = os; i2 = oi; l2 = ol; }
      }
      if (t == 0) { ss[0] = s2; si[0] = i2; sl[0] = l2; }
    }
    __syncthreads();
    float fs = ss[0]; int fi = si[0]; int fl = sl[0];
    if (fi != 0x7fffffff) {
      if (t == fl) heads[bo]++;
      if (t == 0) { out_i[(long long)l * out_stride + kk] = fi; out_s[(long long)l * out_stride + kk] = fs; }
      found++;
    } else if (t == 0) {
      out_i[(long long)l * out_stride + kk] = -1; out_s[(long long)l * out_stride + kk] = -INFINITY;
    }
    __syncthreads();
  }
  if (t == 0 && out_n) out_n[l] = found;
}

// ======================================================================================
static size_t knn_smem(int BQ, int d) { return (size_t)BQ * d * 4 + (size_t)KNN_WARPS * BQ * KMAX * 8 + (size_t)KNN_WARPS * BQ * 4; }

template <int DV, int BQ, int RW>
static int32_t launch_pass(hmsg_ctx* ctx, KnnState* st, const float* dq, int nq, int K, const uint8_t* dmask, int mode, int query_id, int q_base,
                           int last_pass, float2* rowstate) {
  size_t smem = knn_smem(BQ, 128 * DV);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sim_topk<DV, BQ, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  ctx->prof_begin(PROF_KNN);
  k_sim_topk<DV, BQ, RW><<<st->grid, KNN_TPB, smem, ctx->stream>>>(st->E, st->N, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate,
                                                                    st->part_s, st->part_i);
  ctx->prof_end(PROF_KNN, (double)st->N * st->d * 4.0);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

template <int DV>
static int32_t launch_pass_bq(hmsg_ctx* ctx, KnnState* st, int BQ, const float* dq, int nq, int K, const uint8_t* dmask, int mode, int query_id,
                              int q_base, int last_pass, float2* rowstate) {
  switch (BQ) {
    case 1: return launch_pass<DV, 1, 4>(ctx, st, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 2: return launch_pass<DV, 2, 4>(ctx, st, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 4: return launch_pass<DV, 4, 4>(ctx, st, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 8: return launch_pass<DV, 8, 4>(ctx, st, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    default: return launch_pass<DV, 16, 2>(ctx, st, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
  }
}

static int32_t launch_pass_any(hmsg_ctx* ctx, KnnState* st, int BQ, const float* dq, int nq, int K, const uint8_t* dmask, int mode, int query_id,
                               int q_base, int last_pass, float2* rowstate) {
  switch (st->d / 128) {
    case 1: return launch_pass_bq<1>(ctx, st, BQ, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 2: return launch_pass_bq<2>(ctx, st, BQ, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 4: return launch_pass_bq<4>(ctx, st, BQ, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 6: return launch_pass_bq<6>(ctx, st, BQ, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
    case 8: return launch_pass_bq<8>(ctx, st, BQ, dq, nq, K, dmask, mode, query_id, q_base, last_pass, rowstate);
  }
  return ctx->fail(HMSG_ERR_ARG, "knn: unsupported d (128,256,512,768,1024)");
}

static int pick_bq(int nq) { return nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : 16; }
static int g_knn_bq_override = 0;   // bench/tuning hook: HMSG_KNN_BQ env

int32_t knn_set_option(hmsg_ctx*, const char* key, int value) {
  if (!strcmp(key, "knn_bq")) { g_knn_bq_override = value; return HMSG_OK; }
  return -1;
}

int32_t knn_destroy(hmsg_ctx* ctx) {
  KnnState* st = ctx->knn;
  if (!st) return HMSG_OK;
  free_dev(st->E_owned); free_dev(st->q_dev); free_dev(st->mask_dev); free_dev(st->part_s); free_dev(st->part_i);
  free_dev(st->out_s); free_dev(st->out_i); free_dev(st->out_n); free_dev(st->rowstate);
  free_dev(st->dense); free_dev(st->keys); free_dev(st->sort_tmp);
  delete st;
  ctx->knn = nullptr;
  return HMSG_OK;
}

extern "C" int32_t hmsg_index_set(hmsg_ctx* ctx, const float* E, int64_t N, int32_t d, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!E || N <= 0 || N >= (1LL << 31) || d <= 0 || d % 128 != 0 || d > 1024)
    return ctx->fail(HMSG_ERR_ARG, "hmsg_index_set: need E != NULL, 0 < N < 2^31, d multiple of 128 <= 1024");
  if (!ctx->knn) ctx->knn = new KnnState();
  KnnState* st = ctx->knn;
  free_dev(st->E_owned);
  st->E = nullptr;
  if (on_device == 2) {
    st->E = E;
  } else {
    HMSG_CUDA(cudaMalloc((void**)&st->E_owned, (size_t)N * d * 4));
    HMSG_CUDA(cudaMemcpyAsync(st->E_owned, E, (size_t)N * d * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
    st->E = st->E_owned;
  }
  st->N = N; st->d = d;
  st->grid = ctx->sm_count * 2;
  if (const char* e = getenv("HMSG_KNN_BQ")) g_knn_bq_override = atoi(e);
  int32_t rc;
  if ((rc = ctx->reserve(&st->part_s, &st->part_s_bytes, (size_t)st->grid * 16 * KMAX * 4))) return rc;
  if ((rc = ctx->reserve(&st->part_i, &st->part_i_bytes, (size_t)st->grid * 16 * KMAX * 4))) return rc;
  return HMSG_OK;
}

static int32_t stage_inputs(hmsg_ctx* ctx, KnnState* st, const float* Q, size_t qfloats, const uint8_t* row_mask, int on_device, const float** dq,
                            const uint8_t** dmask) {
  int32_t rc;
  *dq = Q; *dmask = row_mask;
  if (!on_device) {
    if ((rc = ctx->reserve(&st->q_dev, &st->q_bytes, qfloats * 4))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(st->q_dev, Q, qfloats * 4, cudaMemcpyHostToDevice, ctx->stream));
    *dq = st->q_dev;
    if (row_mask) {
      if ((rc = ctx->reserve(&st->mask_dev, &st->mask_bytes, (size_t)st->N))) return rc;
      HMSG_CUDA(cudaMemcpyAsync(st->mask_dev, row_mask, (size_t)st->N, cudaMemcpyHostToDevice, ctx->stream));
      *dmask = st->mask_dev;
    }
  }
  return HMSG_OK;
}

// ------------------------------------------------------------------------------------------------
// k > KMAX: the reference's argsort has no k limit (callers pass top_k = len(objects) to rank a whole
// room).  Dense scores [Qp][N] -> one 64-bit key per row (score descending, row ascending; rows that
// are masked out or - in object mode - not won by query_id get the sentinel) -> CUB radix sort ->
// first k.  Not the hot path: one extra N*8-byte key stream on top of the matvec.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sim_dense(const float* __restrict__ E, long long N, int d, const float* __restrict__ Q, int nq,
                                                   float* __restrict__ out);

__device__ __forceinline__ unsigned int f32_desc_key(float s) {          // larger score -> smaller key
  unsigned int u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                       // ascending-orderable
  return ~u;
}

__global__ void __launch_bounds__(256) k_rank_keys(const float* __restrict__ sc, long long N, int Qp, int mode, int query_id,
                                                   const uint8_t* __restrict__ mask, unsigned long long* __restrict__ keys) {
  long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  unsigned long long key = ~0ull;
  if (!mask || mask[row]) {
    float s = sc[(long long)query_id * N + row];
    bool ok = true;
    if (mode == 1) {                                                     // np.argmax(sim, axis=0) == query_id (first maximum wins)
      for (int q = 0; q < Qp; q++) {
        float o = sc[(long long)q * N + row];
        if (o > s || (o == s && q < query_id)) { ok = false; break; }
      }
    }
    if (ok) key = ((unsigned long long)f32_desc_key(s) << 32) | (unsigned long long)(unsigned int)row;
  }
  keys[row] = key;
}

__global__ void k_rank_emit(const unsigned long long* __restrict__ keys, long long N, const float* __restrict__ sc_q, int k,
                            long long* __restrict__ out_i, float* __restrict__ out_s, int* __restrict__ out_n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0 && out_n) {
    int lo = 0, hi = (int)std::min<long long>(N, (long long)k);          // valid keys sort before the sentinel: binary search the count
    while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] != ~0ull) lo = mid + 1; else hi = mid; }
    *out_n = lo;
  }
  if (j >= k) return;
  unsigned long long key = j < N ? keys[j] : ~0ull;
  if (key == ~0ull) { out_i[j] = -1; out_s[j] = -INFINITY; return; }
  long long row = (long long)(key & 0xffffffffull);
  out_i[j] = row;
  out_s[j] = sc_q[row];
}

// one request: Qp query rows at dq (device), results to oi/os/on (device)
static int32_t ranked_request(hmsg_ctx* ctx, KnnState* st, const float* dq, int Qp, int mode, int query_id, int k, const uint8_t* dmask,
                              long long* oi, float* os, int* on) {
  int32_t rc;
  long long N = st->N;
  if ((rc = ctx->reserve(&st->dense, &st->dense_bytes, (size_t)Qp * N * 4))) return rc;
  if ((rc = ctx->reserve(&st->keys, &st->keys_bytes, (size_t)2 * N * 8))) return rc;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp, st->keys, st->keys + N, (int)N, 0, 64, ctx->stream);
  if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
  long long warps = N * Qp;
  ctx->prof_begin(PROF_KNN);
  k_sim_dense<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(st->E, N, st->d, dq, Qp, st->dense);
  ctx->prof_end(PROF_KNN, (double)N * st->d * 4.0);
  k_rank_keys<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(st->dense, N, Qp, mode, query_id, dmask, st->keys);
  HMSG_CUDA(cub::DeviceRadixSort::SortKeys(st->sort_tmp, tmp, st->keys, st->keys + N, (int)N, 0, 64, ctx->stream));
  k_rank_emit<<<(k + 255) / 256, 256, 0, ctx->stream>>>(st->keys + N, N, st->dense + (size_t)query_id * N, k, oi, os, on);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

extern "C" int32_t hmsg_query_topk(hmsg_ctx* ctx, const float* Q, int32_t nq, int32_t k, const uint8_t* row_mask, int64_t* ids, float* scores,
                                   int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  KnnState* st = ctx->knn;
  if (!st || !st->E) return ctx->fail(HMSG_ERR_STATE, "hmsg_query_topk: call hmsg_index_set first");
  if (!Q || nq <= 0 || k <= 0 || !ids || !scores) return ctx->fail(HMSG_ERR_ARG, "hmsg_query_topk: bad argument (k >= 1)");
  const float* dq; const uint8_t* dmask;
  int32_t rc = stage_inputs(ctx, st, Q, (size_t)nq * st->d, row_mask, on_device, &dq, &dmask);
  if (rc) return rc;
  long long* oi = (long long*)ids; float* os = scores;
  if (!on_device) {
    if ((rc = ctx->reserve(&st->out_i, &st->out_i_bytes, (size_t)nq * k * 8))) return rc;
    if ((rc = ctx->reserve(&st->out_s, &st->out_s_bytes, (size_t)nq * k * 4))) return rc;
    oi = st->out_i; os = st->out_s;
  }
  int BQmax = g_knn_bq_override > 0 ? pick_bq(g_knn_bq_override) : 8;
  if (k > KMAX) {                                   // ranked path, one query at a time
    for (int q0 = 0; q0 < nq; q0++)
      if ((rc = ranked_request(ctx, st, dq + (size_t)q0 * st->d, 1, 0, 0, k, dmask, oi + (size_t)q0 * k, os + (size_t)q0 * k, nullptr))) return rc;
  }
  for (int q0 = 0; k <= KMAX && q0 < nq;) {
    int rem = nq - q0;
    int BQ = std::min(pick_bq(rem), BQmax);
    int cnt = std::min(rem, BQ);
    if ((rc = launch_pass_any(ctx, st, BQ, dq + (size_t)q0 * st->d, cnt, k, dmask, 0, 0, 0, 1, nullptr))) return rc;
    k_topk_merge<<<cnt, 1024, 0, ctx->stream>>>(st->part_s, st->part_i, st->grid, BQ, k, oi + (size_t)q0 * k, os + (size_t)q0 * k, nullptr, k);
    HMSG_LAUNCH_CHECK();
    q0 += cnt;
  }
  if (!on_device) {
    HMSG_CUDA(cudaMemcpyAsync(ids, oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(scores, os, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HMSG_OK;
}

extern "C" int32_t hmsg_query_object(hmsg_ctx* ctx, const float* Q, int32_t n_req, int32_t Qp, int32_t query_id, int32_t k,
                                     const uint8_t* row_mask, int64_t* ids, float* scores, int32_t* n_found, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  KnnState* st = ctx->knn;
  if (!st || !st->E) return ctx->fail(HMSG_ERR_STATE, "hmsg_query_object: call hmsg_index_set first");
  if (!Q || n_req <= 0 || Qp <= 0 || query_id < 0 || query_id >= Qp || k <= 0 || !ids || !scores || !n_found)
    return ctx->fail(HMSG_ERR_ARG, "hmsg_query_object: bad argument (k >= 1, 0 <= query_id < Qp)");
  const float* dq; const uint8_t* dmask;
  int32_t rc = stage_inputs(ctx, st, Q, (size_t)n_req * Qp * st->d, row_mask, on_device, &dq, &dmask);
  if (rc) return rc;
  long long* oi = (long long*)ids; float* os = scores; int* on = n_found;
  if (!on_device) {
    if ((rc = ctx->reserve(&st->out_i, &st->out_i_bytes, (size_t)n_req * k * 8))) return rc;
    if ((rc = ctx->reserve(&st->out_s, &st->out_s_bytes, (size_t)n_req * k * 4))) return rc;
    if ((rc = ctx->reserve(&st->out_n, &st->out_n_bytes, (size_t)n_req * 4))) return rc;
    oi = st->out_i; os = st->out_s; on = st->out_n;
  }
  int passes = (Qp + 15) / 16;
  if (passes > 1 && (rc = ctx->reserve(&st->rowstate, &st->rowstate_bytes, (size_t)st->N * 8))) return rc;
  for (int r = 0; r < n_req; r++) {
    const float* qr = dq + (size_t)r * Qp * st->d;
    if (k > KMAX) {
      if ((rc = ranked_request(ctx, st, qr, Qp, 1, query_id, k, dmask, oi + (size_t)r * k, os + (size_t)r * k, on + r))) return rc;
      continue;
    }
    for (int p = 0; p < passes; p++) {
      int q_base = p * 16;
      int cnt = std::min(16, Qp - q_base);
      int BQ = pick_bq(cnt);
      if ((rc = launch_pass_any(ctx, st, BQ, qr + (size_t)q_base * st->d, cnt, k, dmask, 1, query_id, q_base, p == passes - 1,
                                passes > 1 ? st->rowstate : nullptr)))
        return rc;
    }
    int BQl = pick_bq(std::min(16, Qp - (passes - 1) * 16));
    k_topk_merge<<<1, 1024, 0, ctx->stream>>>(st->part_s, st->part_i, st->grid, BQl, k, oi + (size_t)r * k, os + (size_t)r * k, on + r, k);
    HMSG_LAUNCH_CHECK();
  }
  if (!on_device) {
    HMSG_CUDA(cudaMemcpyAsync(ids, oi, (size_t)n_req * k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(scores, os, (size_t)n_req * k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(n_found, on, (size_t)n_req * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HMSG_OK;
}

// dense similarity for small tables (room / label retrieval: graph.py:1452, :3204, :3250):
// one warp per (query, row): scores[q][row] = dot(Q[q], E[row])
__global__ void __launch_bounds__(256) k_sim_dense(const float* __restrict__ E, long long N, int d, const float* __restrict__ Q, int nq,
                                                   float* __restrict__ out) {
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= N * nq) return;
  int q = (int)(w / N);
  long long row = w % N;
  const float4* e = reinterpret_cast<const float4*>(E + row * d);
  const float4* qq = reinterpret_cast<const float4*>(Q + (long long)q * d);
  float a = 0.f;
  for (int j = lane; j < d / 4; j += 32) {
    float4 x = e[j], y = qq[j];
    a = fmaf(x.x, y.x, a); a = fmaf(x.y, y.y, a); a = fmaf(x.z, y.z, a); a = fmaf(x.w, y.w, a);
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[(long long)q * N + row] = a;
}

extern "C" int32_t hmsg_query_scores(hmsg_ctx* ctx, const float* Q, int32_t nq, float* scores, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  KnnState* st = ctx->knn;
  if (!st || !st->E) return ctx->fail(HMSG_ERR_STATE, "hmsg_query_scores: call hmsg_index_set first");
  if (!Q || nq <= 0 || !scores) return ctx->fail(HMSG_ERR_ARG, "hmsg_query_scores: bad argument");
  const float* dq; const uint8_t* dm;
  int32_t rc = stage_inputs(ctx, st, Q, (size_t)nq * st->d, nullptr, on_device, &dq, &dm);
  if (rc) return rc;
  float* ds = scores;
  if (!on_device) {
    if ((rc = ctx->reserve(&st->out_s, &st->out_s_bytes, (size_t)nq * st->N * 4))) return rc;
    ds = st->out_s;
  }
  long long warps = (long long)st->N * nq;
  k_sim_dense<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(st->E, st->N, st->d, dq, nq, ds);
  HMSG_LAUNCH_CHECK();
  if (!on_device) {
    HMSG_CUDA(cudaMemcpyAsync(scores, ds, (size_t)nq * st->N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HMSG_OK;
}
