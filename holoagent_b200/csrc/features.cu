// features.cu - A5/A6/A7 of the HMSG build path: per-mask feature fusion, the per-pixel
// feature map restricted to node-winning pixels, node feature accumulation, and the
// per-mask 3-D node sets.
// Reference: fsr_vln/perception/models/sam_clip_feats_extractor.py:159-190;
//            fsr_vln/memory/hmsg/graph/graph.py:404-415; dataloader/generic.py:140-190.
//
// The reference materialises a dense [H*W,d] fp32 map (629 MB) per frame, normalises it,
// casts to fp16, copies 315 MB to the host and then keeps ONE row per touched node.  Here
// the row is computed only for the pixel that wins its node (largest row-major pixel index
// among the frame's valid pixels mapping to that node == the single-thread index_put_ order,
// SURVEY H1), straight from the frame's M mask embeddings held in L2.
#include "common.cuh"
#include <algorithm>

#define TPB 256
int32_t geometry_nn_winner(hmsg_ctx* ctx, int64_t frame_begin, int n_frames);

// ----------------------------------------------------------------------------- masks
// maskbits[fb][p][w]: bit m of word w set iff pixel p belongs to mask 32*w+m
__global__ void __launch_bounds__(TPB) k_masks_boxes(const uint16_t* depth, long long frame0, int H, int W, float scale, int M, int MW,
                                                     const int32_t* boxes, uint32_t* maskbits) {
  extern __shared__ int32_t sb[];
  int fb = blockIdx.y;
  for (int i = threadIdx.x; i < M * 4; i += blockDim.x) sb[i] = boxes[(long long)fb * M * 4 + i];
  __syncthreads();
  int HW = H * W;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int y = p / W, x = p - y * W;
  bool ok = __fdiv_rn((float)depth[(frame0 + fb) * HW + p], scale) > 0.0f;
  for (int w = 0; w < MW; w++) {
    uint32_t bits = 0;
    if (ok) {
      int mend = min(32, M - w * 32);
      for (int m = 0; m < mend; m++) {
        const int32_t* b = sb + (w * 32 + m) * 4;
        if (x >= b[0] && x < b[0] + b[2] && y >= b[1] && y < b[1] + b[3]) bits |= 1u << m;
      }
    }
    maskbits[((long long)fb * HW + p) * MW + w] = bits;
  }
}

// label image form: pixel p belongs to mask labels[p] (a panoptic / instance-id image; negative or >= M: no mask)
__global__ void __launch_bounds__(TPB) k_masks_labels(const int8_t* __restrict__ labels, const uint16_t* __restrict__ depth, long long frame0, float scale,
                                                      int HW, int M, int MW, uint32_t* __restrict__ maskbits) {
  int fb = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int l = labels[(long long)fb * HW + p];
  const bool ok = l >= 0 && l < M && __fdiv_rn((float)depth[(frame0 + fb) * HW + p], scale) > 0.0f;
  for (int w = 0; w < MW; w++) maskbits[((long long)fb * HW + p) * MW + w] = (ok && (l >> 5) == w) ? (1u << (l & 31)) : 0u;
}

__global__ void __launch_bounds__(TPB) k_masks_dense(const uint8_t* seg, int HW, int M, int MW, uint32_t* maskbits) {
  int fb = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const uint8_t* s = seg + (long long)fb * M * HW + p;
  for (int w = 0; w < MW; w++) {
    uint32_t bits = 0;
    int mend = min(32, M - w * 32);
    for (int m = 0; m < mend; m++) bits |= (s[(long long)(w * 32 + m) * HW] ? 1u : 0u) << m;
    maskbits[((long long)fb * HW + p) * MW + w] = bits;
  }
}

// ----------------------------------------------------------------------------- A5 fuse
// one block per frame.  feats [n, 2M+1, d]: rows 0..M-1 masked crops, M..2M-1 plain crops,
// row 2M the full frame (F_g).  extractor.py:159-175.
template <int DV>   // d = 128*DV ; a lane owns DV float4
__global__ void __launch_bounds__(TPB) k_fuse(const float* __restrict__ feats, int M, const int32_t* __restrict__ mask_cnt, float w_masked,
                                              float w_plain, float* __restrict__ Fp) {
  extern __shared__ float sm[];      // phi[M]
  const int d = 128 * DV;
  int fb = blockIdx.x;
  // SAM returns a different number of masks per frame; the batch is padded to M slots.  The softmax of
  // extractor.py:168-172 runs over the frame's REAL masks only: padded slots take no part in it.
  const int Mr = mask_cnt[fb];
  const float* base = feats + (long long)fb * (2 * M + 1) * d;
  float* out = Fp + (long long)fb * M * d;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float4 g[DV];
  float gn = 0.f;
#pragma unroll
  for (int j = 0; j < DV; j++) {
    g[j] = reinterpret_cast<const float4*>(base + (long long)(2 * M) * d)[lane + 32 * j];
    gn += g[j].x * g[j].x + g[j].y * g[j].y + g[j].z * g[j].z + g[j].w * g[j].w;
  }
  for (int o = 16; o > 0; o >>= 1) gn += __shfl_xor_sync(0xffffffffu, gn, o);
  gn = fmaxf(sqrtf(gn), 1e-6f);
  // pass 1: F_l = normalize(w*masked + (1-w)*plain) -> stored in out ; phi = cos(F_l, F_g)
  for (int m = wid; m < Mr; m += nw) {
    float4 v[DV];
    float nn = 0.f;
#pragma unroll
    for (int j = 0; j < DV; j++) {
      float4 a = reinterpret_cast<const float4*>(base + (long long)m * d)[lane + 32 * j];
      float4 b = reinterpret_cast<const float4*>(base + (long long)(M + m) * d)[lane + 32 * j];
      v[j].x = __fadd_rn(__fmul_rn(w_masked, a.x), __fmul_rn(w_plain, b.x));
      v[j].y = __fadd_rn(__fmul_rn(w_masked, a.y), __fmul_rn(w_plain, b.y));
      v[j].z = __fadd_rn(__fmul_rn(w_masked, a.z), __fmul_rn(w_plain, b.z));
      v[j].w = __fadd_rn(__fmul_rn(w_masked, a.w), __fmul_rn(w_plain, b.w));
      nn += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    float den = fmaxf(sqrtf(nn), 1e-12f);          // F.normalize eps
    float dot = 0.f, ln = 0.f;
#pragma unroll
    for (int j = 0; j < DV; j++) {
      v[j].x = __fdiv_rn(v[j].x, den); v[j].y = __fdiv_rn(v[j].y, den); v[j].z = __fdiv_rn(v[j].z, den); v[j].w = __fdiv_rn(v[j].w, den);
      dot += v[j].x * g[j].x + v[j].y * g[j].y + v[j].z * g[j].z + v[j].w * g[j].w;
      ln += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
      reinterpret_cast<float4*>(out + (long long)m * d)[lane + 32 * j] = v[j];
    }
    for (int o = 16; o > 0; o >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); ln += __shfl_xor_sync(0xffffffffu, ln, o); }
    if (lane == 0) sm[m] = dot / (fmaxf(sqrtf(ln), 1e-6f) * gn);   // CosineSimilarity(eps=1e-6)
  }
  __syncthreads();
  // softmax over the masks of the frame (dim=0)
  float mx = -INFINITY;
  for (int m = 0; m < Mr; m++) mx = fmaxf(mx, sm[m]);
  float den = 0.f;
  for (int m = 0; m < Mr; m++) den += expf(sm[m] - mx);
  // pass 2: F_p = normalize(w_i*F_g + (1-w_i)*F_l)
  for (int m = wid; m < Mr; m += nw) {
    float wi = expf(sm[m] - mx) / den;
    float om = 1.0f - wi;
    float4 v[DV];
    float nn = 0.f;
#pragma unroll
    for (int j = 0; j < DV; j++) {
      float4 l = reinterpret_cast<const float4*>(out + (long long)m * d)[lane + 32 * j];
      v[j].x = __fadd_rn(__fmul_rn(wi, g[j].x), __fmul_rn(om, l.x));
      v[j].y = __fadd_rn(__fmul_rn(wi, g[j].y), __fmul_rn(om, l.y));
      v[j].z = __fadd_rn(__fmul_rn(wi, g[j].z), __fmul_rn(om, l.z));
      v[j].w = __fadd_rn(__fmul_rn(wi, g[j].w), __fmul_rn(om, l.w));
      nn += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    float dn = fmaxf(sqrtf(nn), 1e-12f);
#pragma unroll
    for (int j = 0; j < DV; j++) {
      v[j].x = __fdiv_rn(v[j].x, dn); v[j].y = __fdiv_rn(v[j].y, dn); v[j].z = __fdiv_rn(v[j].z, dn); v[j].w = __fdiv_rn(v[j].w, dn);
      reinterpret_cast<float4*>(out + (long long)m * d)[lane + 32 * j] = v[j];
    }
  }
}

// ----------------------------------------------------------------------------- A5/A6 scatter
// One frame per launch (frames of a batch are launched in order => the fp32 accumulation
// order per node equals the reference's frame order, and no atomics are needed: a node has
// exactly one winning pixel per frame).  A warp scans 32 pixels, then cooperates on each
// winner: lane owns DV float4 of the d-vector.
__device__ __forceinline__ float round_half(float v) { return __half2float(__float2half_rn(v)); }

template <int DV>
__global__ void __launch_bounds__(TPB) k_scatter(const int32_t* __restrict__ pix_idx, const unsigned long long* __restrict__ win,
                                                 const uint32_t* __restrict__ maskbits, const float* __restrict__ Fp, int HW, int M, int MW,
                                                 uint32_t epoch, float* __restrict__ sum_feats, float* __restrict__ counter) {
  const int d = 128 * DV;
  int lane = threadIdx.x & 31;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int base = warp * 32; base < HW; base += nwarps * 32) {
    int p = base + lane;
    int n = (p < HW) ? pix_idx[p] : -1;
    bool is_win = false;
    if (n >= 0) is_win = (win[n] == (((unsigned long long)epoch << 32) | (unsigned)p));
    unsigned ball = __ballot_sync(0xffffffffu, is_win);
    while (ball) {
      int s = __ffs(ball) - 1;
      ball &= ball - 1;
      int ns = __shfl_sync(0xffffffffu, n, s);
      int ps = base + s;
      float4 acc[DV];
#pragma unroll
      for (int j = 0; j < DV; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      bool any = false;
      for (int w = 0; w < MW; w++) {
        uint32_t bits = __ldg(&maskbits[(long long)ps * MW + w]);
        while (bits) {
          int m = __ffs(bits) - 1 + 32 * w;
          bits &= bits - 1;
          any = true;
          const float4* row = reinterpret_cast<const float4*>(Fp + (long long)m * d);
#pragma unroll
          for (int j = 0; j < DV; j++) {
            float4 v = __ldg(&row[lane + 32 * j]);
            acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;   // outfeat[idx] += F_p[i] in mask order
          }
        }
      }
      if (any) {
        float nn = 0.f;
#pragma unroll
        for (int j = 0; j < DV; j++) nn += acc[j].x * acc[j].x + acc[j].y * acc[j].y + acc[j].z * acc[j].z + acc[j].w * acc[j].w;
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        float den = fmaxf(sqrtf(nn), 1e-12f);       // extractor.py:188 F.normalize
        float4* dst = reinterpret_cast<float4*>(sum_feats + (long long)ns * d);
#pragma unroll
        for (int j = 0; j < DV; j++) {
          float4 cur = dst[lane + 32 * j];
          cur.x += round_half(__fdiv_rn(acc[j].x, den));   // extractor.py:189 .half(); graph.py:410
          cur.y += round_half(__fdiv_rn(acc[j].y, den));
          cur.z += round_half(__fdiv_rn(acc[j].z, den));
          cur.w += round_half(__fdiv_rn(acc[j].w, den));
          dst[lane + 32 * j] = cur;
        }
      }
      if (lane == 0) counter[ns] += 1.0f;          // graph.py:411 (once per frame per node)
    }
  }
}

// Node-major batch scatter: one warp per node walks the batch's frames IN ORDER, so the fp32 sum
// per node is formed in exactly the reference's frame order (load once, add each winning frame's
// fp16-rounded feature, store once).  HBM traffic: one 2 KB RMW per node touched per BATCH (a node
// is typically hit by many consecutive frames) instead of one per frame.
template <int DV>
__global__ void __launch_bounds__(TPB) k_scatter_batch(const unsigned long long* __restrict__ win, int n_frames, long long n_nodes,
                                                       const uint32_t* __restrict__ maskbits, const float* __restrict__ Fp, int HW, int M, int MW,
                                                       uint32_t epoch, float* __restrict__ sum_feats, float* __restrict__ counter) {
  const int d = 128 * DV;
  int lane = threadIdx.x & 31;
  long long node = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (node >= n_nodes) return;
  float4 acc[DV];
  bool loaded = false;
  int cnt = 0;
  float4* dst = reinterpret_cast<float4*>(sum_feats + node * d);
  for (int f0 = 0; f0 < n_frames; f0 += 32) {
    int f = f0 + lane;
    unsigned long long tag = (f < n_frames) ? win[(long long)f * n_nodes + node] : 0ULL;
    unsigned hits = __ballot_sync(0xffffffffu, (uint32_t)(tag >> 32) == epoch);
    while (hits) {
      int s = __ffs(hits) - 1;
      hits &= hits - 1;
      unsigned p = __shfl_sync(0xffffffffu, (unsigned)(tag & 0xffffffffu), s);
      int fr = f0 + s;
      if (!loaded) {
#pragma unroll
        for (int j = 0; j < DV; j++) acc[j] = dst[lane + 32 * j];
        loaded = true;
      }
      cnt++;
      float4 v[DV];
#pragma unroll
      for (int j = 0; j < DV; j++) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      bool any = false;
      const uint32_t* mb = maskbits + ((long long)fr * HW + p) * MW;
      const float* fpf = Fp + (long long)fr * M * d;
      for (int w = 0; w < MW; w++) {
        uint32_t bits = __ldg(&mb[w]);
        while (bits) {
          int m = __ffs(bits) - 1 + 32 * w;
          bits &= bits - 1;
          any = true;
          const float4* row = reinterpret_cast<const float4*>(fpf + (long long)m * d);
#pragma unroll
          for (int j = 0; j < DV; j++) { float4 x = __ldg(&row[lane + 32 * j]); v[j].x += x.x; v[j].y += x.y; v[j].z += x.z; v[j].w += x.w; }
        }
      }
      if (any) {
        float nn = 0.f;
#pragma unroll
        for (int j = 0; j < DV; j++) nn += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        float den = fmaxf(sqrtf(nn), 1e-12f);
#pragma unroll
        for (int j = 0; j < DV; j++) {
          acc[j].x += round_half(__fdiv_rn(v[j].x, den)); acc[j].y += round_half(__fdiv_rn(v[j].y, den));
          acc[j].z += round_half(__fdiv_rn(v[j].z, den)); acc[j].w += round_half(__fdiv_rn(v[j].w, den));
        }
      }
    }
  }
  if (cnt) {
#pragma unroll
    for (int j = 0; j < DV; j++) dst[lane + 32 * j] = acc[j];
    if (lane == 0) {
      float c = counter[node];
      for (int i = 0; i < cnt; i++) c += 1.0f;     // counter[idx] += 1 once per frame (graph.py:411)
      counter[node] = c;
    }
  }
}

__global__ void __launch_bounds__(TPB) k_finalize_feats(const float* __restrict__ sum_feats, const float* __restrict__ counter, long long n, int d,
                                                        float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = n * (d / 4);
  if (i >= total) return;
  long long node = i / (d / 4);
  float c = counter[node];
  if (c == 0.0f) c = 1e-5f;                        // graph.py:413
  float4 v = reinterpret_cast<const float4*>(sum_feats)[i];
  v.x = __fdiv_rn(v.x, c); v.y = __fdiv_rn(v.y, c); v.z = __fdiv_rn(v.z, c); v.w = __fdiv_rn(v.w, c);
  reinterpret_cast<float4*>(out)[i] = v;
}

// pixel bounding boxes of the masks: rectangles clipped to the frame / min-max over the set bits of dense masks
__global__ void k_rects_from_boxes(const int32_t* __restrict__ boxes, int n, int H, int W, int32_t* __restrict__ rect) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t* b = boxes + (long long)i * 4;
  int x0 = max(b[0], 0), y0 = max(b[1], 0), x1 = min(b[0] + b[2], W), y1 = min(b[1] + b[3], H);
  if (x1 <= x0 || y1 <= y0) { x0 = y0 = x1 = y1 = 0; }
  rect[i * 4] = x0; rect[i * 4 + 1] = y0; rect[i * 4 + 2] = x1; rect[i * 4 + 3] = y1;
}
__global__ void k_rects_init(int32_t* __restrict__ rect, int n, int H, int W) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rect[i * 4] = W; rect[i * 4 + 1] = H; rect[i * 4 + 2] = 0; rect[i * 4 + 3] = 0;
}
__global__ void __launch_bounds__(TPB) k_rects_from_bits(const uint32_t* __restrict__ maskbits, int HW, int W, int M, int MW, int32_t* __restrict__ rect) {
  const int fb = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = p / W, x = p - y * W;
  for (int w = 0; w < MW; w++) {
    uint32_t bits = (p < HW) ? maskbits[((long long)fb * HW + p) * MW + w] : 0u;
    uint32_t uni = __reduce_or_sync(0xffffffffu, bits);
    while (uni) {
      const int ml = __ffs(uni) - 1;
      uni &= uni - 1;
      const bool has = (bits >> ml) & 1u;
      const unsigned b = __ballot_sync(0xffffffffu, has);
      if (has) {
        const int xmin = __reduce_min_sync(b, x), xmax = __reduce_max_sync(b, x), ymin = __reduce_min_sync(b, y), ymax = __reduce_max_sync(b, y);
        if ((b & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) {
          int32_t* r = rect + ((long long)fb * M + w * 32 + ml) * 4;
          if (xmin < r[0]) atomicMin(&r[0], xmin);
          if (ymin < r[1]) atomicMin(&r[1], ymin);
          if (xmax + 1 > r[2]) atomicMax(&r[2], xmax + 1);
          if (ymax + 1 > r[3]) atomicMax(&r[3], ymax + 1);
        }
      }
    }
  }
}

__global__ void k_fill_i32(int32_t* a, int n, int32_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// ======================================================================================
// host side
// ======================================================================================
static int32_t ensure_batch(hmsg_ctx* ctx, int n_frames, int M) {
  // masks need the frames only (crops + encoder of a batch may run before the node table exists: the host-fed
  // path encodes early batches while later frames are still crossing PCIe); A4-A7 check for the node table themselves
  if (!ctx->depth) return ctx->fail(HMSG_ERR_STATE, "masks: call hmsg_scene_begin / add frames first");
  if (M <= 0 || M > 1024) return ctx->fail(HMSG_ERR_ARG, "masks: M out of range");
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  int MW = (M + 31) / 32;
  int32_t rc;
  if ((rc = ctx->reserve(&ctx->maskbits, &ctx->maskbits_bytes, (size_t)n_frames * hw * MW * 4))) return rc;
  if ((rc = ctx->reserve(&ctx->pix_idx, &ctx->pix_idx_bytes, (size_t)n_frames * hw * 4))) return rc;
  if ((rc = ctx->reserve(&ctx->mask_cnt, &ctx->mask_cnt_bytes, (size_t)std::max(n_frames, 64) * 4))) return rc;
  if ((rc = ctx->reserve(&ctx->mask_rect, &ctx->mask_rect_bytes, (size_t)n_frames * M * 16))) return rc;
  ctx->batch_M = M; ctx->batch_MW = MW; ctx->batch_n = n_frames;
  ctx->pix_idx_for = -1;
  masks3d_invalidate_scratch(ctx);
  // every slot is a real mask until hmsg_masks_counts says otherwise
  ctx->batch_counts.assign(n_frames, M);
  k_fill_i32<<<(n_frames + TPB - 1) / TPB, TPB, 0, ctx->stream>>>(ctx->mask_cnt, n_frames, M);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

// Ragged SAM output (extractor.py:117-124 returns as many masks as SAM finds): counts[i] <= M real masks in frame
// frame_begin + i of the batch just set with hmsg_masks_*; slots past the count are padding that takes no part in the
// softmax of A5, produces no 3-D mask in A7 and is not appended to the N1 merge list.
extern "C" int32_t hmsg_masks_counts(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, const int32_t* counts) {
  if (!ctx) return HMSG_ERR_ARG;
  if (ctx->batch_begin != frame_begin || ctx->batch_n != n || !counts)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_masks_counts: call right after hmsg_masks_* of the same batch");
  for (int i = 0; i < n; i++)
    if (counts[i] < 0 || counts[i] > ctx->batch_M) return ctx->fail(HMSG_ERR_ARG, "hmsg_masks_counts: count outside [0, M]");
  ctx->batch_counts.assign(counts, counts + n);
  HMSG_CUDA(cudaMemcpyAsync(ctx->mask_cnt, ctx->batch_counts.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  masks3d_invalidate_scratch(ctx);
  return HMSG_OK;
}

extern "C" int32_t hmsg_masks_boxes(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const int32_t* xywh, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (frame_begin < 0 || n <= 0 || frame_begin + n > ctx->nframes || !xywh) return ctx->fail(HMSG_ERR_ARG, "hmsg_masks_boxes: bad argument");
  int32_t rc = ensure_batch(ctx, n, M);
  if (rc) return rc;
  const int32_t* dbox = xywh;
  if (!on_device) {
    if ((rc = ctx->reserve(&ctx->boxes_stage, &ctx->boxes_stage_bytes, (size_t)n * M * 16))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(ctx->boxes_stage, xywh, (size_t)n * M * 16, cudaMemcpyHostToDevice, ctx->stream));
    dbox = ctx->boxes_stage;
  }
  int HW = ctx->cam.H * ctx->cam.W;
  dim3 grid((HW + TPB - 1) / TPB, n);
  ctx->wait_frames(frame_begin, n);
  k_masks_boxes<<<grid, TPB, M * 16, ctx->stream>>>(ctx->depth, frame_begin, ctx->cam.H, ctx->cam.W, ctx->cam.scale, M, ctx->batch_MW, dbox,
                                                    ctx->maskbits);
  HMSG_LAUNCH_CHECK();
  k_rects_from_boxes<<<(n * M + TPB - 1) / TPB, TPB, 0, ctx->stream>>>(dbox, n * M, ctx->cam.H, ctx->cam.W, ctx->mask_rect);
  HMSG_LAUNCH_CHECK();
  ctx->batch_begin = frame_begin;
  return HMSG_OK;
}

extern "C" int32_t hmsg_masks_dense(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const uint8_t* seg, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (frame_begin < 0 || n <= 0 || frame_begin + n > ctx->nframes || !seg) return ctx->fail(HMSG_ERR_ARG, "hmsg_masks_dense: bad argument");
  int32_t rc = ensure_batch(ctx, n, M);
  if (rc) return rc;
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  const uint8_t* dseg = seg;
  if (!on_device) {
    if ((rc = ctx->reserve(&ctx->seg_stage, &ctx->seg_stage_bytes, (size_t)n * M * hw))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(ctx->seg_stage, seg, (size_t)n * M * hw, cudaMemcpyHostToDevice, ctx->stream));
    dseg = ctx->seg_stage;
  }
  dim3 grid((unsigned)((hw + TPB - 1) / TPB), n);
  k_masks_dense<<<grid, TPB, 0, ctx->stream>>>(dseg, (int)hw, M, ctx->batch_MW, ctx->maskbits);
  HMSG_LAUNCH_CHECK();
  k_rects_init<<<(n * M + TPB - 1) / TPB, TPB, 0, ctx->stream>>>(ctx->mask_rect, n * M, ctx->cam.H, ctx->cam.W);
  k_rects_from_bits<<<grid, TPB, 0, ctx->stream>>>(ctx->maskbits, (int)hw, ctx->cam.W, M, ctx->batch_MW, ctx->mask_rect);
  HMSG_LAUNCH_CHECK();
  ctx->batch_begin = frame_begin;
  return HMSG_OK;
}

extern "C" int32_t hmsg_masks_labels(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const int8_t* labels, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (frame_begin < 0 || n <= 0 || frame_begin + n > ctx->nframes || !labels || M > 127) return ctx->fail(HMSG_ERR_ARG, "hmsg_masks_labels: bad argument");
  int32_t rc = ensure_batch(ctx, n, M);
  if (rc) return rc;
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  const int8_t* dl = labels;
  if (!on_device) {
    if ((rc = ctx->reserve(&ctx->seg_stage, &ctx->seg_stage_bytes, (size_t)n * hw))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(ctx->seg_stage, labels, (size_t)n * hw, cudaMemcpyHostToDevice, ctx->stream));
    dl = (const int8_t*)ctx->seg_stage;
  }
  dim3 grid((unsigned)((hw + TPB - 1) / TPB), n);
  ctx->wait_frames(frame_begin, n);
  k_masks_labels<<<grid, TPB, 0, ctx->stream>>>(dl, ctx->depth, frame_begin, ctx->cam.scale, (int)hw, M, ctx->batch_MW, ctx->maskbits);
  HMSG_LAUNCH_CHECK();
  k_rects_init<<<(n * M + TPB - 1) / TPB, TPB, 0, ctx->stream>>>(ctx->mask_rect, n * M, ctx->cam.H, ctx->cam.W);
  k_rects_from_bits<<<grid, TPB, 0, ctx->stream>>>(ctx->maskbits, (int)hw, ctx->cam.W, M, ctx->batch_MW, ctx->mask_rect);
  HMSG_LAUNCH_CHECK();
  ctx->batch_begin = frame_begin;
  return HMSG_OK;
}

extern "C" int32_t hmsg_features_begin(hmsg_ctx* ctx, int32_t d) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_features_begin: call hmsg_radius_filter first");
  if (d <= 0 || d % 128 != 0 || d > 1024) return ctx->fail(HMSG_ERR_ARG, "hmsg_features_begin: d must be a multiple of 128, <= 1024");
  size_t nn = (size_t)std::max<int64_t>(ctx->n_nodes, 1);
  if (nn * d > ctx->feat_cap) {
    free_dev(ctx->sum_feats); free_dev(ctx->counter);
    HMSG_CUDA(cudaMalloc((void**)&ctx->sum_feats, nn * d * 4));
    HMSG_CUDA(cudaMalloc((void**)&ctx->counter, nn * 4));
    ctx->feat_cap = nn * d;
  }
  HMSG_CUDA(cudaMemsetAsync(ctx->sum_feats, 0, nn * d * 4, ctx->stream));
  HMSG_CUDA(cudaMemsetAsync(ctx->counter, 0, nn * 4, ctx->stream));
  ctx->d = d;
  return HMSG_OK;
}

template <int DV>
static int32_t launch_fuse_scatter(hmsg_ctx* ctx, int n, int M, const float* dfeats, float w_masked, float w_plain) {
  int HW = ctx->cam.H * ctx->cam.W;
  k_fuse<DV><<<n, TPB, M * sizeof(float), ctx->stream>>>(dfeats, M, ctx->mask_cnt, w_masked, w_plain, ctx->Fp);
  HMSG_LAUNCH_CHECK();
  ctx->prof_begin(PROF_SCATTER);
  static int per_frame = -1;
  if (per_frame < 0) { const char* e = getenv("HMSG_SCATTER_PER_FRAME"); per_frame = e ? atoi(e) : 0; }
  if (per_frame) {   // one launch per frame (pixel-major); kept for A/B comparison
    int blocks = std::min((HW + TPB - 1) / TPB, ctx->sm_count * 8);
    for (int fb = 0; fb < n; fb++) {
      k_scatter<DV><<<blocks, TPB, 0, ctx->stream>>>(ctx->pix_idx + (size_t)fb * HW, ctx->win + (size_t)fb * ctx->n_nodes,
                                                      ctx->maskbits + (size_t)fb * HW * ctx->batch_MW, ctx->Fp + (size_t)fb * M * 128 * DV, HW, M,
                                                      ctx->batch_MW, ctx->epoch, ctx->sum_feats, ctx->counter);
      HMSG_LAUNCH_CHECK();
    }
  } else if (ctx->n_nodes > 0) {
    long long threads = ctx->n_nodes * 32;
    k_scatter_batch<DV><<<(unsigned)((threads + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->win, n, ctx->n_nodes, ctx->maskbits, ctx->Fp, HW, M,
                                                                                        ctx->batch_MW, ctx->epoch, ctx->sum_feats, ctx->counter);
    HMSG_LAUNCH_CHECK();
  }
  ctx->prof_end(PROF_SCATTER, 0.0);
  return HMSG_OK;
}

// winner tags are (epoch << 32 | pixel): the buffer is cleared only when it is (re)allocated - cudaMalloc hands back
// recycled memory whose stale bits would otherwise beat or alias the live epoch - and when the epoch wraps.
static int32_t win_next_epoch(hmsg_ctx* ctx, int64_t n) {
  const unsigned long long* before = ctx->win;
  const size_t before_bytes = ctx->win_bytes;
  int32_t rc = ctx->reserve(&ctx->win, &ctx->win_bytes, (size_t)n * std::max<int64_t>(ctx->n_nodes, 1) * 8);
  if (rc) return rc;
  if (ctx->win != before || ctx->win_bytes != before_bytes || ctx->epoch == 0 || ctx->epoch == 0xFFFFFFFFu) {
    HMSG_CUDA(cudaMemsetAsync(ctx->win, 0, ctx->win_bytes, ctx->stream));
    if (ctx->epoch == 0xFFFFFFFFu) ctx->epoch = 0;
  }
  ctx->epoch++;
  return HMSG_OK;
}

extern "C" int32_t hmsg_fuse_scatter(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const float* feats, float maskedd_weight,
                                     float* F_p_out, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->sum_feats || ctx->d == 0 || !ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_fuse_scatter: call hmsg_features_begin first");
  if (ctx->batch_begin != frame_begin || ctx->batch_n != n || ctx->batch_M != M)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_fuse_scatter: masks of this batch were not set (hmsg_masks_*)");
  if (!feats) return ctx->fail(HMSG_ERR_ARG, "hmsg_fuse_scatter: null feats");
  int d = ctx->d;
  int32_t rc;
  if ((rc = win_next_epoch(ctx, n))) return rc;
  if ((rc = ctx->reserve(&ctx->Fp, &ctx->Fp_bytes, (size_t)n * M * d * 4))) return rc;
  const float* dfeats = feats;
  size_t fbytes = (size_t)n * (2 * M + 1) * d * 4;
  if (!on_device) {
    if ((rc = ctx->reserve(&ctx->feats_stage, &ctx->feats_stage_bytes, fbytes))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(ctx->feats_stage, feats, fbytes, cudaMemcpyHostToDevice, ctx->stream));
    dfeats = ctx->feats_stage;
  }
  if ((rc = geometry_nn_winner(ctx, frame_begin, n))) return rc;
  ctx->pix_idx_for = frame_begin;
  // numpy: maskedd_weight * a + (1 - maskedd_weight) * b with float32 arrays and a Python float:
  // both scalars are rounded to float32 (extractor.py:159-160)
  float w_masked = maskedd_weight;
  float w_plain = (float)(1.0 - (double)maskedd_weight);
  switch (d / 128) {
    case 1: rc = launch_fuse_scatter<1>(ctx, n, M, dfeats, w_masked, w_plain); break;
    case 2: rc = launch_fuse_scatter<2>(ctx, n, M, dfeats, w_masked, w_plain); break;
    case 4: rc = launch_fuse_scatter<4>(ctx, n, M, dfeats, w_masked, w_plain); break;
    case 6: rc = launch_fuse_scatter<6>(ctx, n, M, dfeats, w_masked, w_plain); break;
    case 8: rc = launch_fuse_scatter<8>(ctx, n, M, dfeats, w_masked, w_plain); break;
    default: return ctx->fail(HMSG_ERR_ARG, "hmsg_fuse_scatter: unsupported d (128,256,512,768,1024)");
  }
  if (rc) return rc;
  if (F_p_out) {
    HMSG_CUDA(cudaMemcpyAsync(F_p_out, ctx->Fp, (size_t)n * M * d * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                              ctx->stream));
  }
  if (!on_device) HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_node_feats_finalize(hmsg_ctx* ctx, float* full_feats, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->sum_feats || ctx->d == 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_node_feats_finalize: call hmsg_features_begin first");
  if (!full_feats) return ctx->fail(HMSG_ERR_ARG, "hmsg_node_feats_finalize: null output");
  long long n = ctx->n_nodes;
  int d = ctx->d;
  if (n == 0) return HMSG_OK;
  float* dout = full_feats;
  if (!on_device) {
    int32_t rc = ctx->reserve((char**)&ctx->scratch, &ctx->scratch_bytes, (size_t)n * d * 4);
    if (rc) return rc;
    dout = (float*)ctx->scratch;
  }
  long long total = n * (d / 4);
  k_finalize_feats<<<(unsigned)((total + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->sum_feats, ctx->counter, n, d, dout);
  HMSG_LAUNCH_CHECK();
  if (!on_device) {
    HMSG_CUDA(cudaMemcpyAsync(full_feats, dout, (size_t)n * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HMSG_OK;
}

extern "C" int32_t hmsg_node_feats_raw(hmsg_ctx* ctx, float* sum_features, float* counter) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->sum_feats || ctx->d == 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_node_feats_raw: call hmsg_features_begin first");
  long long n = ctx->n_nodes;
  if (n == 0) return HMSG_OK;
  if (sum_features) HMSG_CUDA(cudaMemcpyAsync(sum_features, ctx->sum_feats, (size_t)n * ctx->d * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (counter) HMSG_CUDA(cudaMemcpyAsync(counter, ctx->counter, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_node_feats_device(hmsg_ctx* ctx, float** sum_features, float** counter, int64_t* n_nodes, int32_t* d) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->sum_feats || ctx->d == 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_node_feats_device: call hmsg_features_begin first");
  if (sum_features) *sum_features = ctx->sum_feats;
  if (counter) *counter = ctx->counter;
  if (n_nodes) *n_nodes = ctx->n_nodes;
  if (d) *d = ctx->d;
  return HMSG_OK;
}

// pixel -> node map of the current mask batch (computed by hmsg_fuse_scatter; recomputed here when a
// caller asks for 3-D masks without having scattered features for this batch)
int32_t features_ensure_pix_idx(hmsg_ctx* ctx) {
  if (ctx->batch_begin < 0) return ctx->fail(HMSG_ERR_STATE, "no mask batch (hmsg_masks_*)");
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "pixel->node map: call hmsg_radius_filter first");
  if (ctx->pix_idx_for == ctx->batch_begin) return HMSG_OK;
  int32_t rc;
  if ((rc = win_next_epoch(ctx, ctx->batch_n))) return rc;
  if ((rc = geometry_nn_winner(ctx, ctx->batch_begin, ctx->batch_n))) return rc;
  ctx->pix_idx_for = ctx->batch_begin;
  return HMSG_OK;
}

// ---- multi-GPU merge: packed partial = [sum_features n*d | counter n | F_p rows]
__global__ void __launch_bounds__(TPB) k_merge_partials(const float* __restrict__ gathered, int world, long long stride, long long count,
                                                        float* __restrict__ sum_feats, float* __restrict__ counter, long long nd) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float a = 0.f;
  for (int r = 0; r < world; r++) a += gathered[(long long)r * stride + i];   // rank order: deterministic
  if (i < nd) sum_feats[i] = a; else counter[i - nd] = a;
}

extern "C" int32_t hmsg_node_feats_pack(hmsg_ctx* ctx, float* dst, const float* Fp_rows, int64_t fp_floats) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->sum_feats || ctx->d == 0 || !dst) return ctx->fail(HMSG_ERR_STATE, "hmsg_node_feats_pack: call hmsg_features_begin first");
  size_t nd = (size_t)ctx->n_nodes * ctx->d;
  HMSG_CUDA(cudaMemcpyAsync(dst, ctx->sum_feats, nd * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(dst + nd, ctx->counter, (size_t)ctx->n_nodes * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  if (Fp_rows && fp_floats > 0)
    HMSG_CUDA(cudaMemcpyAsync(dst + nd + ctx->n_nodes, Fp_rows, (size_t)fp_floats * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_node_feats_merge(hmsg_ctx* ctx, const float* gathered, int32_t world, int64_t stride_floats) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->sum_feats || ctx->d == 0 || !gathered || world < 1) return ctx->fail(HMSG_ERR_STATE, "hmsg_node_feats_merge: bad state/argument");
  long long nd = (long long)ctx->n_nodes * ctx->d, count = nd + ctx->n_nodes;
  if (stride_floats < count) return ctx->fail(HMSG_ERR_ARG, "hmsg_node_feats_merge: stride smaller than the partial");
  if (count == 0) return HMSG_OK;
  k_merge_partials<<<(unsigned)((count + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(gathered, world, stride_floats, count, ctx->sum_feats, ctx->counter, nd);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

// ---- dense per-pixel feature map of ONE frame (extractor.py:177-190), only for API parity of
// extract_feats_per_pixel's first return value: [H*W, d] fp16.  The ingest path never builds it.
template <int DV>
__global__ void __launch_bounds__(TPB) k_pixel_map(const uint32_t* __restrict__ maskbits, const float* __restrict__ Fp, int HW, int MW,
                                                   __half* __restrict__ out) {
  const int d = 128 * DV;
  int lane = threadIdx.x & 31;
  long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= HW) return;
  float4 acc[DV];
#pragma unroll
  for (int j = 0; j < DV; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int w = 0; w < MW; w++) {
    uint32_t bits = maskbits[p * MW + w];
    while (bits) {
      int m = __ffs(bits) - 1 + 32 * w;
      bits &= bits - 1;
      const float4* row = reinterpret_cast<const float4*>(Fp + (long long)m * d);
#pragma unroll
      for (int j = 0; j < DV; j++) { float4 v = __ldg(&row[lane + 32 * j]); acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w; }
    }
  }
  float nn = 0.f;
#pragma unroll
  for (int j = 0; j < DV; j++) nn += acc[j].x * acc[j].x + acc[j].y * acc[j].y + acc[j].z * acc[j].z + acc[j].w * acc[j].w;
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  float den = fmaxf(sqrtf(nn), 1e-12f);
#pragma unroll
  for (int j = 0; j < DV; j++) {
    __half2 a = __floats2half2_rn(__fdiv_rn(acc[j].x, den), __fdiv_rn(acc[j].y, den));
    __half2 b = __floats2half2_rn(__fdiv_rn(acc[j].z, den), __fdiv_rn(acc[j].w, den));
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    reinterpret_cast<uint2*>(out + p * d)[lane + 32 * j] = pk;
  }
}

extern "C" int32_t hmsg_pixel_feature_map(hmsg_ctx* ctx, int64_t frame, uint16_t* out_half) {
  if (!ctx) return HMSG_ERR_ARG;
  if (ctx->batch_begin < 0 || frame < ctx->batch_begin || frame >= ctx->batch_begin + ctx->batch_n || !ctx->Fp)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_pixel_feature_map: frame is not in the batch last passed to hmsg_fuse_scatter");
  if (!out_half) return ctx->fail(HMSG_ERR_ARG, "hmsg_pixel_feature_map: null output");
  int HW = ctx->cam.H * ctx->cam.W, d = ctx->d, M = ctx->batch_M, MW = ctx->batch_MW;
  int fb = (int)(frame - ctx->batch_begin);
  int32_t rc = ctx->reserve((char**)&ctx->scratch, &ctx->scratch_bytes, (size_t)HW * d * 2);
  if (rc) return rc;
  __half* dout = (__half*)ctx->scratch;
  const uint32_t* mb = ctx->maskbits + (size_t)fb * HW * MW;
  const float* fp = ctx->Fp + (size_t)fb * M * d;
  unsigned blocks = (unsigned)(((long long)HW * 32 + TPB - 1) / TPB);
  switch (d / 128) {
    case 1: k_pixel_map<1><<<blocks, TPB, 0, ctx->stream>>>(mb, fp, HW, MW, dout); break;
    case 2: k_pixel_map<2><<<blocks, TPB, 0, ctx->stream>>>(mb, fp, HW, MW, dout); break;
    case 4: k_pixel_map<4><<<blocks, TPB, 0, ctx->stream>>>(mb, fp, HW, MW, dout); break;
    case 6: k_pixel_map<6><<<blocks, TPB, 0, ctx->stream>>>(mb, fp, HW, MW, dout); break;
    case 8: k_pixel_map<8><<<blocks, TPB, 0, ctx->stream>>>(mb, fp, HW, MW, dout); break;
    default: return ctx->fail(HMSG_ERR_ARG, "hmsg_pixel_feature_map: unsupported d");
  }
  HMSG_LAUNCH_CHECK();
  HMSG_CUDA(cudaMemcpyAsync(out_half, dout, (size_t)HW * d * 2, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}
