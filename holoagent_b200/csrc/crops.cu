// crops.cu - A8 (N3): mask crops + open_clip preprocessing on device, bit-exact in uint8.
// Reference: fsr_vln/memory/hmsg/utils/sam_utils.py:58-81 (increase_bbox_by_margin), :119-146
// (crop_all_bounding_boxs: cv2.resize(crop,(512,512)) bilinear), :149-181 (crop_image/crop_bbox);
// fsr_vln/memory/hmsg/utils/clip_utils.py:72-73,88-89 (open_clip preprocess: PIL bicubic
// antialiased Resize(224) + CenterCrop + ToTensor + Normalize).
//
// Both resamplers are restated in their own fixed-point arithmetic (validated bit for bit
// against cv2 4.13 / PIL 12.2 in tests/test_oracle.py):
//   cv2 INTER_LINEAR 8U : 11-bit coefficients, horizontal pass in int32, vertical
//                         (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 ; x coefficients are zeroed
//                         at the clamped borders, y coefficients are not (rows are clamped instead)
//   PIL ImagingResample : precompute_coeffs in double (host), 22-bit fixed point, horizontal
//                         pass then vertical pass with clip8 in between
// Per frame the kernels emit [2M+1, 3, 224, 224] fp32: M background-blocked crops, M plain
// crops (bbox + margin), the full frame (extractor.py:147-158 order).
#include "common.cuh"
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#define CROP_MID 512
#define ROWS_PER_BLOCK 16

struct ResampleTable {
  int in_size = 0, out_size = 0, ksize = 0;
  int* bounds = nullptr;   // [out][2] (xmin, count)
  int* kk = nullptr;       // [out][ksize]
};

// Banded PIL pass as int8 tensor-core MMAs (m16n8k32, s32 accumulate - exact).  Eight consecutive outputs
// of the 512 -> 224 antialiased bicubic pass read a window of <= 31 consecutive inputs, so one k=32 MMA covers a
// tile of 8 outputs; the 22-bit fixed-point coefficients are split into three 8-bit digit planes
// (k = d2*65536 + d1*256 + d0, d0/d1 unsigned, d2 signed) and recombined by shifting the accumulator
// between the planes (wrap-around int32 arithmetic is exact because the true sum fits).
struct MmaTable {
  uint32_t* frag = nullptr;   // [28 tiles][3 planes (d2, d1, d0)][32 lanes][2] B fragments
  int* x0 = nullptr;          // [28] window start (multiple of 4)
};

struct CropState {
  std::map<std::pair<int, int>, ResampleTable> tables;
  MmaTable mma;
  uint8_t* t1p = nullptr; size_t t1p_bytes = 0;      // [crops, 3, 224, 512] planar, transposed (dy fastest)
  __half* lut_h = nullptr; float* lut_f = nullptr;   // [3][256] ToTensor + Normalize of a uint8 value
  uint32_t* t1 = nullptr; size_t t1_bytes = 0;       // [crops, rows, 224] packed rgb
  float* out = nullptr;  size_t out_bytes = 0;       // [crops, 3, 224, 224]
  int32_t* boxes = nullptr; size_t boxes_bytes = 0;
};
static std::map<hmsg_ctx*, CropState*> g_crop_states;

// PIL bicubic_filter (a = -0.5)
static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// PIL precompute_coeffs + normalize_coeffs_8bpc (Resample.c), box = whole image
static void pil_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int& ksize) {
  double scale = (double)in_size / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  double support = 2.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; xx++) {
    double center = (xx + 0.5) * scale;
    double ww = 0.0, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; x++) { double w = bicubic_filter((x + xmin - center + 0.5) * ss); k[x] = w; ww += w; }
    for (int x = 0; x < xmax; x++) if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < xmax; x++) {
      double v = k[x];
      kk[(size_t)xx * ksize + x] = (v < 0) ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
    }
    bounds[xx * 2] = xmin; bounds[xx * 2 + 1] = xmax;
  }
}

static int32_t get_table(hmsg_ctx* ctx, CropState* cs, int in_size, int out_size, ResampleTable** out) {
  auto key = std::make_pair(in_size, out_size);
  auto it = cs->tables.find(key);
  if (it == cs->tables.end()) {
    ResampleTable t; t.in_size = in_size; t.out_size = out_size;
    std::vector<int> b, k;
    pil_coeffs(in_size, out_size, b, k, t.ksize);
    HMSG_CUDA(cudaMalloc((void**)&t.bounds, b.size() * 4));
    HMSG_CUDA(cudaMalloc((void**)&t.kk, k.size() * 4));
    HMSG_CUDA(cudaMemcpy(t.bounds, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    HMSG_CUDA(cudaMemcpy(t.kk, k.data(), k.size() * 4, cudaMemcpyHostToDevice));
    it = cs->tables.emplace(key, t).first;
  }
  *out = &it->second;
  return HMSG_OK;
}

// cv2 resize coefficient for destination index d (INTER_LINEAR, 8U): source index + 11-bit pair
__device__ __forceinline__ void cv_coef(int d, double scale, int src, bool clamp_coef, int& s, int& a0, int& a1) {
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_coef) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// ---- pass 1 (crops): cv2 bilinear to 512x512 fused with the PIL horizontal pass 512 -> 224.
// grid = (512/ROWS_PER_BLOCK, 2M, n_frames); T1[crop][dy][ox] = packed r | g<<8 | b<<16.
// A thread owns one output column of the PIL pass and keeps its (<= KS) coefficients in registers.
template <int KS>
__global__ void __launch_bounds__(256) k_crop_rows(const uint8_t* __restrict__ rgb, const uint32_t* __restrict__ maskbits, long long frame0,
                                                   int H, int W, int M, int MW, const int32_t* __restrict__ boxes, int margin,
                                                   const int* __restrict__ bounds, const int* __restrict__ kk, int rows_alloc, uint32_t* __restrict__ t1) {
  __shared__ short s_xofs[CROP_MID];
  __shared__ short s_a[CROP_MID][2];
  __shared__ int s_h[2][CROP_MID * 3];      // horizontally interpolated source rows, already >> 4
  __shared__ uint32_t s_row[CROP_MID + 16];   // +16 zero pad: taps beyond a column's count carry zero coefficients
  __shared__ int s_ry[ROWS_PER_BLOCK][3];   // per destination row: source row index (unclamped), beta0, beta1
  const int fb = blockIdx.z, ci = blockIdx.y;
  const bool masked = ci < M;
  const int m = masked ? ci : ci - M;
  const int32_t* b = boxes + ((long long)fb * M + m) * 4;
  int x = b[0], y = b[1], w = b[2], h = b[3];
  if (!masked) {   // increase_bbox_by_margin (sam_utils.py:58-81)
    x -= margin; y -= margin; w += 2 * margin; h += 2 * margin;
    if (x < 0) { w += x; x = 0; }
    if (y < 0) { h += y; y = 0; }
  }
  // numpy slicing image[y:y+h, x:x+w] truncates at the image border
  int x1 = min(x + w, W), y1 = min(y + h, H);
  x = min(max(x, 0), W); y = min(max(y, 0), H);
  const int cw = x1 - x, ch = y1 - y;
  const long long crop_id = (long long)fb * (2 * M + 1) + ci;
  uint32_t* dst = t1 + crop_id * (long long)rows_alloc * 224;   // the vertical pass strides crops by rows_alloc = max(512, H)
  const int dy0 = blockIdx.x * ROWS_PER_BLOCK;
  if (cw <= 0 || ch <= 0) {   // empty crop: cv2.resize would raise in the reference; emit zeros
    for (int i = threadIdx.x; i < ROWS_PER_BLOCK * 224; i += blockDim.x) dst[(long long)dy0 * 224 + i] = 0;
    return;
  }
  const double scale_x = __ddiv_rn(1.0, __ddiv_rn((double)CROP_MID, (double)cw));
  const double scale_y = __ddiv_rn(1.0, __ddiv_rn((double)CROP_MID, (double)ch));
  for (int d = threadIdx.x; d < CROP_MID; d += blockDim.x) {
    int s, a0, a1;
    cv_coef(d, scale_x, cw, true, s, a0, a1);
    s_xofs[d] = (short)s; s_a[d][0] = (short)a0; s_a[d][1] = (short)a1;
  }
  if (threadIdx.x < 16) s_row[CROP_MID + threadIdx.x] = 0;
  if (threadIdx.x < ROWS_PER_BLOCK) {
    int sy, b0, b1;
    cv_coef(dy0 + threadIdx.x, scale_y, ch, false, sy, b0, b1);
    s_ry[threadIdx.x][0] = sy; s_ry[threadIdx.x][1] = b0; s_ry[threadIdx.x][2] = b1;
  }
  // PIL coefficients of this thread's output column
  const int ox = threadIdx.x;
  int kreg[KS]; int pxmin = 0;
  if (ox < 224) {
    pxmin = __ldg(&bounds[ox * 2]);
#pragma unroll
    for (int k = 0; k < KS; k++) kreg[k] = __ldg(&kk[ox * KS + k]);   // entries beyond pcnt are zero
  }
  __syncthreads();
  const uint8_t* img = rgb + (frame0 + fb) * (long long)H * W * 3;
  const uint32_t* mb = maskbits + (long long)fb * H * W * MW;
  int tag0 = -1, tag1 = -1;   // source rows currently held in s_h[sl0], s_h[sl0^1]
  int sl0 = 0;
  for (int r = 0; r < ROWS_PER_BLOCK; r++) {
    const int dy = dy0 + r;
    const int sy = s_ry[r][0];
    const unsigned b0s = (unsigned)s_ry[r][1] << 16, b1s = (unsigned)s_ry[r][2] << 16;   // (b*h)>>16 == umulhi(b<<16, h)
    const int r0 = min(max(sy, 0), ch - 1), r1 = min(max(sy + 1, 0), ch - 1);
    int need0 = 1, need1 = 1;
    if (tag0 == r0 && tag1 == r1) { need0 = need1 = 0; }
    else if (tag1 == r0) { sl0 ^= 1; tag0 = tag1; tag1 = -1; need0 = 0; }
    for (int which = 0; which < 2; which++) {
      if (which == 0 ? !need0 : !need1) continue;
      const int sr = which == 0 ? r0 : r1;
      int* hb = s_h[which == 0 ? sl0 : (sl0 ^ 1)];
      const uint8_t* srow = img + ((long long)(y + sr) * W + x) * 3;
      const uint32_t* mrow = mb + ((long long)(y + sr) * W + x) * MW;
      for (int d = threadIdx.x; d < CROP_MID; d += blockDim.x) {
        int s0 = s_xofs[d], s1 = min(s0 + 1, cw - 1);
        int a0 = s_a[d][0], a1 = s_a[d][1];
        if (masked) {   // crop_image: image * segmentation (sam_utils.py:159)
          a0 *= (mrow[(long long)s0 * MW + (m >> 5)] >> (m & 31)) & 1;
          a1 *= (mrow[(long long)s1 * MW + (m >> 5)] >> (m & 31)) & 1;
        }
#pragma unroll
        for (int c = 0; c < 3; c++) hb[d * 3 + c] = ((int)srow[s0 * 3 + c] * a0 + (int)srow[s1 * 3 + c] * a1) >> 4;
      }
    }
    tag0 = r0; tag1 = r1;
    __syncthreads();
    const int* h0 = s_h[sl0];
    const int* h1 = s_h[sl0 ^ 1];
    for (int d = threadIdx.x; d < CROP_MID; d += blockDim.x) {
      uint32_t pk = 0;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        int v = (int)(__umulhi(b0s, (unsigned)h0[d * 3 + c]) + __umulhi(b1s, (unsigned)h1[d * 3 + c]) + 2u) >> 2;
        pk |= (uint32_t)(v & 255) << (8 * c);
      }
      s_row[d] = pk;
    }
    __syncthreads();
    if (ox < 224) {   // PIL horizontal pass 512 -> 224
      int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
#pragma unroll
      for (int k = 0; k < KS; k++) {
        uint32_t px = s_row[pxmin + k];
        a0 += (int)(px & 255) * kreg[k]; a1 += (int)((px >> 8) & 255) * kreg[k]; a2 += (int)((px >> 16) & 255) * kreg[k];
      }
      a0 = min(max(a0 >> 22, 0), 255); a1 = min(max(a1 >> 22, 0), 255); a2 = min(max(a2 >> 22, 0), 255);
      dst[(long long)dy * 224 + ox] = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16);
    }
  }
}

// ---- pass 1 (full frame): PIL horizontal pass W -> nw, only columns [left, left+224)
// grid = (H, n_frames); T1[crop][y][ox] packed, `rows_alloc` rows per crop
__global__ void __launch_bounds__(256) k_frame_rows(const uint8_t* __restrict__ rgb, long long frame0, int H, int W, int M, int left,
                                                    const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int rows_alloc,
                                                    uint32_t* __restrict__ t1) {
  const int fb = blockIdx.y, yy = blockIdx.x;
  const uint8_t* srow = rgb + ((frame0 + fb) * (long long)H + yy) * W * 3;
  const long long crop_id = (long long)fb * (2 * M + 1) + 2 * M;
  uint32_t* dst = t1 + crop_id * (long long)rows_alloc * 224 + (long long)yy * 224;
  int ox = threadIdx.x;
  if (ox >= 224) return;
  int sx = ox + left;
  int xmin = __ldg(&bounds[sx * 2]), cnt = __ldg(&bounds[sx * 2 + 1]);
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int k = 0; k < cnt; k++) {
    int kv = __ldg(&kk[sx * ksize + k]);
#pragma unroll
    for (int c = 0; c < 3; c++) acc[c] += (int)srow[(xmin + k) * 3 + c] * kv;
  }
  uint32_t pk = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) pk |= (uint32_t)min(max(acc[c] >> 22, 0), 255) << (8 * c);
  dst[ox] = pk;
}

// ---- pass 2: PIL vertical pass + ToTensor + Normalize -> fp32 planar
// grid = (ceil(224*224/256), crops per frame in this launch, n_frames); crop = frame*(2M+1) + crop0 + blockIdx.y
__global__ void __launch_bounds__(256) k_crop_cols(const uint32_t* __restrict__ t1, int rows_alloc, int top, const int* __restrict__ bounds,
                                                   const int* __restrict__ kk, int ksize, int crop0, int crops_per_frame,
                                                   float* __restrict__ out) {
  const long long crop = (long long)blockIdx.z * crops_per_frame + crop0 + blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= 224 * 224) return;
  int oy = p / 224, ox = p - oy * 224;
  int sy = oy + top;
  int ymin = __ldg(&bounds[sy * 2]), cnt = __ldg(&bounds[sy * 2 + 1]);
  const uint32_t* src = t1 + crop * (long long)rows_alloc * 224 + (long long)ymin * 224 + ox;
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int k = 0; k < cnt; k++) {
    int kv = __ldg(&kk[sy * ksize + k]);
    uint32_t px = __ldg(src + (long long)k * 224);
    acc[0] += (int)(px & 255) * kv; acc[1] += (int)((px >> 8) & 255) * kv; acc[2] += (int)((px >> 16) & 255) * kv;
  }
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    int v = min(max(acc[c] >> 22, 0), 255);
    float f = __fdiv_rn((float)v, 255.0f);                       // ToTensor
    f = __fdiv_rn(__fsub_rn(f, mean[c]), stdv[c]);               // Normalize
    out[(crop * 3 + c) * (long long)(224 * 224) + p] = f;
  }
}

// same pass, but emits the encoder's patch matrix directly: A0[crop*G*G + py*G + px][c*P*P + iy*P + ix] = fp16(value)
// (what k_im2col would produce from the fp32 crop) - saves one 1.2 GB write + read per 32-frame batch.
__global__ void __launch_bounds__(256) k_crop_cols_patches(const uint32_t* __restrict__ t1, int rows_alloc, int top, const int* __restrict__ bounds,
                                                           const int* __restrict__ kk, int ksize, int crop0, int crops_per_frame, int P, int G, int Kpad,
                                                           __half* __restrict__ a0) {
  const long long crop = (long long)blockIdx.z * crops_per_frame + crop0 + blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= 224 * 224) return;
  int oy = p / 224, ox = p - oy * 224;
  int sy = oy + top;
  int ymin = __ldg(&bounds[sy * 2]), cnt = __ldg(&bounds[sy * 2 + 1]);
  const uint32_t* src = t1 + crop * (long long)rows_alloc * 224 + (long long)ymin * 224 + ox;
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int k = 0; k < cnt; k++) {
    int kv = __ldg(&kk[sy * ksize + k]);
    uint32_t px = __ldg(src + (long long)k * 224);
    acc[0] += (int)(px & 255) * kv; acc[1] += (int)((px >> 8) & 255) * kv; acc[2] += (int)((px >> 16) & 255) * kv;
  }
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  const int py = oy / P, iy = oy - py * P, pxx = ox / P, ix = ox - pxx * P;
  __half* dst = a0 + (crop * G * G + py * G + pxx) * (long long)Kpad + iy * P + ix;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    int v = min(max(acc[c] >> 22, 0), 255);
    float f = __fdiv_rn((float)v, 255.0f);
    f = __fdiv_rn(__fsub_rn(f, mean[c]), stdv[c]);
    dst[c * P * P] = __float2half_rn(f);
  }
}

// ------------------------------------------------------------------------------------------
// tensor-core path (default): same arithmetic, the two PIL passes of the 2M mask crops as IMMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void imma_u8s8(int* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void imma_u8u8(int* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// acc[i] = 2^21 + sum_k A[.][k] * coef[k][.] for the three digit planes b[0..5] = (d2, d1, d0) fragments
// (the PIL rounding constant rides along in the second shift: (acc << 8) + 2^21 is one IMAD)
__device__ __forceinline__ void imma_banded(int* acc, const uint32_t* a, const uint32_t* b) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0;
  imma_u8s8(acc, a, b);
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] = (int)((unsigned)acc[i] << 8);
  imma_u8u8(acc, a, b + 2);
#pragma unroll
  for (int i = 0; i < 4; i++) acc[i] = (int)((unsigned)acc[i] * 256u + (1u << 21));
  imma_u8u8(acc, a, b + 4);
}
__device__ __forceinline__ int pil_clip8(int acc) { return min(max(acc >> 22, 0), 255); }

#define PL_LD 560   // 512 + 32 zero pad (+16: row stride of 140 words keeps the 8-row A-fragment loads bank-conflict free): the last tile's 32-wide window may run past column 511 (zero coefficients there)

// pass 1: cv2 bilinear to 512x512 (16 rows per block) into planar uint8 rows in shared memory, then the PIL
// horizontal pass as banded IMMA; output T1P[crop][c][ox / 16][dy / 4][ox % 16][dy % 4]: the vertical pass reads its A
// fragments (4 consecutive dy of one column) as 4-byte words, 8 columns of a tile side by side in one 32-byte sector.
// Bilinear stage: a thread owns two of the 512 columns - its source offsets and 11-bit x coefficients live in
// registers, and so do the horizontally interpolated values of the two source rows the current destination row
// blends (they are reused while the destination rows keep mapping to the same source rows, i.e. whenever the crop
// is magnified).  No shared-memory round trip and no barrier inside the row loop.
__global__ void __launch_bounds__(256) k_crop_rows_mma(const uint8_t* __restrict__ rgb, const uint32_t* __restrict__ maskbits, long long frame0,
                                                       int H, int W, int M, int MW, const int32_t* __restrict__ boxes, int margin,
                                                       const uint32_t* __restrict__ frag, const int* __restrict__ wx0, uint8_t* __restrict__ t1p) {
  __shared__ __align__(16) uint8_t s_pl[3][ROWS_PER_BLOCK][PL_LD];
  __shared__ __align__(16) uint8_t s_out[3 * 224 * 16];
  __shared__ int s_ry[ROWS_PER_BLOCK][3];
  const int fb = blockIdx.z, ci = blockIdx.y;
  const bool masked = ci < M;
  const int m = masked ? ci : ci - M;
  const int32_t* b = boxes + ((long long)fb * M + m) * 4;
  int x = b[0], y = b[1], w = b[2], h = b[3];
  if (!masked) {   // increase_bbox_by_margin (sam_utils.py:58-81)
    x -= margin; y -= margin; w += 2 * margin; h += 2 * margin;
    if (x < 0) { w += x; x = 0; }
    if (y < 0) { h += y; y = 0; }
  }
  int x1 = min(x + w, W), y1 = min(y + h, H);
  x = min(max(x, 0), W); y = min(max(y, 0), H);
  const int cw = x1 - x, ch = y1 - y;
  const long long crop_id = (long long)fb * (2 * M + 1) + ci;
  uint8_t* dst = t1p + crop_id * (long long)(3 * 224 * CROP_MID);
  const int dy0 = blockIdx.x * ROWS_PER_BLOCK;
  if (cw <= 0 || ch <= 0) {   // empty crop: zeros (the reference's cv2.resize would raise)
    for (int i = threadIdx.x; i < 3 * 14 * 16; i += blockDim.x)
      *reinterpret_cast<uint4*>(dst + ((long long)(i >> 4) * 128 + (dy0 >> 2)) * 64 + (i & 15) * 16) = make_uint4(0, 0, 0, 0);
    return;
  }
  const double scale_x = __ddiv_rn(1.0, __ddiv_rn((double)CROP_MID, (double)cw));
  const double scale_y = __ddiv_rn(1.0, __ddiv_rn((double)CROP_MID, (double)ch));
  int cs0[2], cs1[2], cm0[2], cm1[2], ca0[2], ca1[2];   // byte offsets (3 per pixel), mask word offsets, 11-bit coefficients
#pragma unroll
  for (int q = 0; q < 2; q++) {
    int sx, a0, a1;
    cv_coef(threadIdx.x + 256 * q, scale_x, cw, true, sx, a0, a1);
    const int sx1 = min(sx + 1, cw - 1);
    cs0[q] = sx * 3; cs1[q] = sx1 * 3; cm0[q] = sx * MW; cm1[q] = sx1 * MW; ca0[q] = a0; ca1[q] = a1;
  }
  for (int i = threadIdx.x; i < 3 * ROWS_PER_BLOCK * 8; i += blockDim.x)   // zero the 32-byte pad of every planar row
    *reinterpret_cast<uint32_t*>(&s_pl[0][0][0] + (i >> 3) * PL_LD + CROP_MID + (i & 7) * 4) = 0u;
  if (threadIdx.x < ROWS_PER_BLOCK) {
    int sy, b0, b1;
    cv_coef(dy0 + threadIdx.x, scale_y, ch, false, sy, b0, b1);
    s_ry[threadIdx.x][0] = sy; s_ry[threadIdx.x][1] = b0; s_ry[threadIdx.x][2] = b1;
  }
  __syncthreads();
  const uint8_t* img = rgb + ((frame0 + fb) * (long long)H * W + (long long)y * W + x) * 3;
  const uint32_t* mb = maskbits + ((long long)fb * H * W + (long long)y * W + x) * MW + (m >> 5);
  const int mbit = m & 31;
  int hp[2][3], hc[2][3];     // interpolated source rows tagp (upper) / tagc (lower), already >> 4
  int tagp = -1, tagc = -1;
  auto load_row = [&](int sr, int (&hh)[2][3]) {
    const uint8_t* srow = img + (long long)sr * W * 3;
    const uint32_t* mrow = mb + (long long)sr * W * MW;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      int a0 = ca0[q], a1 = ca1[q];
      if (masked) {   // crop_image: image * segmentation (sam_utils.py:159)
        a0 *= (__ldg(mrow + cm0[q]) >> mbit) & 1;
        a1 *= (__ldg(mrow + cm1[q]) >> mbit) & 1;
      }
#pragma unroll
      for (int c = 0; c < 3; c++) hh[q][c] = ((int)__ldg(srow + cs0[q] + c) * a0 + (int)__ldg(srow + cs1[q] + c) * a1) >> 4;
    }
  };
#pragma unroll 1
  for (int r = 0; r < ROWS_PER_BLOCK; r++) {
    const int sy = s_ry[r][0];
    const unsigned b0s = (unsigned)s_ry[r][1] << 16, b1s = (unsigned)s_ry[r][2] << 16;   // (b*h)>>16 == umulhi(b<<16, h)
    const int r0 = min(max(sy, 0), ch - 1), r1 = min(max(sy + 1, 0), ch - 1);
    if (tagp != r0) {
      if (tagc == r0) {
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
          for (int c = 0; c < 3; c++) hp[q][c] = hc[q][c];
      } else {
        load_row(r0, hp);
      }
      tagp = r0;
    }
    if (tagc != r1) { load_row(r1, hc); tagc = r1; }
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int v = (int)(__umulhi(b0s, (unsigned)hp[q][c]) + __umulhi(b1s, (unsigned)hc[q][c]) + 2u) >> 2;
        s_pl[c][r][threadIdx.x + 256 * q] = (uint8_t)v;
      }
  }
  __syncthreads();
  // PIL horizontal pass 512 -> 224: 28 tiles of 8 outputs x 3 channels; A = 16 rows x 32 window bytes, B = coefficient digit planes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // tile j -> warp (j + c) & 7 would balance 84 items as 11/10; simpler and nearly as even: each warp takes whole tiles
  // (B fragments loaded once per tile, reused by the three channels): 28 tiles over 8 warps = 4,4,4,4,3,3,3,3
  for (int j = warp; j < 28; j += 8) {
    const int x0 = __ldg(&wx0[j]);
    uint32_t bf[6];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(frag) + (j * 3 + q) * 32 + lane);
      bf[2 * q] = v.x; bf[2 * q + 1] = v.y;
    }
    const int ox = j * 8 + 2 * t;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      uint32_t a[4];
      const uint8_t* p0 = &s_pl[c][g][x0 + 4 * t];
      a[0] = *reinterpret_cast<const uint32_t*>(p0);      a[1] = *reinterpret_cast<const uint32_t*>(p0 + 8 * PL_LD);
      a[2] = *reinterpret_cast<const uint32_t*>(p0 + 16); a[3] = *reinterpret_cast<const uint32_t*>(p0 + 8 * PL_LD + 16);
      int acc[4];
      imma_banded(acc, a, bf);
      // stage in the T1P tile order: [c][ox / 16][row quad][ox % 16][row % 4]
      uint8_t* o = s_out + ((c * 14 + (ox >> 4)) * 4 + (g >> 2)) * 64 + (ox & 15) * 4 + (g & 3);
      o[0] = (uint8_t)pil_clip8(acc[0]);    o[4] = (uint8_t)pil_clip8(acc[1]);      // (row g, ox), (row g, ox + 1)
      o[128] = (uint8_t)pil_clip8(acc[2]);  o[132] = (uint8_t)pil_clip8(acc[3]);    // rows g + 8: two row quads further
    }
  }
  __syncthreads();
  // 256 contiguous bytes per (channel, 16-column tile): this block's four row quads
  for (int i = threadIdx.x; i < 3 * 14 * 16; i += blockDim.x)
    *reinterpret_cast<uint4*>(dst + ((long long)(i >> 4) * 128 + (dy0 >> 2)) * 64 + (i & 15) * 16) = *reinterpret_cast<const uint4*>(s_out + i * 16);
}

// pass 2: PIL vertical pass 512 -> 224 as banded IMMA + ToTensor + Normalize (256-entry table per channel).
// grid = (7 row bands of 32 output rows, 2M crops, n frames); a warp owns one 8-row output group of the band and half
// of the 14 column tiles.  PATCHES: emit the encoder's fp16 patch matrix (P = 32, G = 7), else fp32 NCHW.
template <bool PATCHES>
__global__ void __launch_bounds__(256) k_crop_cols_mma(const uint8_t* __restrict__ t1p, const uint32_t* __restrict__ frag, const int* __restrict__ wy0,
                                                       const __half* __restrict__ lut_h, const float* __restrict__ lut_f, int crops_per_frame,
                                                       int Kpad, __half* __restrict__ a0, float* __restrict__ out) {
  __shared__ __half s_lh[3][256];
  __shared__ float s_lf[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    if (PATCHES) (&s_lh[0][0])[i] = lut_h[i]; else (&s_lf[0][0])[i] = lut_f[i];
  }
  __syncthreads();
  const long long crop = (long long)blockIdx.z * crops_per_frame + blockIdx.y;
  const uint8_t* src = t1p + crop * (long long)(3 * 224 * CROP_MID);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int band = blockIdx.x;                 // output rows [32*band, 32*band + 32)
  const int grp = band * 4 + (warp & 3);       // 8-row output group (0..27)
  const int y0 = __ldg(&wy0[grp]);
  uint32_t bf[6];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(frag) + (grp * 3 + q) * 32 + lane);
    bf[2 * q] = v.x; bf[2 * q + 1] = v.y;
  }
  const int oy = grp * 8 + 2 * t;              // this thread's output rows: oy, oy + 1
  const int mt0 = (warp >> 2) * 7;
  // A fragments of the next column tile are fetched while the current one is multiplied
  auto load_a = [&](int mt, uint32_t (&a)[3][4]) {
    const uint8_t* p = src + ((long long)mt * 128 + (y0 >> 2) + t) * 64 + g * 4;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const uint8_t* p0 = p + (long long)c * 14 * 128 * 64;
      a[c][0] = __ldg(reinterpret_cast<const uint32_t*>(p0));            a[c][1] = __ldg(reinterpret_cast<const uint32_t*>(p0 + 32));        // columns g, g + 8
      a[c][2] = __ldg(reinterpret_cast<const uint32_t*>(p0 + 4 * 64));   a[c][3] = __ldg(reinterpret_cast<const uint32_t*>(p0 + 4 * 64 + 32)); // dy + 16
    }
  };
  uint32_t acur[3][4], anext[3][4];
  load_a(mt0, acur);
#pragma unroll 1
  for (int mt = mt0; mt < mt0 + 7; mt++) {
    if (mt + 1 < mt0 + 7) load_a(mt + 1, anext);
    const int ox = mt * 16 + g;                // this thread's output columns: ox, ox + 8
#pragma unroll
    for (int c = 0; c < 3; c++) {
      int acc[4];
      imma_banded(acc, acur[c], bf);
      // acc[0]: (ox, oy)  acc[1]: (ox, oy+1)  acc[2]: (ox+8, oy)  acc[3]: (ox+8, oy+1).  Lanes g and g^1 (lane ^ 4) hold
      // neighbouring columns: the even one ends up with the pair of row oy, the odd one with the pair of row oy+1.
#pragma unroll
      for (int hx = 0; hx < 2; hx++) {
        const int v0 = pil_clip8(acc[2 * hx]), v1 = pil_clip8(acc[2 * hx + 1]);
        const int got = __shfl_xor_sync(0xffffffffu, (g & 1) ? v0 : v1, 4);
        const int lo = (g & 1) ? got : v0, hi = (g & 1) ? v1 : got;      // values of columns (xx, xx + 1), xx even
        const int xx = (ox & ~1) + hx * 8, yy = oy + (g & 1);
        if (PATCHES) {
          const int py = yy >> 5, iy = yy & 31, pxx = xx >> 5, ix = xx & 31;
          *reinterpret_cast<__half2*>(a0 + (crop * 49 + py * 7 + pxx) * (long long)Kpad + c * 1024 + iy * 32 + ix) = __halves2half2(s_lh[c][lo], s_lh[c][hi]);
        } else {
          *reinterpret_cast<float2*>(out + (crop * 3 + c) * (long long)(224 * 224) + yy * 224 + xx) = make_float2(s_lf[c][lo], s_lf[c][hi]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int e = 0; e < 4; e++) acur[c][e] = anext[c][e];
  }
}

int32_t crops_destroy(hmsg_ctx* ctx) {
  auto it = g_crop_states.find(ctx);
  if (it == g_crop_states.end()) return HMSG_OK;
  CropState* cs = it->second;
  for (auto& kv : cs->tables) { cudaFree(kv.second.bounds); cudaFree(kv.second.kk); }
  free_dev(cs->t1); free_dev(cs->out); free_dev(cs->boxes); free_dev(cs->t1p); free_dev(cs->lut_h); free_dev(cs->lut_f);
  free_dev(cs->mma.frag); free_dev(cs->mma.x0);
  delete cs;
  g_crop_states.erase(it);
  return HMSG_OK;
}

// B fragments of the banded 512 -> 224 PIL pass (both directions of a mask crop use this table).
// frag [28 tiles][3 planes (d2, d1, d0)][32 lanes][2 regs], x0 [28] window starts.  Returns 0, or 1 / 2 when a tile's
// window exceeds 32 inputs / a coefficient does not fit three 8-bit digits (never for 512 -> 224).
static int build_mma_table(std::vector<uint32_t>& frag, std::vector<int>& x0s) {
  std::vector<int> bounds, kk; int ksize = 0;
  pil_coeffs(CROP_MID, 224, bounds, kk, ksize);
  frag.assign((size_t)28 * 3 * 32 * 2, 0u);
  x0s.assign(28, 0);
  for (int j = 0; j < 28; j++) {
    const int x0 = bounds[(8 * j) * 2] & ~3;
    x0s[j] = x0;
    for (int n = 0; n < 8; n++) {
      const int ox = 8 * j + n, xmin = bounds[ox * 2], cnt = bounds[ox * 2 + 1];
      if (xmin < x0 || xmin + cnt > x0 + 32) return 1;
      for (int k = 0; k < 32; k++) {
        const int idx = x0 + k - xmin;
        const int coef = (idx >= 0 && idx < cnt) ? kk[(size_t)ox * ksize + idx] : 0;
        if (coef >= (1 << 23) || coef < -(1 << 23)) return 2;
        const uint32_t dig[3] = {(uint32_t)((coef >> 16) & 255), (uint32_t)((coef >> 8) & 255), (uint32_t)(coef & 255)};   // d2 (signed), d1, d0
        // B fragment (32x8, "col"): lane = n*4 + t holds k = 4t..4t+3 in reg 0 and k = 16+4t..16+4t+3 in reg 1
        const int t = (k & 15) >> 2, reg = k >> 4, byte = k & 3;
        for (int q = 0; q < 3; q++) frag[(((size_t)j * 3 + q) * 32 + (n * 4 + t)) * 2 + reg] |= dig[q] << (8 * byte);
      }
    }
  }
  return 0;
}

// host-only (no ctx, no GPU): the table above, for the CPU test that replays the banded MMA against real PIL
extern "C" int32_t hmsg_debug_pil_mma_table(uint32_t* frag_out, int32_t* x0_out) {
  if (!frag_out || !x0_out) return HMSG_ERR_ARG;
  std::vector<uint32_t> frag; std::vector<int> x0s;
  if (build_mma_table(frag, x0s)) return HMSG_ERR_STATE;
  memcpy(frag_out, frag.data(), frag.size() * 4);
  for (int j = 0; j < 28; j++) x0_out[j] = x0s[j];
  return HMSG_OK;
}

static int32_t get_mma_table(hmsg_ctx* ctx, CropState* cs) {
  if (cs->mma.frag) return HMSG_OK;
  std::vector<uint32_t> frag; std::vector<int> x0s;
  const int bad = build_mma_table(frag, x0s);
  if (bad == 1) return ctx->fail(HMSG_ERR_STATE, "crops: PIL window of an 8-output tile exceeds 32 inputs");
  if (bad == 2) return ctx->fail(HMSG_ERR_STATE, "crops: PIL coefficient does not fit three 8-bit digits");
  // ToTensor + Normalize of every uint8 value, in the same individually rounded float32 operations as torchvision
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  std::vector<float> lf(768); std::vector<__half> lh(768);
  for (int c = 0; c < 3; c++)
    for (int v = 0; v < 256; v++) {
      volatile float f = (float)v / 255.0f;
      volatile float d = f - mean[c];
      volatile float r = d / stdv[c];
      lf[c * 256 + v] = r; lh[c * 256 + v] = __float2half_rn(r);
    }
  HMSG_CUDA(cudaMalloc((void**)&cs->mma.frag, frag.size() * 4));
  HMSG_CUDA(cudaMalloc((void**)&cs->mma.x0, 28 * 4));
  HMSG_CUDA(cudaMalloc((void**)&cs->lut_f, 768 * 4));
  HMSG_CUDA(cudaMalloc((void**)&cs->lut_h, 768 * 2));
  HMSG_CUDA(cudaMemcpy(cs->mma.frag, frag.data(), frag.size() * 4, cudaMemcpyHostToDevice));
  HMSG_CUDA(cudaMemcpy(cs->mma.x0, x0s.data(), 28 * 4, cudaMemcpyHostToDevice));
  HMSG_CUDA(cudaMemcpy(cs->lut_f, lf.data(), 768 * 4, cudaMemcpyHostToDevice));
  HMSG_CUDA(cudaMemcpy(cs->lut_h, lh.data(), 768 * 2, cudaMemcpyHostToDevice));
  return HMSG_OK;
}

static int g_crops_mma = 1;   // hmsg_set_option("crops_mma", 0) selects the scalar kernels (same results)
int32_t crops_set_option(hmsg_ctx* ctx, const char* key, int value) {
  (void)ctx;
  if (!strcmp(key, "crops_mma")) { g_crops_mma = value; return HMSG_OK; }
  return -1;
}

// a0 != nullptr: emit the fp16 patch matrix (P, G, Kpad describe it) instead of fp32 NCHW crops
int32_t crops_run(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const int32_t* xywh, int32_t bbox_margin, int32_t on_device,
                  float** crops_dev_out, __half* a0, int P, int G, int Kpad) {
  if (ctx->batch_begin != frame_begin || ctx->batch_n != n || ctx->batch_M != M)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_make_crops: masks of this batch were not set (hmsg_masks_*)");
  if (!xywh) return ctx->fail(HMSG_ERR_ARG, "hmsg_make_crops: null argument");
  const int H = ctx->cam.H, W = ctx->cam.W;
  CropState*& cs = g_crop_states[ctx];
  if (!cs) cs = new CropState();
  int32_t rc;
  // open_clip Resize(224): shorter side -> 224, the other int(224 * long / short); CenterCrop(224)
  int nw, nh;
  if (W <= H) { nw = 224; nh = (int)(224.0 * H / W); } else { nh = 224; nw = (int)(224.0 * W / H); }
  if (W == H) { nw = nh = 224; }
  const int left = (int)nearbyint((nw - 224) / 2.0), top = (int)nearbyint((nh - 224) / 2.0);
  ResampleTable *tc, *tw, *th;
  if ((rc = get_table(ctx, cs, CROP_MID, 224, &tc))) return rc;
  if ((rc = get_table(ctx, cs, W, nw, &tw))) return rc;
  if ((rc = get_table(ctx, cs, H, nh, &th))) return rc;
  const int rows_alloc = std::max(CROP_MID, H);
  const long long ncrops = (long long)n * (2 * M + 1);
  if ((rc = ctx->reserve(&cs->t1, &cs->t1_bytes, (size_t)ncrops * rows_alloc * 224 * 4))) return rc;
  if (!a0 && (rc = ctx->reserve(&cs->out, &cs->out_bytes, (size_t)ncrops * 3 * 224 * 224 * 4))) return rc;
  const int32_t* dbox = xywh;
  if (!on_device) {
    if ((rc = ctx->reserve(&cs->boxes, &cs->boxes_bytes, (size_t)n * M * 16))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(cs->boxes, xywh, (size_t)n * M * 16, cudaMemcpyHostToDevice, ctx->stream));
    dbox = cs->boxes;
  }
  // tensor-core path for the 2M mask crops (the patch-matrix form needs the ViT-B/32 geometry P = 32, G = 7)
  const bool use_mma = g_crops_mma != 0 && (!a0 || (P == 32 && G == 7));
  if (use_mma) {
    if ((rc = get_mma_table(ctx, cs))) return rc;
    if ((rc = ctx->reserve(&cs->t1p, &cs->t1p_bytes, (size_t)ncrops * 3 * 224 * CROP_MID + 1024))) return rc;
  }
  ctx->wait_frames(frame_begin, n);
  ctx->prof_begin(PROF_CROPS);
  if (tc->ksize != 11) return ctx->fail(HMSG_ERR_STATE, "hmsg_make_crops: unexpected PIL kernel size for 512->224");
  if (use_mma)
    k_crop_rows_mma<<<dim3(CROP_MID / ROWS_PER_BLOCK, 2 * M, n), 256, 0, ctx->stream>>>(ctx->rgb, ctx->maskbits, frame_begin, H, W, M, ctx->batch_MW, dbox,
                                                                                         bbox_margin, cs->mma.frag, cs->mma.x0, cs->t1p);
  else
    k_crop_rows<11><<<dim3(CROP_MID / ROWS_PER_BLOCK, 2 * M, n), 256, 0, ctx->stream>>>(ctx->rgb, ctx->maskbits, frame_begin, H, W, M, ctx->batch_MW, dbox,
                                                                                         bbox_margin, tc->bounds, tc->kk, rows_alloc, cs->t1);
  HMSG_LAUNCH_CHECK();
  k_frame_rows<<<dim3(H, n), 256, 0, ctx->stream>>>(ctx->rgb, frame_begin, H, W, M, left, tw->bounds, tw->kk, tw->ksize, rows_alloc, cs->t1);
  HMSG_LAUNCH_CHECK();
  const int pb = (224 * 224 + 255) / 256;
  if (a0) {
    if (use_mma)
      k_crop_cols_mma<true><<<dim3(7, 2 * M, n), 256, 0, ctx->stream>>>(cs->t1p, cs->mma.frag, cs->mma.x0, cs->lut_h, cs->lut_f, 2 * M + 1, Kpad, a0, nullptr);
    else
      k_crop_cols_patches<<<dim3(pb, 2 * M, n), 256, 0, ctx->stream>>>(cs->t1, rows_alloc, 0, tc->bounds, tc->kk, tc->ksize, 0, 2 * M + 1, P, G, Kpad, a0);
    HMSG_LAUNCH_CHECK();
    k_crop_cols_patches<<<dim3(pb, 1, n), 256, 0, ctx->stream>>>(cs->t1, rows_alloc, top, th->bounds, th->kk, th->ksize, 2 * M, 2 * M + 1, P, G, Kpad, a0);
    HMSG_LAUNCH_CHECK();
  } else {
    if (use_mma)
      k_crop_cols_mma<false><<<dim3(7, 2 * M, n), 256, 0, ctx->stream>>>(cs->t1p, cs->mma.frag, cs->mma.x0, cs->lut_h, cs->lut_f, 2 * M + 1, 0, nullptr, cs->out);
    else
      k_crop_cols<<<dim3(pb, 2 * M, n), 256, 0, ctx->stream>>>(cs->t1, rows_alloc, 0, tc->bounds, tc->kk, tc->ksize, 0, 2 * M + 1, cs->out);
    HMSG_LAUNCH_CHECK();
    k_crop_cols<<<dim3(pb, 1, n), 256, 0, ctx->stream>>>(cs->t1, rows_alloc, top, th->bounds, th->kk, th->ksize, 2 * M, 2 * M + 1, cs->out);
    HMSG_LAUNCH_CHECK();
  }
  ctx->prof_end(PROF_CROPS, (double)ncrops * 3 * 224 * 224 * (a0 ? 2 : 4));
  if (crops_dev_out) *crops_dev_out = cs->out;
  return HMSG_OK;
}

extern "C" int32_t hmsg_make_crops(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const int32_t* xywh, int32_t bbox_margin,
                                   int32_t on_device, float** crops_dev_out) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!crops_dev_out) return ctx->fail(HMSG_ERR_ARG, "hmsg_make_crops: null argument");
  return crops_run(ctx, frame_begin, n, M, xywh, bbox_margin, on_device, crops_dev_out, nullptr, 0, 0, 0);
}

extern "C" int32_t hmsg_crops_read(hmsg_ctx* ctx, int64_t n_crops, float* host_out) {
  if (!ctx) return HMSG_ERR_ARG;
  auto it = g_crop_states.find(ctx);
  if (it == g_crop_states.end() || !it->second->out) return ctx->fail(HMSG_ERR_STATE, "hmsg_crops_read: call hmsg_make_crops first");
  size_t bytes = (size_t)n_crops * 3 * 224 * 224 * 4;
  if (bytes > it->second->out_bytes) return ctx->fail(HMSG_ERR_ARG, "hmsg_crops_read: more crops requested than were made");
  HMSG_CUDA(cudaMemcpyAsync(host_out, it->second->out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}
