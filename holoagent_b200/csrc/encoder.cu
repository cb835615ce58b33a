// encoder.cu - A9: open_clip ViT visual tower forward (ViT-B/32 for d=512) on sm_100a.
// Reference: fsr_vln/memory/hmsg/utils/clip_utils.py:63-94 (encode_image + F.normalize);
//            model construction fsr_vln/memory/hmsg/graph/graph.py:112-119 (fp16 weights).
//
// All dense contractions (patch embedding, QKV, out-proj, MLP fc/proj, final projection) run
// through ONE persistent warp-specialised GEMM kernel written directly against the Blackwell
// tensor-core path:  TMA (cp.async.bulk.tensor, 128B swizzle) -> 4-stage smem ring ->
// tcgen05.mma (cta_group::1, kind::f16, M=128 N=256 K=16) issued by one elected thread ->
// fp32 accumulators in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1 -> tcgen05.ld epilogue with fused bias / GELU / residual.
//   C[M,N] = A[M,K] (fp16, K-major) x W[N,K]^T (fp16, K-major, torch Linear layout)
// fp16 operands / fp32 accumulate / fp32 residual stream, LayerNorm and softmax in fp32.
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a pipeline bug must surface as a trap (CUDA error), never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint64_t it = 0; it < (1ull << 28); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]   (single-CTA, fp16/bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives TMEM lane (base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) |
//   [32,46) SBO >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ------------------------------------------------------------------------------------------
// GEMM kernel
// ------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KB
constexpr int OUT_STAGE_BYTES = BM * 128;    // 16 KB: 128 rows x one 128-byte swizzle span
constexpr int GEMM_THREADS = 640;            // 4 control warps + 16 epilogue warps
constexpr int COLV_BYTES = 2 * BN * 4;       // per-tile column vectors staged in smem: bias / b' [BN] | sres [BN]
constexpr int GEMM_SMEM = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 2 * OUT_STAGE_BYTES + 512 /*align: checked in the kernel*/ + 256 /*barriers*/ + COLV_BYTES;

// EPI_F16_LN / EPI_F16_LN_GELU: the GEMM consumes the fp16 residual stream x directly; LayerNorm is folded in:
//   LN(x) W^T + b = rstd_r * (x W''^T - mean_r * sres_n) + b'_n,  W''[n,k] = gamma_k W[n,k] - mean_k(gamma W[n,:])  (rows sum to ~0,
//   so the row mean of x drops out of the product; sres_n = what is left of the row sum after fp16 rounding), b' = b + W beta.
//   mean_r / rstd_r come from per-row (sum, sum of squares) accumulated by the epilogue that wrote x.
// EPI_F16_RESID_STATS: x_new = fp16(x_old + acc + bias) written in place (fp16 residual stream, like the reference's
//   precision='fp16' tower, graph.py:117) + the row statistics of the ROUNDED values for the next LayerNorm.
enum { EPI_F16_BIAS = 0, EPI_F16_BIAS_GELU = 1, EPI_F32_RESIDUAL = 2, EPI_F32_STORE = 3, EPI_F16_LN = 4, EPI_F16_LN_GELU = 5, EPI_F16_RESID_STATS = 6 };

struct GemmArgs {
  int M, N, K;
  const float* bias;     // [N] or null
  int quick_gelu;
  // folded-LayerNorm / fp16-residual epilogues
  const float* sres;     // [N] residual row sums of the folded weight
  const float2* stats_in;   // [rows * stat_stride] (mean, rstd) of the A rows (k_rowstats over the parts below)
  float2* stats_out;        // [rows * stat_stride][stat_parts]: every (N tile, column group) owner writes its own part - no
                            // atomics, no zeroing, bit-reproducible; null = not needed
  int stat_parts;           // width / 64
  const __half* resid;      // x_old (same buffer the output tile goes to)
  long long ldr;            // row stride of resid in elements
  int stat_stride;          // row r of this GEMM = stats row r * stat_stride (class-token rows: T)
  float inv_w;              // 1 / width
  float ln_eps;
};

// erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far inside the 1e-3 embedding contract):
// one MUFU.EX2 + one MUFU.RCP + 7 FMA instead of libdevice erff's ~40 instructions
__device__ __forceinline__ float gelu_erf(float x) {
  float z = fabsf(x) * 0.70710678118654752440f;
  float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  float hh = 0.5f * x * (p * __expf(-z * z));   // 0.5 x erfc(|x|/sqrt2)
  return x > 0.f ? x - hh : hh;                 // 0.5 x (1 + erf(x/sqrt2))
}
// packed (2 x fp32) GELU on Blackwell's FFMA2/FMUL2/FADD2 pipes: ~9.5 instructions and one MUFU per element
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  // erfc(z) = 1 / (1 + a1 z + ... + a6 z^6)^16  (Abramowitz-Stegun 7.1.28, |err| < 3e-7), with z = |x|/sqrt(2)
  // folded into the coefficients: one MUFU.RCP per element, everything else on the packed FFMA2/FMUL2 pipe.
  // Branch-free form: gelu(x) = 0.5 x (1 + erf(x/sqrt2)) = max(x, 0) - 0.5 |x| erfc(|x|/sqrt2)  (both signs of x).
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  float2 p = __ffma2_rn(make_float2(5.38297490e-06f, 5.38297490e-06f), ax, make_float2(4.88906371e-05f, 4.88906371e-05f));
  p = __ffma2_rn(p, ax, make_float2(3.80035744e-05f, 3.80035744e-05f));
  p = __ffma2_rn(p, ax, make_float2(3.27762635e-03f, 3.27762635e-03f));
  p = __ffma2_rn(p, ax, make_float2(2.11410057e-02f, 2.11410057e-02f));
  p = __ffma2_rn(p, ax, make_float2(4.98673469e-02f, 4.98673469e-02f));
  p = __ffma2_rn(p, ax, make_float2(1.f, 1.f));
  float2 r = make_float2(rcp_approx(p.x), rcp_approx(p.y));
  r = __fmul2_rn(r, r); r = __fmul2_rn(r, r); r = __fmul2_rn(r, r); r = __fmul2_rn(r, r);      // ^16 = erfc(|x|/sqrt2)
  const float2 t = __fmul2_rn(ax, r);
  return __ffma2_rn(make_float2(-0.5f, -0.5f), t, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float gelu_quick(float x) { return x / (1.0f + __expf(-1.702f * x)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One warp's share of a 128-row x 64-column fp16 output chunk: 32 accumulator columns of TMEM lane `trow` -> bias / folded
// LayerNorm / GELU / residual -> fp16 -> 128B-swizzled staging row.  xo = the thread's 32 x_old halfs (EPI_F16_RESID_STATS).
// sb / ss: this warp's 32 entries of the tile's bias (b') and sres vectors in shared memory (null bias: sb == nullptr)
template <int EPI>
__device__ __forceinline__ void epi_f16_group(const uint32_t* r, const GemmArgs& g, const float* sb, const float* ss, uint8_t* srow, int sw, int half,
                                              float mean, float rstd, const uint4* xo, float& ssum, float& ssq) {
  constexpr bool LN = (EPI == EPI_F16_LN || EPI == EPI_F16_LN_GELU);
  constexpr bool GELU = (EPI == EPI_F16_BIAS_GELU || EPI == EPI_F16_LN_GELU);
  const float* bp = sb;
#pragma unroll
  for (int q4 = 0; q4 < 4; q4++) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = __uint_as_float(r[q4 * 8 + e]);
    if (LN) {
      const float4 s0 = reinterpret_cast<const float4*>(ss)[q4 * 2], s1 = reinterpret_cast<const float4*>(ss)[q4 * 2 + 1];   // LDS, warp-uniform address
      const float4 b0 = reinterpret_cast<const float4*>(bp)[q4 * 2], b1 = reinterpret_cast<const float4*>(bp)[q4 * 2 + 1];
      const float2 nm = make_float2(-mean, -mean), rs = make_float2(rstd, rstd);       // packed FFMA2: two columns per instruction
      float2 t0 = __ffma2_rn(rs, __ffma2_rn(nm, make_float2(s0.x, s0.y), make_float2(v[0], v[1])), make_float2(b0.x, b0.y));
      float2 t1 = __ffma2_rn(rs, __ffma2_rn(nm, make_float2(s0.z, s0.w), make_float2(v[2], v[3])), make_float2(b0.z, b0.w));
      float2 t2 = __ffma2_rn(rs, __ffma2_rn(nm, make_float2(s1.x, s1.y), make_float2(v[4], v[5])), make_float2(b1.x, b1.y));
      float2 t3 = __ffma2_rn(rs, __ffma2_rn(nm, make_float2(s1.z, s1.w), make_float2(v[6], v[7])), make_float2(b1.z, b1.w));
      v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y; v[4] = t2.x; v[5] = t2.y; v[6] = t3.x; v[7] = t3.y;
    } else if (bp) {
      float4 b0 = reinterpret_cast<const float4*>(bp)[q4 * 2], b1 = reinterpret_cast<const float4*>(bp)[q4 * 2 + 1];
      float2 s0 = __fadd2_rn(make_float2(v[0], v[1]), make_float2(b0.x, b0.y)), s1 = __fadd2_rn(make_float2(v[2], v[3]), make_float2(b0.z, b0.w));
      float2 s2 = __fadd2_rn(make_float2(v[4], v[5]), make_float2(b1.x, b1.y)), s3 = __fadd2_rn(make_float2(v[6], v[7]), make_float2(b1.z, b1.w));
      v[0] = s0.x; v[1] = s0.y; v[2] = s1.x; v[3] = s1.y; v[4] = s2.x; v[5] = s2.y; v[6] = s3.x; v[7] = s3.y;
    }
    if (EPI == EPI_F16_RESID_STATS) {
      const uint4 x4 = xo[q4];
      const uint32_t xw[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float2 xf = __half22float2(*reinterpret_cast<const __half2*>(&xw[e]));
        v[2 * e] += xf.x; v[2 * e + 1] += xf.y;
      }
    }
    if (GELU) {
      if (g.quick_gelu) {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = gelu_quick(v[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) { float2 gg = gelu_erf2(make_float2(v[2 * e], v[2 * e + 1])); v[2 * e] = gg.x; v[2 * e + 1] = gg.y; }
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      pk[e] = *reinterpret_cast<uint32_t*>(&hh);
      if (EPI == EPI_F16_RESID_STATS) {        // statistics of the values the next GEMM will actually read
        const float2 hf = __half22float2(hh);
        ssum += hf.x + hf.y;
        ssq = fmaf(hf.x, hf.x, fmaf(hf.y, hf.y, ssq));
      }
    }
    *reinterpret_cast<uint4*>(srow + (((half * 4 + q4) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Epilogue: 16 warps in two groups of 8 (two warps per TMEM lane quarter).  A group converts one
// 128-row x 128-byte chunk of the accumulator tile (64 fp16 / 32 fp32 columns) into its own
// 128B-swizzled staging buffer and one elected thread hands it to TMA (store, or reduce-add for
// the in-place fp32 residual: x += acc + bias happens in L2, the SM never reads x).
template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_f16(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sO = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * OUT_STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;    // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* scol = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  if (threadIdx.x == 0 && (reinterpret_cast<uint8_t*>(scol) + COLV_BYTES) - smem_raw > GEMM_SMEM) asm volatile("trap;");   // window less aligned than assumed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.M + BM - 1) / BM, n_tiles = g.N / BN, k_blocks = g.K / BK;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb = tile / n_tiles, nb = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(sA + s * A_STAGE_BYTES, &tmA, kb * BK, mb * BM, &full[s]);
          tma_load_2d(sB + s * B_STAGE_BYTES, &tmB, kb * BK, nb * BN, &full[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both, N=256, M=128
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < k_blocks; kb++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA + s * A_STAGE_BYTES));
          uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sB + s * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            // advance 16 halfs = 32 B along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[s]);                       // frees the smem slot when these MMAs retire
          if (kb == k_blocks - 1) umma_commit(&tfull[as]);   // accumulator complete -> epilogue
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (16 warps) =====================
    // two groups of 8 warps; in a group two warps share each TMEM lane quarter and split the
    // chunk's columns (a warp loads 32 fp32 columns for fp16 output, 16 for fp32 output)
    constexpr bool F16OUT = (EPI != EPI_F32_RESIDUAL && EPI != EPI_F32_STORE);
    constexpr bool LN = (EPI == EPI_F16_LN || EPI == EPI_F16_LN_GELU);
    constexpr bool RESID = (EPI == EPI_F16_RESID_STATS);
    constexpr int CH_COLS = F16OUT ? 64 : 32;       // columns per 128-byte staging row
    constexpr int WCOLS = CH_COLS / 2;              // columns per warp
    constexpr int NCH = BN / CH_COLS;
    const int ew = warp - 4;
    const int q = ew & 3;                            // == warp % 4: TMEM lane quarter of this warp
    const int half = (ew >> 2) & 1;
    const int grp = ew >> 3;
    uint8_t* stg = sO + grp * OUT_STAGE_BYTES;
    const int trow = q * 32 + lane;                  // row inside the tile == TMEM lane
    const bool issuer = (q == 0 && half == 0 && lane == 0);
    uint8_t* srow = stg + trow * 128;
    const int sw = trow & 7;
    int as = 0; uint32_t aph = 0;
    const int et = threadIdx.x - 128;                // epilogue thread 0..511: entry of the column-vector buffer it fills
    constexpr int ROWS_PER_TILE = BM;
    const int row_in_tile0 = 0;
    const int tile_step = gridDim.x;
    auto load_colv = [&](int nbx) -> float {          // et < BN: bias / b'[nbx * BN + et]; else sres[nbx * BN + et - BN]
      if (et < BN) return g.bias ? __ldg(g.bias + nbx * BN + et) : 0.f;
      return (LN && g.sres) ? __ldg(g.sres + nbx * BN + et - BN) : 0.f;
    };
    float colv = 0.f;
    float2 ms_cur = make_float2(0.f, 0.f), ms_cur_next = make_float2(0.f, 0.f);
    if ((int)blockIdx.x < num_tiles) {
      colv = load_colv((int)blockIdx.x % n_tiles);
      if (LN) { const long long r0 = (long long)((int)blockIdx.x / n_tiles) * BM + trow; if (r0 < g.M) ms_cur = g.stats_in[r0 * g.stat_stride]; }
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb = tile / n_tiles, nb = tile % n_tiles;
      const int row0 = mb * BM;
      const long long grow = (long long)row0 + trow;
      const bool row_ok = grow < g.M;
      float ssum = 0.f, ssq = 0.f;
      const float mean = ms_cur.x, rstd = ms_cur.y;   // LayerNorm statistics of this row (k_rowstats), fetched one tile ahead
      // column vectors of this tile -> shared memory (fetched one tile ahead into `colv`), next tile's prefetches in flight
      scol[et] = colv;
      named_bar_sync(3, 512);
      {
        const int ntile = tile + tile_step;
        if (ntile < num_tiles) {
          const int nmb = ntile / n_tiles, nnb = ntile % n_tiles;
          colv = load_colv(nnb);
          if (LN) { const long long nrow = (long long)nmb * ROWS_PER_TILE + row_in_tile0 + trow; if (nrow < g.M) ms_cur_next = g.stats_in[nrow * g.stat_stride]; }
        }
      }
      uint4 xo[4];
      if (RESID) {                                   // first chunk's x_old: in flight while the MMAs of this tile finish
        const int c0 = nb * BN + grp * CH_COLS + half * WCOLS;
        if (row_ok) {
          const uint4* xp = reinterpret_cast<const uint4*>(g.resid + grow * g.ldr + c0);
#pragma unroll
          for (int i = 0; i < 4; i++) xo[i] = xp[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) xo[i] = make_uint4(0, 0, 0, 0);
        }
      }
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int ch = grp; ch < NCH; ch += 2) {
        if (issuer) tma_wait_read0();                // previous store out of this buffer has drained it
        named_bar_sync(1 + grp, 256);
        const int col0 = nb * BN + ch * CH_COLS;
        uint32_t r[WCOLS];
        if (F16OUT) tmem_ld_32x32(t_addr + (uint32_t)(ch * CH_COLS + half * WCOLS), r);
        else tmem_ld_32x16(t_addr + (uint32_t)(ch * CH_COLS + half * WCOLS), r);
        tmem_ld_wait();
        if constexpr (F16OUT) {
          epi_f16_group<EPI>(r, g, g.bias ? scol + ch * CH_COLS + half * WCOLS : nullptr, scol + BN + ch * CH_COLS + half * WCOLS, srow, sw, half, mean, rstd,
                             xo, ssum, ssq);
          if (RESID && ch + 2 < NCH) {              // next chunk's x_old: overlaps the staging barrier + TMA store of this one
            if (row_ok) {
              const uint4* xp = reinterpret_cast<const uint4*>(g.resid + grow * g.ldr + col0 + 2 * CH_COLS + half * WCOLS);
#pragma unroll
              for (int i = 0; i < 4; i++) xo[i] = xp[i];
            }
          }
        } else {
          const float* bp = g.bias ? scol + ch * CH_COLS + half * WCOLS : nullptr;
#pragma unroll
          for (int q8 = 0; q8 < 4; q8++) {
            float4 v;
            v.x = __uint_as_float(r[q8 * 4 + 0]); v.y = __uint_as_float(r[q8 * 4 + 1]);
            v.z = __uint_as_float(r[q8 * 4 + 2]); v.w = __uint_as_float(r[q8 * 4 + 3]);
            if (bp) { float4 b = reinterpret_cast<const float4*>(bp)[q8]; v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
            *reinterpret_cast<float4*>(srow + (((half * 4 + q8) ^ sw) << 4)) = v;
          }
        }
        fence_proxy_async();                         // generic-proxy smem writes -> visible to the TMA (async proxy)
        named_bar_sync(1 + grp, 256);
        if (issuer) {
          if (EPI == EPI_F32_RESIDUAL) tma_reduce_add_2d(&tmO, stg, col0, row0);
          else tma_store_2d(&tmO, stg, col0, row0);
          tma_commit_group();
        }
      }
      if (RESID && g.stats_out && row_ok)            // this thread's 64 of the row's 256 tile columns = part (nb, grp, half)
        g.stats_out[grow * g.stat_stride * g.stat_parts + nb * 4 + grp * 2 + half] = make_float2(ssum, ssq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
      ms_cur = ms_cur_next;
      named_bar_sync(3, 512);                        // every warp is done reading this tile's column vectors
    }
    if (issuer) tma_wait_all0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a CTA pair (cluster of 2, same TPC) computes a 256x256 tile.
// Each CTA stages its own 128 rows of A and 128 of the 256 B rows, so per k-block a CTA writes
// 32 KB and the tensor core reads 8 KB per MMA from each CTA's shared memory: shared-memory
// traffic per SM drops from ~190 B/clk (1-CTA, over the 128 B/clk port) to ~128 B/clk, which is
// what caps the 1-CTA kernel at ~68 % tensor-pipe activity.  One thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M=256); commits are multicast to both CTAs' barriers; both CTAs run
// their own 16-warp epilogue on their own 128 TMEM lanes.
// ------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6;
constexpr int A2_STAGE_BYTES = 128 * BK * 2;   // 16 KB
constexpr int B2_STAGE_BYTES = 128 * BK * 2;   // 16 KB (this CTA's half of the 256 B rows)
// 227 KB is the opt-in limit: the alignment slack is 512 B here (the dynamic window starts 1024-aligned in practice; checked below)
constexpr int GEMM2_SMEM = STAGES2 * (A2_STAGE_BYTES + B2_STAGE_BYTES) + 2 * OUT_STAGE_BYTES + 512 + 256 + COLV_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  // executed by both CTAs; the peer bit of the barrier address is cleared so the bytes land on the leader's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {   // arrive on the same barrier in CTA `cta` of the cluster
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
k_gemm_f16_2sm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES2 * A2_STAGE_BYTES;
  uint8_t* sO = smem + STAGES2 * (A2_STAGE_BYTES + B2_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * OUT_STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES2]  (used in the leader)
  uint64_t* empty = bars + STAGES2;        // [STAGES2]
  uint64_t* tfull = bars + 2 * STAGES2;    // [2]
  uint64_t* tempty = bars + 2 * STAGES2 + 2;   // [2]  (used in the leader: 32 arrivals = 16 warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);
  float* scol = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  if (threadIdx.x == 0 && (reinterpret_cast<uint8_t*>(scol) + COLV_BYTES) - smem_raw > GEMM2_SMEM) asm volatile("trap;");   // window less aligned than assumed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_tiles = (g.M + 255) / 256, n_tiles = g.N / BN, k_blocks = g.K / BK;
  const int num_tiles = m_tiles * n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES2; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        int mb = tile / n_tiles, nb = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[s], 2 * (A2_STAGE_BYTES + B2_STAGE_BYTES));
          tma_load_2d_2sm(sA + s * A2_STAGE_BYTES, &tmA, kb * BK, mb * 256 + (int)rank * 128, &full[s]);
          tma_load_2d_2sm(sB + s * B2_STAGE_BYTES, &tmB, kb * BK, nb * BN + (int)rank * 128, &full[s]);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      // D=f32, A=B=f16, K-major both, N=256, M=256 (two CTAs x 128 rows)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < k_blocks; kb++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA + s * A2_STAGE_BYTES));
          uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sB + s * B2_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) umma_f16_2sm(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          umma_commit_2sm(&empty[s]);
          if (kb == k_blocks - 1) umma_commit_2sm(&tfull[as]);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (16 warps per CTA, own 128 TMEM lanes) =====================
    constexpr bool F16OUT = (EPI != EPI_F32_RESIDUAL && EPI != EPI_F32_STORE);
    constexpr bool LN = (EPI == EPI_F16_LN || EPI == EPI_F16_LN_GELU);
    constexpr bool RESID = (EPI == EPI_F16_RESID_STATS);
    constexpr int CH_COLS = F16OUT ? 64 : 32;       // columns per 128-byte staging row
    constexpr int WCOLS = CH_COLS / 2;              // columns per warp
    constexpr int NCH = BN / CH_COLS;
    const int ew = warp - 4;
    const int q = ew & 3;                            // == warp % 4: TMEM lane quarter of this warp
    const int half = (ew >> 2) & 1;
    const int grp = ew >> 3;
    uint8_t* stg = sO + grp * OUT_STAGE_BYTES;
    const int trow = q * 32 + lane;                  // row inside the tile == TMEM lane
    const bool issuer = (q == 0 && half == 0 && lane == 0);
    uint8_t* srow = stg + trow * 128;
    const int sw = trow & 7;
    int as = 0; uint32_t aph = 0;
    const int et = threadIdx.x - 128;                // epilogue thread 0..511: entry of the column-vector buffer it fills
    constexpr int ROWS_PER_TILE = 256;
    const int row_in_tile0 = (int)rank * 128;
    const int tile_step = n_clusters;
    auto load_colv = [&](int nbx) -> float {          // et < BN: bias / b'[nbx * BN + et]; else sres[nbx * BN + et - BN]
      if (et < BN) return g.bias ? __ldg(g.bias + nbx * BN + et) : 0.f;
      return (LN && g.sres) ? __ldg(g.sres + nbx * BN + et - BN) : 0.f;
    };
    float colv = 0.f;
    float2 ms_cur = make_float2(0.f, 0.f), ms_cur_next = make_float2(0.f, 0.f);
    if (cluster_id < num_tiles) {
      colv = load_colv(cluster_id % n_tiles);
      if (LN) { const long long r0 = (long long)(cluster_id / n_tiles) * 256 + row_in_tile0 + trow; if (r0 < g.M) ms_cur = g.stats_in[r0 * g.stat_stride]; }
    }
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      int mb = tile / n_tiles, nb = tile % n_tiles;
      const int row0 = mb * 256 + (int)rank * 128;
      const long long grow = (long long)row0 + trow;
      const bool row_ok = grow < g.M;
      float ssum = 0.f, ssq = 0.f;
      const float mean = ms_cur.x, rstd = ms_cur.y;   // LayerNorm statistics of this row (k_rowstats), fetched one tile ahead
      // column vectors of this tile -> shared memory (fetched one tile ahead into `colv`), next tile's prefetches in flight
      scol[et] = colv;
      named_bar_sync(3, 512);
      {
        const int ntile = tile + tile_step;
        if (ntile < num_tiles) {
          const int nmb = ntile / n_tiles, nnb = ntile % n_tiles;
          colv = load_colv(nnb);
          if (LN) { const long long nrow = (long long)nmb * ROWS_PER_TILE + row_in_tile0 + trow; if (nrow < g.M) ms_cur_next = g.stats_in[nrow * g.stat_stride]; }
        }
      }
      uint4 xo[4];
      if (RESID) {                                   // first chunk's x_old: in flight while the MMAs of this tile finish
        const int c0 = nb * BN + grp * CH_COLS + half * WCOLS;
        if (row_ok) {
          const uint4* xp = reinterpret_cast<const uint4*>(g.resid + grow * g.ldr + c0);
#pragma unroll
          for (int i = 0; i < 4; i++) xo[i] = xp[i];
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++) xo[i] = make_uint4(0, 0, 0, 0);
        }
      }
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int ch = grp; ch < NCH; ch += 2) {
        if (issuer) tma_wait_read0();                // previous store out of this buffer has drained it
        named_bar_sync(1 + grp, 256);
        const int col0 = nb * BN + ch * CH_COLS;
        uint32_t r[WCOLS];
        if (F16OUT) tmem_ld_32x32(t_addr + (uint32_t)(ch * CH_COLS + half * WCOLS), r);
        else tmem_ld_32x16(t_addr + (uint32_t)(ch * CH_COLS + half * WCOLS), r);
        tmem_ld_wait();
        if constexpr (F16OUT) {
          epi_f16_group<EPI>(r, g, g.bias ? scol + ch * CH_COLS + half * WCOLS : nullptr, scol + BN + ch * CH_COLS + half * WCOLS, srow, sw, half, mean, rstd,
                             xo, ssum, ssq);
          if (RESID && ch + 2 < NCH) {              // next chunk's x_old: overlaps the staging barrier + TMA store of this one
            if (row_ok) {
              const uint4* xp = reinterpret_cast<const uint4*>(g.resid + grow * g.ldr + col0 + 2 * CH_COLS + half * WCOLS);
#pragma unroll
              for (int i = 0; i < 4; i++) xo[i] = xp[i];
            }
          }
        } else {
          const float* bp = g.bias ? scol + ch * CH_COLS + half * WCOLS : nullptr;
#pragma unroll
          for (int q8 = 0; q8 < 4; q8++) {
            float4 v;
            v.x = __uint_as_float(r[q8 * 4 + 0]); v.y = __uint_as_float(r[q8 * 4 + 1]);
            v.z = __uint_as_float(r[q8 * 4 + 2]); v.w = __uint_as_float(r[q8 * 4 + 3]);
            if (bp) { float4 b = reinterpret_cast<const float4*>(bp)[q8]; v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
            *reinterpret_cast<float4*>(srow + (((half * 4 + q8) ^ sw) << 4)) = v;
          }
        }
        fence_proxy_async();                         // generic-proxy smem writes -> visible to the TMA (async proxy)
        named_bar_sync(1 + grp, 256);
        if (issuer) {
          if (EPI == EPI_F32_RESIDUAL) tma_reduce_add_2d(&tmO, stg, col0, row0);
          else tma_store_2d(&tmO, stg, col0, row0);
          tma_commit_group();
        }
      }
      if (RESID && g.stats_out && row_ok)            // this thread's 64 of the row's 256 tile columns = part (nb, grp, half)
        g.stats_out[grow * g.stat_stride * g.stat_parts + nb * 4 + grp * 2 + half] = make_float2(ssum, ssq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);   // the leader's MMA thread waits for both CTAs' epilogues
      if (++as == 2) { as = 0; aph ^= 1; }
      ms_cur = ms_cur_next;
      named_bar_sync(3, 512);                        // every warp is done reading this tile's column vectors
    }
    if (issuer) tma_wait_all0();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// element-wise / small kernels
// ------------------------------------------------------------------------------------------
// fp32 -> fp16 (weights upload), optional transpose [R,C] -> [C,R]
__global__ void k_f32_to_f16(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
__global__ void k_f32_to_f16_T(const float* __restrict__ in, __half* __restrict__ out, int R, int C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * C) return;
  int r = (int)(i / C), c = (int)(i % C);
  out[(long long)c * R + r] = __float2half_rn(in[i]);
}
// conv weight [W, Kc] -> [W, Kpad] fp16 zero padded
__global__ void k_pad_weight(const float* __restrict__ in, __half* __restrict__ out, int W, int Kc, int Kpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)W * Kpad) return;
  int r = (int)(i / Kpad), c = (int)(i % Kpad);
  out[i] = (c < Kc) ? __float2half_rn(in[(long long)r * Kc + c]) : __float2half_rn(0.f);
}

// im2col for the stride==patch convolution: x [B,3,S,S] fp32 -> A0 [B*G*G, Kpad] fp16,
// column = c*P*P + iy*P + ix (conv1.weight flattening).  One block per (b, c, image row).
__global__ void __launch_bounds__(256) k_im2col(const float* __restrict__ x, __half* __restrict__ a0, long long nrows, int S, int P, int G, int Kpad) {
  // one warp per image row (b, c, y); requires S % 4 == 0 and P % 4 == 0 (float4 in, 4 halfs out)
  long long rowid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // b*3*S + c*S + y
  int lane = threadIdx.x & 31;
  if (rowid >= nrows) return;
  int y = (int)(rowid % S); int c = (int)((rowid / S) % 3); long long b = rowid / (3 * S);
  int py = y / P, iy = y % P;
  if (py >= G) return;
  const float4* src = reinterpret_cast<const float4*>(x + rowid * S);
  for (int t4 = lane; t4 < (G * P) / 4; t4 += 32) {
    float4 v = __ldg(src + t4);
    int t = t4 * 4;
    int px = t / P, ix = t % P;
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(a0 + ((b * G * G + py * G + px)) * Kpad + c * P * P + iy * P + ix) = pk;
  }
}
// any patch size (ViT-L/14: P = 14, Kc = 588 -> Kpad = 640): one thread per A0 element, pad columns zero filled
__global__ void __launch_bounds__(256) k_im2col_generic(const float* __restrict__ x, __half* __restrict__ a0, long long n, int S, int P, int G, int Kc,
                                                        int Kpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int col = (int)(i % Kpad);
  const long long row = i / Kpad;
  float v = 0.f;
  if (col < Kc) {
    const int c = col / (P * P), rem = col - c * P * P, iy = rem / P, ix = rem - iy * P;
    const int pidx = (int)(row % (G * G)); const long long b = row / (G * G);
    const int py = pidx / G, px = pidx - py * G;
    v = __ldg(x + ((b * 3 + c) * S + py * P + iy) * (long long)S + px * P + ix);
  }
  a0[i] = __float2half_rn(v);
}
__global__ void k_zero_pad_cols(__half* a0, long long rows, int Kc, int Kpad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int padw = Kpad - Kc;
  if (i >= rows * padw) return;
  a0[(i / padw) * Kpad + Kc + (i % padw)] = __float2half_rn(0.f);
}

// warp-per-row LayerNorm helpers; W % 128 == 0, W <= 1536
template <int NV>   // NV = W/128 float4 per lane
__device__ __forceinline__ void ln_row(float4* v, const float* __restrict__ gam, const float* __restrict__ bet, int lane, float eps) {
  constexpr int W = NV * 128;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; j++) s += v[j].x + v[j].y + v[j].z + v[j].w;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / (float)W;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  float rstd = rsqrtf(q / (float)W + eps);
#pragma unroll
  for (int j = 0; j < NV; j++) {
    float4 gm = __ldg(reinterpret_cast<const float4*>(gam) + lane + 32 * j);
    float4 bt = __ldg(reinterpret_cast<const float4*>(bet) + lane + 32 * j);
    v[j].x = (v[j].x - mean) * rstd * gm.x + bt.x;
    v[j].y = (v[j].y - mean) * rstd * gm.y + bt.y;
    v[j].z = (v[j].z - mean) * rstd * gm.z + bt.z;
    v[j].w = (v[j].w - mean) * rstd * gm.w + bt.w;
  }
}

// tokens: x[b*T + t] = ln_pre( (t==0 ? cls : patch[b*(T-1)+t-1]) + pos[t] )   -> fp32 residual stream
template <int NV>
__global__ void __launch_bounds__(256) k_embed_lnpre(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                                                     const float* __restrict__ gam, const float* __restrict__ bet, float* __restrict__ x,
                                                     long long rows, int T) {
  constexpr int W = NV * 128;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long b = row / T; int t = (int)(row % T);
  const float4* src = (t == 0) ? reinterpret_cast<const float4*>(cls) : reinterpret_cast<const float4*>(patch + (b * (T - 1) + t - 1) * W);
  const float4* pp = reinterpret_cast<const float4*>(pos + (long long)t * W);
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    float4 a = src[lane + 32 * j], p = __ldg(pp + lane + 32 * j);
    v[j] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
  ln_row<NV>(v, gam, bet, lane, 1e-5f);
#pragma unroll
  for (int j = 0; j < NV; j++) reinterpret_cast<float4*>(x + row * W)[lane + 32 * j] = v[j];
}

// h = LN(x) as fp16 ; row_stride/row_map: used for ln_post on the class-token rows (stride T)
template <int NV>
__global__ void __launch_bounds__(256) k_layernorm_f16(const float* __restrict__ x, const float* __restrict__ gam, const float* __restrict__ bet,
                                                       __half* __restrict__ h, long long rows, long long in_row_stride) {
  constexpr int W = NV * 128;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(x + row * in_row_stride * W);
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) v[j] = src[lane + 32 * j];
  ln_row<NV>(v, gam, bet, lane, 1e-5f);
#pragma unroll
  for (int j = 0; j < NV; j++) {
    __half2 a = __floats2half2_rn(v[j].x, v[j].y), b = __floats2half2_rn(v[j].z, v[j].w);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    reinterpret_cast<uint2*>(h + row * W)[lane + 32 * j] = pk;
  }
}

// ---- folded LayerNorm (see the EPI_F16_LN comment): one warp per output row n of a Linear [N,K] that follows a LayerNorm
__global__ void __launch_bounds__(256) k_fold_ln(const float* __restrict__ Wsrc, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 const float* __restrict__ bias, int N, int K, __half* __restrict__ Wf, float* __restrict__ sres,
                                                 float* __restrict__ bprime) {
  int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* w = Wsrc + (long long)n * K;
  double m = 0.0, bb = 0.0;
  for (int k = lane; k < K; k += 32) { m += (double)gamma[k] * (double)w[k]; bb += (double)beta[k] * (double)w[k]; }
  for (int o = 16; o > 0; o >>= 1) { m += __shfl_xor_sync(0xffffffffu, m, o); bb += __shfl_xor_sync(0xffffffffu, bb, o); }
  const double mean = m / (double)K;
  double rs = 0.0;
  for (int k = lane; k < K; k += 32) {
    const __half h = __float2half_rn((float)((double)gamma[k] * (double)w[k] - mean));
    Wf[(long long)n * K + k] = h;
    rs += (double)__half2float(h);
  }
  for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
  if (lane == 0) { sres[n] = (float)rs; bprime[n] = (float)((double)bias[n] + bb); }
}

// tokens -> fp16 residual stream + (sum, sum of squares) of the rounded row for the first folded LayerNorm
template <int NV>
__global__ void __launch_bounds__(256) k_embed_lnpre_f16(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                                                         const float* __restrict__ gam, const float* __restrict__ bet, __half* __restrict__ x,
                                                         float2* __restrict__ stats, long long rows, int T) {
  constexpr int W = NV * 128;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long b = row / T; int t = (int)(row % T);
  const float4* src = (t == 0) ? reinterpret_cast<const float4*>(cls) : reinterpret_cast<const float4*>(patch + (b * (T - 1) + t - 1) * W);
  const float4* pp = reinterpret_cast<const float4*>(pos + (long long)t * W);
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    float4 a = src[lane + 32 * j], p = __ldg(pp + lane + 32 * j);
    v[j] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
  ln_row<NV>(v, gam, bet, lane, 1e-5f);
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    __half2 a = __floats2half2_rn(v[j].x, v[j].y), c = __floats2half2_rn(v[j].z, v[j].w);
    const float2 fa = __half22float2(a), fc = __half22float2(c);
    s += (fa.x + fa.y) + (fc.x + fc.y);
    q += fa.x * fa.x + fa.y * fa.y + fc.x * fc.x + fc.y * fc.y;
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&c));
    reinterpret_cast<uint2*>(x + row * W)[lane + 32 * j] = pk;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) {
    const float mean = s * (1.0f / (float)W);
    stats[row] = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, q * (1.0f / (float)W)), 0.f) + 1e-5f));
  }
}

// (sum, sumsq) parts written by the residual epilogues -> (mean, rstd) per row, summed in part order (bit-reproducible)
__global__ void __launch_bounds__(256) k_rowstats(const float2* __restrict__ parts, int P, long long rows, long long stride, float inv_w, float eps,
                                                  float2* __restrict__ ms) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float2* sp = parts + r * stride * P;
  float sx = 0.f, sy = 0.f;
  for (int p = 0; p < P; p++) { const float2 v = sp[p]; sx += v.x; sy += v.y; }
  const float mean = sx * inv_w;
  ms[r * stride] = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, sy * inv_w), 0.f) + eps));
}

// h = LN(x) with x in fp16 (ln_post on the class-token rows of the fp16 residual stream)
template <int NV>
__global__ void __launch_bounds__(256) k_layernorm_h2h(const __half* __restrict__ x, const float* __restrict__ gam, const float* __restrict__ bet,
                                                       __half* __restrict__ h, long long rows, long long in_row_stride) {
  constexpr int W = NV * 128;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint2* src = reinterpret_cast<const uint2*>(x + row * in_row_stride * W);
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const uint2 u = src[lane + 32 * j];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), c = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    v[j] = make_float4(a.x, a.y, c.x, c.y);
  }
  ln_row<NV>(v, gam, bet, lane, 1e-5f);
#pragma unroll
  for (int j = 0; j < NV; j++) {
    __half2 a = __floats2half2_rn(v[j].x, v[j].y), b = __floats2half2_rn(v[j].z, v[j].w);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    reinterpret_cast<uint2*>(h + row * W)[lane + 32 * j] = pk;
  }
}

// out[b] = in[b] / max(||in[b]||, 1e-12)   (F.normalize), warp per row, D % 128 == 0
__global__ void __launch_bounds__(256) k_l2norm_rows(const float* __restrict__ in, float* __restrict__ out, int rows, int D, int do_norm) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(in + (long long)row * D);
  float s = 0.f;
  for (int j = lane; j < D / 4; j += 32) { float4 v = src[j]; s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float den = do_norm ? fmaxf(sqrtf(s), 1e-12f) : 1.0f;
  for (int j = lane; j < D / 4; j += 32) {
    float4 v = src[j];
    v.x = __fdiv_rn(v.x, den); v.y = __fdiv_rn(v.y, den); v.z = __fdiv_rn(v.z, den); v.w = __fdiv_rn(v.w, den);
    reinterpret_cast<float4*>(out + (long long)row * D)[j] = v;
  }
}

// ------------------------------------------------------------------------------------------
// attention: T <= 64 tokens, head dim 64.  One warp per (image, head), mma.sync m16n8k16
// (S = Q K^T and O = P V are 50x50x64 - far below a tcgen05 tile; 1.7 % of the model FLOPs).
// ------------------------------------------------------------------------------------------
constexpr int ATT_LD = 72;   // padded row stride in halfs: conflict-free 32-bit fragment loads
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(128) k_attention_mma(const __half* __restrict__ qkv, __half* __restrict__ o, int B, int T, int heads, int W,
                                                       float scale) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long pair = (long long)blockIdx.x * 4 + warp;       // (image, head)
  if (pair >= (long long)B * heads) return;
  int b = (int)(pair / heads), h = (int)(pair % heads);
  __half* sQ = reinterpret_cast<__half*>(att_smem) + (size_t)warp * 3 * 64 * ATT_LD;
  __half* sK = sQ + 64 * ATT_LD;
  __half* sVt = sK + 64 * ATT_LD;      // [dcol][key]
  const int ld = 3 * W;
  // ---- load Q, K (row-major) and V (transposed); rows >= T are zero
  for (int i = lane; i < 64 * 8; i += 32) {
    int r = i >> 3, ch = i & 7;          // 8 x 16-byte chunks per 64-half row
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q;
    if (r < T) {
      const __half* base = qkv + ((long long)b * T + r) * ld + h * 64 + ch * 8;
      q = *reinterpret_cast<const uint4*>(base);
      k = *reinterpret_cast<const uint4*>(base + W);
      v = *reinterpret_cast<const uint4*>(base + 2 * W);
    }
    *reinterpret_cast<uint4*>(sQ + r * ATT_LD + ch * 8) = q;
    *reinterpret_cast<uint4*>(sK + r * ATT_LD + ch * 8) = k;
    const __half* vh = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int e = 0; e < 8; e++) sVt[(ch * 8 + e) * ATT_LD + r] = vh[e];
  }
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
  const int m_tiles = (T + 15) / 16;
  for (int mi = 0; mi < m_tiles; mi++) {
    // ---- S = Q K^T for rows [16mi, 16mi+16), all 64 key columns
    float s[8][4];
#pragma unroll
    for (int ni = 0; ni < 8; ni++) { s[ni][0] = s[ni][1] = s[ni][2] = s[ni][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      uint32_t a[4];
      const __half* qa = sQ + (mi * 16 + g) * ATT_LD + ks * 16 + 2 * t;
      a[0] = *reinterpret_cast<const uint32_t*>(qa);
      a[1] = *reinterpret_cast<const uint32_t*>(qa + 8 * ATT_LD);
      a[2] = *reinterpret_cast<const uint32_t*>(qa + 8);
      a[3] = *reinterpret_cast<const uint32_t*>(qa + 8 * ATT_LD + 8);
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
        uint32_t bb[2];
        const __half* kb = sK + (ni * 8 + g) * ATT_LD + ks * 16 + 2 * t;
        bb[0] = *reinterpret_cast<const uint32_t*>(kb);
        bb[1] = *reinterpret_cast<const uint32_t*>(kb + 8);
        mma_16816(s[ni], a, bb);
      }
    }
    // ---- softmax over keys (rows g and g+8 of this tile); columns >= T masked
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int ni = 0; ni < 8; ni++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        int col = ni * 8 + 2 * t + e;
        float v0 = (col < T) ? s[ni][e] * scale : -INFINITY;
        float v1 = (col < T) ? s[ni][2 + e] * scale : -INFINITY;
        s[ni][e] = v0; s[ni][2 + e] = v1;
        mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int ni = 0; ni < 8; ni++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        float p0 = __expf(s[ni][e] - mx0), p1 = __expf(s[ni][2 + e] - mx1);
        s[ni][e] = p0; s[ni][2 + e] = p1;
        sum0 += p0; sum1 += p1;
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    // ---- O = P V : P (fp32, normalised, rounded to fp16) as A fragments
    float oacc[8][4];
#pragma unroll
    for (int ni = 0; ni < 8; ni++) { oacc[ni][0] = oacc[ni][1] = oacc[ni][2] = oacc[ni][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      uint32_t a[4];
      __half2 h0 = __floats2half2_rn(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
      __half2 h1 = __floats2half2_rn(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
      __half2 h2 = __floats2half2_rn(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
      __half2 h3 = __floats2half2_rn(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
      a[0] = *reinterpret_cast<uint32_t*>(&h0); a[1] = *reinterpret_cast<uint32_t*>(&h1);
      a[2] = *reinterpret_cast<uint32_t*>(&h2); a[3] = *reinterpret_cast<uint32_t*>(&h3);
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
        uint32_t bb[2];
        const __half* vb = sVt + (ni * 8 + g) * ATT_LD + kk * 16 + 2 * t;
        bb[0] = *reinterpret_cast<const uint32_t*>(vb);
        bb[1] = *reinterpret_cast<const uint32_t*>(vb + 8);
        mma_16816(oacc[ni], a, bb);
      }
    }
    int r0 = mi * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ni = 0; ni < 8; ni++) {
      int col = h * 64 + ni * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<__half2*>(o + ((long long)b * T + r0) * W + col) = __floats2half2_rn(oacc[ni][0], oacc[ni][1]);
      if (r1 < T) *reinterpret_cast<__half2*>(o + ((long long)b * T + r1) * W + col) = __floats2half2_rn(oacc[ni][2], oacc[ni][3]);
    }
  }
}

// ---- attention v2: persistent warps, cp.async double-buffered (image, head) tiles, ldmatrix
// fragment loads (V through ldmatrix.trans, so no transposition pass), P kept in registers.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

constexpr int ATT2_WARPS = 4;
constexpr int ATT2_TILE = 64 * ATT_LD;            // halfs per matrix
constexpr int ATT2_SMEM = ATT2_WARPS * 2 * 3 * ATT2_TILE * 2;

__global__ void __launch_bounds__(ATT2_WARPS * 32, 1) k_attention_mma2(const __half* __restrict__ qkv, __half* __restrict__ o, int B, int T, int heads,
                                                                       int W, float scale) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* wbase = reinterpret_cast<__half*>(att_smem) + (size_t)warp * 2 * 3 * ATT2_TILE;
  // zero both buffers once: rows >= T are never written by cp.async and V padding rows must be finite (P = 0 there)
  for (int i = lane; i < 2 * 3 * ATT2_TILE / 8; i += 32) reinterpret_cast<uint4*>(wbase)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const long long npairs = (long long)B * heads;
  const long long gw = (long long)blockIdx.x * ATT2_WARPS + warp, tw = (long long)gridDim.x * ATT2_WARPS;
  const int ld = 3 * W;
  auto issue = [&](long long pair, int bi) {
    int b = (int)(pair / heads), h = (int)(pair % heads);
    __half* sQ = wbase + bi * 3 * ATT2_TILE;
    const __half* src0 = qkv + ((long long)b * T) * ld + h * 64;
    for (int i = lane; i < T * 8; i += 32) {
      int r = i >> 3, ch = i & 7;
      const __half* src = src0 + (long long)r * ld + ch * 8;
      __half* dst = sQ + r * ATT_LD + ch * 8;
      cp_async16(dst, src);
      cp_async16(dst + ATT2_TILE, src + W);
      cp_async16(dst + 2 * ATT2_TILE, src + 2 * W);
    }
  };
  if (gw < npairs) issue(gw, 0);
  cp_async_commit();
  int cur = 0;
  const int g = lane >> 2, t = lane & 3;
  const int m_tiles = (T + 15) / 16;
  for (long long pair = gw; pair < npairs; pair += tw) {
    if (pair + tw < npairs) issue(pair + tw, cur ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const __half* sQ = wbase + cur * 3 * ATT2_TILE;
    const __half* sK = sQ + ATT2_TILE;
    const __half* sV = sK + ATT2_TILE;
    const int b = (int)(pair / heads), h = (int)(pair % heads);
    for (int mi = 0; mi < m_tiles; mi++) {
      float s[8][4];
#pragma unroll
      for (int ni = 0; ni < 8; ni++) { s[ni][0] = s[ni][1] = s[ni][2] = s[ni][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        uint32_t a[4];
        ldsm_x4(a, sQ + (mi * 16 + (lane & 15)) * ATT_LD + ks * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int np = 0; np < 4; np++) {     // two key tiles per ldmatrix.x4
          uint32_t bb[4];
          ldsm_x4(bb, sK + ((np * 2 + (lane >> 4)) * 8 + (lane & 7)) * ATT_LD + ks * 16 + ((lane >> 3) & 1) * 8);
          mma_16816(s[np * 2], a, bb);
          mma_16816(s[np * 2 + 1], a, bb + 2);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int col = ni * 8 + 2 * t + e;
          float v0 = (col < T) ? s[ni][e] * scale : -INFINITY;
          float v1 = (col < T) ? s[ni][2 + e] * scale : -INFINITY;
          s[ni][e] = v0; s[ni][2 + e] = v1;
          mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          float p0 = __expf(s[ni][e] - mx0), p1 = __expf(s[ni][2 + e] - mx1);
          s[ni][e] = p0; s[ni][2 + e] = p1;
          sum0 += p0; sum1 += p1;
        }
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
      float oacc[8][4];
#pragma unroll
      for (int ni = 0; ni < 8; ni++) { oacc[ni][0] = oacc[ni][1] = oacc[ni][2] = oacc[ni][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        uint32_t a[4];
        __half2 h0 = __floats2half2_rn(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
        __half2 h1 = __floats2half2_rn(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
        __half2 h2 = __floats2half2_rn(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
        __half2 h3 = __floats2half2_rn(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
        a[0] = *reinterpret_cast<uint32_t*>(&h0); a[1] = *reinterpret_cast<uint32_t*>(&h1);
        a[2] = *reinterpret_cast<uint32_t*>(&h2); a[3] = *reinterpret_cast<uint32_t*>(&h3);
#pragma unroll
        for (int np = 0; np < 4; np++) {     // two d-column tiles per ldmatrix.x4.trans
          uint32_t bb[4];
          ldsm_x4_t(bb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LD + (np * 2 + (lane >> 4)) * 8);
          mma_16816(oacc[np * 2], a, bb);
          mma_16816(oacc[np * 2 + 1], a, bb + 2);
        }
      }
      const int r0 = mi * 16 + g, r1 = r0 + 8;
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
        int col = h * 64 + ni * 8 + 2 * t;
        if (r0 < T) *reinterpret_cast<__half2*>(o + ((long long)b * T + r0) * W + col) = __floats2half2_rn(oacc[ni][0], oacc[ni][1]);
        if (r1 < T) *reinterpret_cast<__half2*>(o + ((long long)b * T + r1) * W + col) = __floats2half2_rn(oacc[ni][2], oacc[ni][3]);
      }
    }
    __syncwarp();
    cur ^= 1;
  }
  cp_async_wait<0>();
}


// GROUPS (image, head) tiles per block, 2 warps each; NBUF = 2 double-buffers the tile (prefetch while computing),
// NBUF = 1 trades the prefetch for twice as many resident warps.
// KT = number of 8-key tiles that hold valid keys (7 for the 50 tokens of ViT-B/32): score MMAs, softmax terms and
// P fragments of the all-padding tile are not computed at all.
template <int ATT3_GROUPS, int NBUF, int KT, int WPG = 2>
__global__ void __launch_bounds__(ATT3_GROUPS * WPG * 32, 1) k_attention_mma3(const __half* __restrict__ qkv, __half* __restrict__ o, int B, int T, int heads,
                                                                       int W, float scale, int q_tiles) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  // two warps share one (image, head) tile: 8 warps per SM (2 per scheduler) hide ldmatrix / HMMA latency,
  // and each warp handles every other 16-row query tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int GT = WPG * 32;    // threads of a group: WPG warps share one (image, head) tile (4: one 16-row query tile each at T <= 64)
  const int grp = warp / WPG, wsub = warp % WPG, l64 = wsub * 32 + lane;
  __half* wbase = reinterpret_cast<__half*>(att_smem) + (size_t)grp * NBUF * 3 * ATT2_TILE;
  // zero both buffers once: rows >= T are never written by cp.async and V padding rows must be finite (P = 0 there)
  for (int i = l64; i < NBUF * 3 * ATT2_TILE / 8; i += GT) reinterpret_cast<uint4*>(wbase)[i] = make_uint4(0, 0, 0, 0);
  named_bar_sync(1 + grp, GT);
  const long long npairs = (long long)B * heads;
  const long long gw = (long long)blockIdx.x * ATT3_GROUPS + grp, tw = (long long)gridDim.x * ATT3_GROUPS;
  const int ld = 3 * W;
  // (image, head) of a pair index: 32-bit division (B * heads < 2^31 is checked by the launcher; the 64-bit form is a
  // ~100-instruction subroutine that sat at the head of every tile's load chain)
  auto issue = [&](long long pair, int bi) {
    const int b = (int)((unsigned)pair / (unsigned)heads), h = (int)((unsigned)pair - (unsigned)b * (unsigned)heads);
    __half* sQ = wbase + bi * 3 * ATT2_TILE;
    const __half* src0 = qkv + ((long long)b * T) * ld + h * 64;
    for (int i = l64; i < T * 8; i += GT) {
      int r = i >> 3, ch = i & 7;
      const __half* src = src0 + (long long)r * ld + ch * 8;
      __half* dst = sQ + r * ATT_LD + ch * 8;
      cp_async16(dst, src);
      cp_async16(dst + ATT2_TILE, src + W);
      cp_async16(dst + 2 * ATT2_TILE, src + 2 * W);
    }
  };
  if (NBUF == 2) {
    if (gw < npairs) issue(gw, 0);
    cp_async_commit();
  }
  int cur = 0;
  const int g = lane >> 2, t = lane & 3;
  const int m_tiles = min((T + 15) / 16, q_tiles);   // q_tiles = 1: class-token query only (last layer)
  for (long long pair = gw; pair < npairs; pair += tw) {
    if (NBUF == 2) {
      if (pair + tw < npairs) issue(pair + tw, cur ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      issue(pair, 0);
      cp_async_commit();
      cp_async_wait<0>();
    }
    named_bar_sync(1 + grp, GT);       // both warps' cp.async data is visible to both
    const __half* sQ = wbase + cur * 3 * ATT2_TILE;
    const __half* sK = sQ + ATT2_TILE;
    const __half* sV = sK + ATT2_TILE;
    const int b = (int)((unsigned)pair / (unsigned)heads), h = (int)((unsigned)pair - (unsigned)b * (unsigned)heads);
    // A warp owns the query tiles mi0 = wsub and mi1 = wsub + WPG (T <= 64: at most 4 tiles, so two per warp at WPG = 2) and
    // computes them TOGETHER: every K fragment (scores) and every V fragment (output) is fetched from shared memory once and
    // feeds the MMAs of both tiles.  ncu (r2t): the LSU data pipe was the busiest unit of this kernel (75 % of peak), 576 of
    // its 1146 wavefronts per (image, head) were ldmatrix reads, half of them the second tile re-reading the same K and V.
    const int mi0 = wsub, mi1 = wsub + WPG;
    const bool two = WPG < 4 && mi1 < m_tiles;      // m_tiles <= 4: with four warps per tile the second tile never exists
    if (mi0 < m_tiles) {
      float s[2][8][4];
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) { s[q][ni][0] = s[q][ni][1] = s[q][ni][2] = s[q][ni][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        uint32_t a0[4], a1[4];
        ldsm_x4(a0, sQ + (mi0 * 16 + (lane & 15)) * ATT_LD + ks * 16 + (lane >> 4) * 8);
        if (two) ldsm_x4(a1, sQ + (mi1 * 16 + (lane & 15)) * ATT_LD + ks * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int np = 0; np < 4; np++) {     // two key tiles per ldmatrix.x4
          if (np * 2 < KT) {
            uint32_t bb[4];
            ldsm_x4(bb, sK + ((np * 2 + (lane >> 4)) * 8 + (lane & 7)) * ATT_LD + ks * 16 + ((lane >> 3) & 1) * 8);
            mma_16816(s[0][np * 2], a0, bb);
            if (np * 2 + 1 < KT) mma_16816(s[0][np * 2 + 1], a0, bb + 2);
            if (two) {
              mma_16816(s[1][np * 2], a1, bb);
              if (np * 2 + 1 < KT) mma_16816(s[1][np * 2 + 1], a1, bb + 2);
            }
          }
        }
      }
      // softmax over the raw scores: exp((s - max) * scale) = exp2(s * c - max * c), c = scale * log2(e); only the last
      // valid tile can hold padding keys when KT == ceil(T / 8).  The normalised probabilities are packed to the fp16 A
      // fragments of the second MMA right away (16 registers per query tile instead of 28 floats).
      uint32_t pf[2][4][4];
      const float cexp = scale * 1.4426950408889634f;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q == 0 || two) {
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int ni = 0; ni < KT; ni++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
              if (KT == 8 || ni == KT - 1) {   // KT < 8 is chosen only when T > 8 (KT - 1); the generic KT = 8 checks every tile
                const int col = ni * 8 + 2 * t + e;
                if (col >= T) { s[q][ni][e] = -INFINITY; s[q][ni][2 + e] = -INFINITY; }
              }
              mx0 = fmaxf(mx0, s[q][ni][e]); mx1 = fmaxf(mx1, s[q][ni][2 + e]);
            }
          }
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
          const float nm0 = -mx0 * cexp, nm1 = -mx1 * cexp;
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int ni = 0; ni < KT; ni++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
              float p0 = ex2_approx(fmaf(s[q][ni][e], cexp, nm0)), p1 = ex2_approx(fmaf(s[q][ni][2 + e], cexp, nm1));
              s[q][ni][e] = p0; s[q][ni][2 + e] = p1;
              sum0 += p0; sum1 += p1;
            }
          }
          sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
          sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
          const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
          for (int kk = 0; kk < (KT + 1) / 2; kk++) {
            __half2 h0 = __floats2half2_rn(s[q][2 * kk][0] * inv0, s[q][2 * kk][1] * inv0);
            __half2 h1 = __floats2half2_rn(s[q][2 * kk][2] * inv1, s[q][2 * kk][3] * inv1);
            pf[q][kk][0] = *reinterpret_cast<uint32_t*>(&h0); pf[q][kk][1] = *reinterpret_cast<uint32_t*>(&h1);
            if (2 * kk + 1 < KT) {
              __half2 h2 = __floats2half2_rn(s[q][2 * kk + 1][0] * inv0, s[q][2 * kk + 1][1] * inv0);
              __half2 h3 = __floats2half2_rn(s[q][2 * kk + 1][2] * inv1, s[q][2 * kk + 1][3] * inv1);
              pf[q][kk][2] = *reinterpret_cast<uint32_t*>(&h2); pf[q][kk][3] = *reinterpret_cast<uint32_t*>(&h3);
            } else {
              pf[q][kk][2] = 0u; pf[q][kk][3] = 0u;     // the all-padding tile: P = 0
            }
          }
        }
      }
      float oacc[2][8][4];
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) { oacc[q][ni][0] = oacc[q][ni][1] = oacc[q][ni][2] = oacc[q][ni][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < (KT + 1) / 2; kk++) {
#pragma unroll
        for (int np = 0; np < 4; np++) {     // two d-column tiles per ldmatrix.x4.trans
          uint32_t bb[4];
          ldsm_x4_t(bb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LD + (np * 2 + (lane >> 4)) * 8);
          mma_16816(oacc[0][np * 2], pf[0][kk], bb);
          mma_16816(oacc[0][np * 2 + 1], pf[0][kk], bb + 2);
          if (two) {
            mma_16816(oacc[1][np * 2], pf[1][kk], bb);
            mma_16816(oacc[1][np * 2 + 1], pf[1][kk], bb + 2);
          }
        }
      }
      // The warp's own Q rows are dead now (nobody else reads them): the output tile goes there as fp16 and leaves as
      // whole 128-byte rows (16 B per lane) instead of 16 half-filled 32-byte sectors per store instruction.
      __half* sO = const_cast<__half*>(sQ);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q == 0 || two) {
          const int mi = q ? mi1 : mi0;
#pragma unroll
          for (int ni = 0; ni < 8; ni++) {
            *reinterpret_cast<__half2*>(sO + (mi * 16 + g) * ATT_LD + ni * 8 + 2 * t) = __floats2half2_rn(oacc[q][ni][0], oacc[q][ni][1]);
            *reinterpret_cast<__half2*>(sO + (mi * 16 + g + 8) * ATT_LD + ni * 8 + 2 * t) = __floats2half2_rn(oacc[q][ni][2], oacc[q][ni][3]);
          }
        }
      }
      __syncwarp();
      __half* obase = o + ((long long)b * T) * W + h * 64;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q == 0 || two) {
          const int mi = q ? mi1 : mi0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int r = mi * 16 + j * 4 + (lane >> 3), ch = lane & 7;
            if (r < T) *reinterpret_cast<uint4*>(obase + (long long)r * W + ch * 8) = *reinterpret_cast<const uint4*>(sO + r * ATT_LD + ch * 8);
          }
        }
      }
    }
    named_bar_sync(1 + grp, GT);       // both warps are done with buf[cur] before it is refilled
    if (NBUF == 2) cur ^= 1;
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// attention for any T (ViT-L/14 = the reference's default `models.clip.type`: 257 tokens, graph.py:98-104).
// One CTA per (image, head): K and V of the head are staged once in shared memory (cp.async, rows padded to a
// multiple of 16 and zero filled); each warp owns 16-row query tiles (staged through a private smem slab for
// ldmatrix) and walks the keys in blocks of 64 with an online softmax (running max / sum in fp32), so the
// 16 x T score row never exists.  Same mma.sync m16n8k16 fragments as the T <= 64 kernels above.
// ------------------------------------------------------------------------------------------
constexpr int ATTF_WARPS = 6;
static inline int attf_smem_bytes(int Tp, int HD = 64) { return (2 * Tp + ATTF_WARPS * 16) * (HD + 8) * 2; }

// HD = head dim: 64 (ViT-B/32, B/16, L/14) or 80 (ViT-H/14, graph.py:105-111)
template <int HD>
__global__ void __launch_bounds__(ATTF_WARPS * 32) k_attention_flash(const __half* __restrict__ qkv, __half* __restrict__ o, int B, int T, int heads,
                                                                    int W, float scale, int Tp, int q_tiles) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  constexpr int LD = HD + 8, CPR = HD / 8;   // padded row stride (halfs), 16-byte chunks per row
  __half* sK = reinterpret_cast<__half*>(att_smem);
  __half* sV = sK + (size_t)Tp * LD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* sQ = sV + (size_t)Tp * LD + (size_t)warp * 16 * LD;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * W;
  const __half* src0 = qkv + ((long long)b * T) * ld + h * HD;
  for (int i = threadIdx.x; i < Tp * CPR; i += ATTF_WARPS * 32) {
    const int r = i / CPR, ch = i - r * CPR;
    __half* dk = sK + r * LD + ch * 8;
    __half* dv = sV + r * LD + ch * 8;
    if (r < T) {
      const __half* src = src0 + (long long)r * ld + ch * 8;
      cp_async16(dk, src + W);
      cp_async16(dv, src + 2 * W);
    } else {   // padding keys: finite zeros (their scores are masked, P = 0 there)
      *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int m_tiles = min((T + 15) / 16, q_tiles);
  for (int mi = warp; mi < m_tiles; mi += ATTF_WARPS) {
    // stage this warp's 16 query rows
    __syncwarp();
    for (int i = lane; i < 16 * CPR; i += 32) {
      const int r = i / CPR, ch = i - r * CPR, row = mi * 16 + r;
      __half* dq = sQ + r * LD + ch * 8;
      if (row < T) cp_async16(dq, src0 + (long long)row * ld + ch * 8);
      else *reinterpret_cast<uint4*>(dq) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    uint32_t aq[HD / 16][4];
#pragma unroll
    for (int ks = 0; ks < HD / 16; ks++) ldsm_x4(aq[ks], sQ + (lane & 15) * LD + ks * 16 + (lane >> 4) * 8);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float oacc[HD / 8][4];
#pragma unroll
    for (int ni = 0; ni < HD / 8; ni++) { oacc[ni][0] = oacc[ni][1] = oacc[ni][2] = oacc[ni][3] = 0.f; }
    for (int kb = 0; kb < Tp; kb += 64) {
      const int nt = min(4, (Tp - kb) >> 4);   // 16-key steps in this block (warp uniform)
      float s[8][4];
#pragma unroll
      for (int ni = 0; ni < 8; ni++) { s[ni][0] = s[ni][1] = s[ni][2] = s[ni][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < HD / 16; ks++) {
#pragma unroll
        for (int np = 0; np < 4; np++) {
          if (np < nt) {
            uint32_t bb[4];
            ldsm_x4(bb, sK + (kb + (np * 2 + (lane >> 4)) * 8 + (lane & 7)) * LD + ks * 16 + ((lane >> 3) & 1) * 8);
            mma_16816(s[np * 2], aq[ks], bb);
            mma_16816(s[np * 2 + 1], aq[ks], bb + 2);
          }
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int col = kb + ni * 8 + 2 * t + e;
          const float v0 = (col < T) ? s[ni][e] * scale : -INFINITY;
          const float v1 = (col < T) ? s[ni][2 + e] * scale : -INFINITY;
          s[ni][e] = v0; s[ni][2 + e] = v1;
          bm0 = fmaxf(bm0, v0); bm1 = fmaxf(bm1, v1);
        }
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      // every block holds at least one valid key (kb < T), so the new maxima are finite
      const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
      const float al0 = __expf(m0 - n0), al1 = __expf(m1 - n1);   // exp(-inf) = 0 on the first block
      m0 = n0; m1 = n1;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int ni = 0; ni < 8; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const float p0 = __expf(s[ni][e] - n0), p1 = __expf(s[ni][2 + e] - n1);
          s[ni][e] = p0; s[ni][2 + e] = p1;
          ps0 += p0; ps1 += p1;
        }
      }
      l0 = l0 * al0 + ps0; l1 = l1 * al1 + ps1;   // per-thread partial row sums (the quad is reduced once at the end)
#pragma unroll
      for (int ni = 0; ni < HD / 8; ni++) { oacc[ni][0] *= al0; oacc[ni][1] *= al0; oacc[ni][2] *= al1; oacc[ni][3] *= al1; }
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        if (kk < nt) {
          uint32_t a[4];
          __half2 h0 = __floats2half2_rn(s[2 * kk][0], s[2 * kk][1]);
          __half2 h1 = __floats2half2_rn(s[2 * kk][2], s[2 * kk][3]);
          __half2 h2 = __floats2half2_rn(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          __half2 h3 = __floats2half2_rn(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          a[0] = *reinterpret_cast<uint32_t*>(&h0); a[1] = *reinterpret_cast<uint32_t*>(&h1);
          a[2] = *reinterpret_cast<uint32_t*>(&h2); a[3] = *reinterpret_cast<uint32_t*>(&h3);
#pragma unroll
          for (int np = 0; np < HD / 16; np++) {
            uint32_t bb[4];
            ldsm_x4_t(bb, sV + (kb + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + (np * 2 + (lane >> 4)) * 8);
            mma_16816(oacc[np * 2], a, bb);
            mma_16816(oacc[np * 2 + 1], a, bb + 2);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = mi * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int ni = 0; ni < HD / 8; ni++) {
      const int col = h * HD + ni * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<__half2*>(o + ((long long)b * T + r0) * W + col) = __floats2half2_rn(oacc[ni][0] * inv0, oacc[ni][1] * inv0);
      if (r1 < T) *reinterpret_cast<__half2*>(o + ((long long)b * T + r1) * W + col) = __floats2half2_rn(oacc[ni][2] * inv1, oacc[ni][3] * inv1);
    }
  }
}

// straightforward fp32 reference attention (debug: HMSG_ATTN_SIMPLE=1), one block per (image, head)
__global__ void __launch_bounds__(128) k_attention_simple(const __half* __restrict__ qkv, __half* __restrict__ o, int B, int T, int heads, int W,
                                                          float scale) {
  __shared__ __half sQ[64][66], sK[64][66], sV[64][66];
  __shared__ float sP[4][64];
  int b = blockIdx.x / heads, h = blockIdx.x % heads;
  int ld = 3 * W;
  for (int i = threadIdx.x; i < T * 64; i += blockDim.x) {
    int r = i / 64, c = i % 64;
    const __half* base = qkv + ((long long)b * T + r) * ld + h * 64 + c;
    sQ[r][c] = base[0]; sK[r][c] = base[W]; sV[r][c] = base[2 * W];
  }
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < T; i += 4) {
    float sc[2]; float mx = -INFINITY;
    for (int jj = 0; jj < 2; jj++) {
      int j = lane + 32 * jj; float a = -INFINITY;
      if (j < T) { a = 0.f; for (int k = 0; k < 64; k++) a += __half2float(sQ[i][k]) * __half2float(sK[j][k]); a *= scale; }
      sc[jj] = a; mx = fmaxf(mx, a);
    }
    for (int of = 16; of > 0; of >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, of));
    float sum = 0.f;
    for (int jj = 0; jj < 2; jj++) { int j = lane + 32 * jj; float p = (j < T) ? expf(sc[jj] - mx) : 0.f; sc[jj] = p; sum += p; }
    for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
    for (int jj = 0; jj < 2; jj++) sP[warp][lane + 32 * jj] = sc[jj] / sum;
    __syncwarp();
    for (int dd = 0; dd < 2; dd++) {
      int dcol = lane + 32 * dd; float a = 0.f;
      for (int j = 0; j < T; j++) a += sP[warp][j] * __half2float(sV[j][dcol]);
      o[((long long)b * T + i) * W + h * 64 + dcol] = __float2half_rn(a);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// host state
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

struct LayerW {
  float *ln1_g, *ln1_b, *bqkv, *bo, *ln2_g, *ln2_b, *bfc, *bproj;
  __half *wqkv, *wo, *wfc, *wproj;
  // ln_1 folded into in_proj, ln_2 folded into c_fc (k_fold_ln)
  __half *wqkv_f, *wfc_f;
  float *sres_qkv, *bqkv_f, *sres_fc, *bfc_f;
};

struct VitState {
  hmsg_vit_desc desc{};
  int T = 0, G = 0, Kc = 0, Kpad = 0;
  PFN_encodeTiled encode = nullptr;
  __half* wconv = nullptr;
  float *cls = nullptr, *pos = nullptr, *lnpre_g = nullptr, *lnpre_b = nullptr, *lnpost_g = nullptr, *lnpost_b = nullptr;
  __half* wout = nullptr;   // [out_dim, width]
  std::vector<LayerW> layers;
  std::vector<void*> allocs;
  // workspaces for `cap` images
  int cap = 0;
  float* x = nullptr;       // [cap*T, W] fp32 residual (ln_fold = 0)
  __half* xh = nullptr;     // [cap*T, W] fp16 residual (ln_fold = 1)
  float2* statsA = nullptr; // [cap*T][W/64] partial (sum, sumsq) of the rows of xh for ln_1
  float2* statsB = nullptr; // [cap*T] (mean, rstd) of the rows of xh for the next folded LayerNorm (k_rowstats)
  bool ln_fold = true;      // LayerNorm folded into the consuming GEMM, fp16 residual stream
  __half* h = nullptr;      // [cap*T, W]
  __half* qkv = nullptr;    // [cap*T, 3W]   (aliases: patch-embed output fp32 [cap*(T-1), W])
  __half* gbuf = nullptr;   // [cap*T, mlp]  (aliases: im2col A0 fp16 [cap*(T-1), Kpad])
  __half* pooled = nullptr; // [cap, W]
  float* proj_out = nullptr;// [cap, out_dim]
  float* in_stage = nullptr; size_t in_stage_bytes = 0;
  float* out_stage = nullptr; size_t out_stage_bytes = 0;
  bool attn_simple = false;
  bool attn_v1 = false;
  bool attn_v2 = false;
  bool attn_v3_db = false;   // v3 with 4 double-buffered tiles per SM instead of 8 single-buffered ones
  int attn_v3_wide = 0;      // 5 / 6: that many tiles per SM with FOUR warps each (one query tile per warp)
  bool attn_flash = false;   // force the any-T kernel (always used when T > 64)
  bool last_cls_only = true; // last layer: only the class-token row feeds ln_post/proj, so only that row is computed past K/V
  int flash_smem_set = 0;
  bool smem_attr_set = false;
};

static int32_t make_tmap(hmsg_ctx* ctx, VitState* vs, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         int elem_bytes = 2, uint64_t ld_elems = 0) {
  // 2-D row-major tensor [rows, cols] (leading dimension ld_elems), box = (128 bytes of columns) x box_rows, 128B swizzle
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {(ld_elems ? ld_elems : cols) * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = vs->encode(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ctx->fail(HMSG_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return HMSG_OK;
}

static int32_t get_encoder(hmsg_ctx* ctx, PFN_encodeTiled* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) return ctx->fail(HMSG_ERR_CUDA, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
  *out = (PFN_encodeTiled)fn;
  return HMSG_OK;
}

template <int EPI>
static int32_t launch_gemm_t(hmsg_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& g) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_f16<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return ctx->fail(HMSG_ERR_CUDA, std::string("gemm smem attr: ") + cudaGetErrorString(e));
    attr = true;
  }
  int tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
  int grid = std::min(tiles, ctx->sm_count);
  ctx->prof_begin(PROF_GEMM);
  k_gemm_f16<EPI><<<grid, GEMM_THREADS, GEMM_SMEM, ctx->stream>>>(ta, tb, to, g);
  ctx->prof_end(PROF_GEMM, 2.0 * g.M * (double)g.N * g.K);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

template <int EPI>
static int32_t launch_gemm2_t(hmsg_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const GemmArgs& g) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_f16_2sm<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM);
    if (e != cudaSuccess) return ctx->fail(HMSG_ERR_CUDA, std::string("gemm2 smem attr: ") + cudaGetErrorString(e));
    attr = true;
  }
  int tiles = ((g.M + 255) / 256) * (g.N / BN);
  int grid = 2 * std::min(tiles, ctx->sm_count / 2);
  ctx->prof_begin(PROF_GEMM);
  k_gemm_f16_2sm<EPI><<<grid, GEMM_THREADS, GEMM2_SMEM, ctx->stream>>>(ta, tb, to, g);
  ctx->prof_end(PROF_GEMM, 2.0 * g.M * (double)g.N * g.K);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

static int g_gemm_2sm = -1;
int32_t vit_set_option(hmsg_ctx* ctx, const char* key, int value) {
  if (!strcmp(key, "gemm_2sm")) { g_gemm_2sm = value; return HMSG_OK; }
  if (!strcmp(key, "ln_fold")) {
    if (!ctx->vit) return ctx->fail(HMSG_ERR_STATE, "hmsg_set_option(ln_fold): load the encoder first");
    ctx->vit->ln_fold = value != 0;
    return HMSG_OK;
  }
  if (!strcmp(key, "last_layer_cls_only")) {
    if (!ctx->vit) return ctx->fail(HMSG_ERR_STATE, "hmsg_set_option(last_layer_cls_only): load the encoder first");
    ctx->vit->last_cls_only = value != 0;
    return HMSG_OK;
  }
  if (!strcmp(key, "attn_variant")) {   // 0: v3 (2 warps per tile, cp.async + ldmatrix), 1: v1, 2: fp32 reference kernel, 3: v2, 4: v3 double buffered, 5: any-T online-softmax kernel
    if (!ctx->vit) return ctx->fail(HMSG_ERR_STATE, "hmsg_set_option(attn_variant): load the encoder first");
    ctx->vit->attn_v1 = value == 1; ctx->vit->attn_simple = value == 2; ctx->vit->attn_v2 = value == 3; ctx->vit->attn_v3_db = value == 4; ctx->vit->attn_flash = value == 5; ctx->vit->attn_v3_wide = value == 6 ? 5 : (value == 7 ? 6 : 0); ctx->vit->smem_attr_set = false;
    return HMSG_OK;
  }
  return -1;
}

// extra operands of the folded-LayerNorm / fp16-residual epilogues
struct GemmExtra {
  const float* sres = nullptr;
  const float2* stats_in = nullptr;
  float2* stats_out = nullptr;
  const __half* resid = nullptr;
  long long ldr = 0;
  int stat_stride = 1;
};

// C = A[M,K] * Wt[N,K]^T with epilogue
static int32_t gemm(hmsg_ctx* ctx, VitState* vs, int epi, const __half* A, const __half* Wt, int M, int N, int K, const float* bias, void* out,
                    int ldo, long long lda = 0, const GemmExtra* ex = nullptr) {
  if (N % BN != 0 || K % BK != 0 || M <= 0) return ctx->fail(HMSG_ERR_ARG, "gemm: N must be a multiple of 256 and K of 64");
  CUtensorMap ta, tb, to;
  int32_t rc;
  if (g_gemm_2sm < 0) { const char* e = getenv("HMSG_GEMM_2SM"); g_gemm_2sm = e ? atoi(e) : 1; }   // default: 2-CTA pairs
  const bool two_sm = g_gemm_2sm != 0 && M > 128;
  if ((rc = make_tmap(ctx, vs, &ta, A, (uint64_t)M, (uint64_t)K, BM, 2, (uint64_t)lda))) return rc;
  if ((rc = make_tmap(ctx, vs, &tb, Wt, (uint64_t)N, (uint64_t)K, two_sm ? 128 : BN))) return rc;
  const bool f16out = (epi != EPI_F32_RESIDUAL && epi != EPI_F32_STORE);
  if ((rc = make_tmap(ctx, vs, &to, out, (uint64_t)M, (uint64_t)N, BM, f16out ? 2 : 4, (uint64_t)ldo))) return rc;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.quick_gelu = vs->desc.quick_gelu;
  g.stat_stride = 1; g.inv_w = 1.0f / (float)K; g.ln_eps = 1e-5f; g.stat_parts = vs->desc.width / 64;
  if (ex) { g.sres = ex->sres; g.stats_in = ex->stats_in; g.stats_out = ex->stats_out; g.resid = ex->resid; g.ldr = ex->ldr; g.stat_stride = ex->stat_stride; }
  if ((epi == EPI_F16_LN || epi == EPI_F16_LN_GELU) && (!g.sres || !g.stats_in || !bias)) return ctx->fail(HMSG_ERR_ARG, "gemm: folded-LN epilogue without operands");
  if (epi == EPI_F16_RESID_STATS && !g.resid) return ctx->fail(HMSG_ERR_ARG, "gemm: residual epilogue without x");
  if (two_sm) {
    switch (epi) {
      case EPI_F16_BIAS: return launch_gemm2_t<EPI_F16_BIAS>(ctx, ta, tb, to, g);
      case EPI_F16_BIAS_GELU: return launch_gemm2_t<EPI_F16_BIAS_GELU>(ctx, ta, tb, to, g);
      case EPI_F32_RESIDUAL: return launch_gemm2_t<EPI_F32_RESIDUAL>(ctx, ta, tb, to, g);
      case EPI_F16_LN: return launch_gemm2_t<EPI_F16_LN>(ctx, ta, tb, to, g);
      case EPI_F16_LN_GELU: return launch_gemm2_t<EPI_F16_LN_GELU>(ctx, ta, tb, to, g);
      case EPI_F16_RESID_STATS: return launch_gemm2_t<EPI_F16_RESID_STATS>(ctx, ta, tb, to, g);
      default: return launch_gemm2_t<EPI_F32_STORE>(ctx, ta, tb, to, g);
    }
  }
  switch (epi) {
    case EPI_F16_BIAS: return launch_gemm_t<EPI_F16_BIAS>(ctx, ta, tb, to, g);
    case EPI_F16_BIAS_GELU: return launch_gemm_t<EPI_F16_BIAS_GELU>(ctx, ta, tb, to, g);
    case EPI_F32_RESIDUAL: return launch_gemm_t<EPI_F32_RESIDUAL>(ctx, ta, tb, to, g);
    case EPI_F16_LN: return launch_gemm_t<EPI_F16_LN>(ctx, ta, tb, to, g);
    case EPI_F16_LN_GELU: return launch_gemm_t<EPI_F16_LN_GELU>(ctx, ta, tb, to, g);
    case EPI_F16_RESID_STATS: return launch_gemm_t<EPI_F16_RESID_STATS>(ctx, ta, tb, to, g);
    default: return launch_gemm_t<EPI_F32_STORE>(ctx, ta, tb, to, g);
  }
}

int32_t vit_destroy(hmsg_ctx* ctx) {
  VitState* vs = ctx->vit;
  if (!vs) return HMSG_OK;
  for (void* p : vs->allocs) cudaFree(p);
  free_dev(vs->x); free_dev(vs->h); free_dev(vs->qkv); free_dev(vs->gbuf); free_dev(vs->pooled); free_dev(vs->proj_out);
  free_dev(vs->xh); free_dev(vs->statsA); free_dev(vs->statsB);
  free_dev(vs->in_stage); free_dev(vs->out_stage);
  delete vs;
  ctx->vit = nullptr;
  return HMSG_OK;
}

template <typename T>
static int32_t dalloc(hmsg_ctx* ctx, VitState* vs, T** p, size_t n) {
  HMSG_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  vs->allocs.push_back(*p);
  return HMSG_OK;
}

extern "C" int32_t hmsg_encoder_load(hmsg_ctx* ctx, const hmsg_vit_desc* desc, const float* blob, int64_t blob_floats) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!desc || !blob) return ctx->fail(HMSG_ERR_ARG, "hmsg_encoder_load: null argument");
  const hmsg_vit_desc& d = *desc;
  if (d.image <= 0 || d.patch <= 0 || d.layers <= 0 || d.heads <= 0 || d.width % 128 != 0 || d.width > 1536 || (d.width / d.heads != 64 && d.width / d.heads != 80) || d.width % d.heads != 0 || d.image % d.patch != 0 || d.mlp % 256 != 0 || d.out_dim % 256 != 0 ||
      (3 * d.width) % 256 != 0 || d.width % 256 != 0)
    return ctx->fail(HMSG_ERR_ARG, "hmsg_encoder_load: unsupported ViT shape (need head_dim 64 or 80, width/mlp/out_dim multiples of 256)");
  int G = d.image / d.patch, T = G * G + 1;
  if (T > 4096) return ctx->fail(HMSG_ERR_ARG, "hmsg_encoder_load: more than 4096 tokens per image is not supported");
  if (attf_smem_bytes((T + 15) / 16 * 16, d.width / d.heads) > 220 * 1024)
    return ctx->fail(HMSG_ERR_ARG, "hmsg_encoder_load: K/V of one head do not fit in shared memory (T too large)");
  vit_destroy(ctx);
  VitState* vs = new VitState();
  ctx->vit = vs;
  vs->desc = d; vs->G = G; vs->T = T; vs->Kc = 3 * d.patch * d.patch; vs->Kpad = (vs->Kc + 63) / 64 * 64;
  int32_t rc = get_encoder(ctx, &vs->encode);
  if (rc) return rc;
  const int W = d.width;
  int64_t expect = (int64_t)W * vs->Kc + W + (int64_t)T * W + 2 * W +
                   (int64_t)d.layers * (2 * W + 3LL * W * W + 3 * W + (int64_t)W * W + W + 2 * W + (int64_t)d.mlp * W + d.mlp + (int64_t)W * d.mlp + W) +
                   2 * W + (int64_t)W * d.out_dim;
  if (blob_floats != expect)
    return ctx->fail(HMSG_ERR_ARG, "hmsg_encoder_load: blob has " + std::to_string(blob_floats) + " floats, expected " + std::to_string(expect));
  // stage the whole blob on the device once, then convert / slice
  float* dblob = nullptr;
  HMSG_CUDA(cudaMalloc((void**)&dblob, (size_t)blob_floats * 4));
  vs->allocs.push_back(dblob);   // fp32 vectors are used in place
  HMSG_CUDA(cudaMemcpyAsync(dblob, blob, (size_t)blob_floats * 4, cudaMemcpyHostToDevice, ctx->stream));
  size_t off = 0;
  auto take = [&](size_t n) { float* p = dblob + off; off += n; return p; };
  auto to_f16 = [&](const float* src, __half** dst, size_t n) -> int32_t {
    int32_t r = dalloc(ctx, vs, dst, n);
    if (r) return r;
    k_f32_to_f16<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, *dst, (long long)n);
    HMSG_LAUNCH_CHECK();
    return HMSG_OK;
  };
  {
    float* cw = take((size_t)W * vs->Kc);
    if ((rc = dalloc(ctx, vs, &vs->wconv, (size_t)W * vs->Kpad))) return rc;
    long long n = (long long)W * vs->Kpad;
    k_pad_weight<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(cw, vs->wconv, W, vs->Kc, vs->Kpad);
    HMSG_LAUNCH_CHECK();
  }
  vs->cls = take(W); vs->pos = take((size_t)T * W); vs->lnpre_g = take(W); vs->lnpre_b = take(W);
  vs->layers.resize(d.layers);
  for (int l = 0; l < d.layers; l++) {
    LayerW& L = vs->layers[l];
    L.ln1_g = take(W); L.ln1_b = take(W);
    float* wqkv = take(3LL * W * W); L.bqkv = take(3 * W);
    float* wo = take((size_t)W * W); L.bo = take(W);
    L.ln2_g = take(W); L.ln2_b = take(W);
    float* wfc = take((size_t)d.mlp * W); L.bfc = take(d.mlp);
    float* wpj = take((size_t)W * d.mlp); L.bproj = take(W);
    if ((rc = to_f16(wqkv, &L.wqkv, 3LL * W * W))) return rc;
    if ((rc = to_f16(wo, &L.wo, (size_t)W * W))) return rc;
    if ((rc = to_f16(wfc, &L.wfc, (size_t)d.mlp * W))) return rc;
    if ((rc = to_f16(wpj, &L.wproj, (size_t)W * d.mlp))) return rc;
    if ((rc = dalloc(ctx, vs, &L.wqkv_f, (size_t)3 * W * W))) return rc;
    if ((rc = dalloc(ctx, vs, &L.wfc_f, (size_t)d.mlp * W))) return rc;
    if ((rc = dalloc(ctx, vs, &L.sres_qkv, (size_t)3 * W))) return rc;
    if ((rc = dalloc(ctx, vs, &L.bqkv_f, (size_t)3 * W))) return rc;
    if ((rc = dalloc(ctx, vs, &L.sres_fc, (size_t)d.mlp))) return rc;
    if ((rc = dalloc(ctx, vs, &L.bfc_f, (size_t)d.mlp))) return rc;
    k_fold_ln<<<(3 * W * 32 + 255) / 256, 256, 0, ctx->stream>>>(wqkv, L.ln1_g, L.ln1_b, L.bqkv, 3 * W, W, L.wqkv_f, L.sres_qkv, L.bqkv_f);
    HMSG_LAUNCH_CHECK();
    k_fold_ln<<<(d.mlp * 32 + 255) / 256, 256, 0, ctx->stream>>>(wfc, L.ln2_g, L.ln2_b, L.bfc, d.mlp, W, L.wfc_f, L.sres_fc, L.bfc_f);
    HMSG_LAUNCH_CHECK();
  }
  vs->lnpost_g = take(W); vs->lnpost_b = take(W);
  {
    float* pj = take((size_t)W * d.out_dim);   // [W, out] (x @ proj) -> [out, W]
    if ((rc = dalloc(ctx, vs, &vs->wout, (size_t)W * d.out_dim))) return rc;
    long long n = (long long)W * d.out_dim;
    k_f32_to_f16_T<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(pj, vs->wout, W, d.out_dim);
    HMSG_LAUNCH_CHECK();
  }
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (const char* e = getenv("HMSG_LN_FOLD")) vs->ln_fold = atoi(e) != 0;
  if (const char* e = getenv("HMSG_ATTN_SIMPLE")) vs->attn_simple = atoi(e) != 0;
  if (const char* e = getenv("HMSG_ATTN_V1")) vs->attn_v1 = atoi(e) != 0;
  if (const char* e = getenv("HMSG_ATTN_WIDE")) { int v = atoi(e); vs->attn_v3_wide = (v == 5 || v == 6) ? v : 0; }
  return HMSG_OK;
}

static int32_t ensure_ws(hmsg_ctx* ctx, VitState* vs, int B) {
  if (B <= vs->cap) return HMSG_OK;
  free_dev(vs->x); free_dev(vs->h); free_dev(vs->qkv); free_dev(vs->gbuf); free_dev(vs->pooled); free_dev(vs->proj_out);
  free_dev(vs->xh); free_dev(vs->statsA); free_dev(vs->statsB);
  vs->cap = 0;
  const hmsg_vit_desc& d = vs->desc;
  size_t R = (size_t)B * vs->T, W = d.width;
  size_t qkv_bytes = std::max(R * 3 * W * 2, (size_t)B * (vs->T - 1) * W * 4);
  size_t g_bytes = std::max(R * (size_t)d.mlp * 2, (size_t)B * (vs->T - 1) * vs->Kpad * 2);
  HMSG_CUDA(cudaMalloc((void**)&vs->x, R * W * 4));
  HMSG_CUDA(cudaMalloc((void**)&vs->xh, R * W * 2));
  HMSG_CUDA(cudaMalloc((void**)&vs->statsA, R * 8 * (W / 64)));
  HMSG_CUDA(cudaMalloc((void**)&vs->statsB, R * 8));
  HMSG_CUDA(cudaMalloc((void**)&vs->h, R * W * 2));
  HMSG_CUDA(cudaMalloc((void**)&vs->qkv, qkv_bytes));
  HMSG_CUDA(cudaMalloc((void**)&vs->gbuf, g_bytes));
  HMSG_CUDA(cudaMalloc((void**)&vs->pooled, (size_t)B * W * 2));
  HMSG_CUDA(cudaMalloc((void**)&vs->proj_out, (size_t)B * d.out_dim * 4));
  vs->cap = B;
  return HMSG_OK;
}

template <int NV>
static int32_t forward_chunk_t(hmsg_ctx* ctx, VitState* vs, const float* dx, int B, float* dout, int normalize, bool patches_ready = false) {
  const hmsg_vit_desc& d = vs->desc;
  const int W = d.width, T = vs->T, G = vs->G;
  const long long R = (long long)B * T, RP = (long long)B * (T - 1);
  int32_t rc;
  __half* a0 = vs->gbuf;
  float* patch = reinterpret_cast<float*>(vs->qkv);
  const bool fast_im2col = d.image % 4 == 0 && d.patch % 4 == 0;
  if (!patches_ready && !fast_im2col) {
    ctx->prof_begin(PROF_ELTWISE);
    long long n = RP * vs->Kpad;
    k_im2col_generic<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dx, a0, n, d.image, d.patch, G, vs->Kc, vs->Kpad);
    ctx->prof_end(PROF_ELTWISE, (double)B * 3 * d.image * d.image * 6);
    HMSG_LAUNCH_CHECK();
  }
  if (!patches_ready && fast_im2col && vs->Kpad != vs->Kc) {
    long long n = RP * (vs->Kpad - vs->Kc);
    k_zero_pad_cols<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a0, RP, vs->Kc, vs->Kpad);
    HMSG_LAUNCH_CHECK();
  }
  if (!patches_ready && fast_im2col) {
    ctx->prof_begin(PROF_ELTWISE);
    long long nrows = (long long)B * 3 * d.image;
    k_im2col<<<(unsigned)((nrows * 32 + 255) / 256), 256, 0, ctx->stream>>>(dx, a0, nrows, d.image, d.patch, G, vs->Kpad);
    ctx->prof_end(PROF_ELTWISE, (double)B * 3 * d.image * d.image * 6);
    HMSG_LAUNCH_CHECK();
  }
  if ((rc = gemm(ctx, vs, EPI_F32_STORE, a0, vs->wconv, (int)RP, W, vs->Kpad, nullptr, patch, W))) return rc;
  const int HD = W / d.heads;
  const float scale = 1.0f / sqrtf((float)HD);
  const bool cls_last = vs->last_cls_only && T > 1 && !vs->attn_simple && !vs->attn_v1 && !vs->attn_v2;
  // attention launcher shared by both residual-stream forms: qkv [R,3W] -> h [R,W]
  auto attention = [&](int q_tiles) -> int32_t {
    ctx->prof_begin(PROF_ATTN);
    if (T > 64 || vs->attn_flash || HD != 64) {
      const int Tp = (T + 15) / 16 * 16;
      const int sm = attf_smem_bytes(Tp, HD);
      if (vs->flash_smem_set != sm) {
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_flash<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_flash<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        vs->flash_smem_set = sm;
      }
      if (HD == 80) k_attention_flash<80><<<(unsigned)(B * d.heads), ATTF_WARPS * 32, sm, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, Tp, q_tiles);
      else k_attention_flash<64><<<(unsigned)(B * d.heads), ATTF_WARPS * 32, sm, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, Tp, q_tiles);
    } else if (vs->attn_simple) {
      k_attention_simple<<<(unsigned)(B * d.heads), 128, 0, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale);
    } else if (vs->attn_v1) {
      size_t sm = (size_t)4 * 3 * 64 * ATT_LD * 2;
      if (!vs->smem_attr_set) {
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        vs->smem_attr_set = true;
      }
      long long pairs = (long long)B * d.heads;
      k_attention_mma<<<(unsigned)((pairs + 3) / 4), 128, sm, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale);
    } else if (vs->attn_v2) {
      if (!vs->smem_attr_set) {
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma2, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
        vs->smem_attr_set = true;
      }
      long long pairs = (long long)B * d.heads;
      int grid = (int)std::min<long long>((pairs + ATT2_WARPS - 1) / ATT2_WARPS, ctx->sm_count);
      k_attention_mma2<<<grid, ATT2_WARPS * 32, ATT2_SMEM, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale);
    } else {
      constexpr int SM3 = 4 * 2 * 3 * ATT2_TILE * 2;   // == 8 * 1 * 3 * ATT2_TILE * 2
      if (!vs->smem_attr_set) {
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma3<4, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM3));
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma3<8, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM3));
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma3<8, 1, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM3));
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma3<5, 1, 7, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM3));
        HMSG_CUDA(cudaFuncSetAttribute(k_attention_mma3<6, 1, 7, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM3));
        vs->smem_attr_set = true;
      }
      long long pairs = (long long)B * d.heads;
      if (pairs >= (1ll << 31)) return ctx->fail(HMSG_ERR_ARG, "attention: B * heads must be below 2^31");
      if (vs->attn_v3_wide && T > 48 && T <= 56) {
        const int G = vs->attn_v3_wide;
        int grid = (int)std::min<long long>((pairs + G - 1) / G, ctx->sm_count);
        if (G == 5) k_attention_mma3<5, 1, 7, 4><<<grid, 5 * 128, 5 * 3 * ATT2_TILE * 2, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, q_tiles);
        else        k_attention_mma3<6, 1, 7, 4><<<grid, 6 * 128, 6 * 3 * ATT2_TILE * 2, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, q_tiles);
      } else if (vs->attn_v3_db) {
        int grid = (int)std::min<long long>((pairs + 3) / 4, ctx->sm_count);
        k_attention_mma3<4, 2, 8><<<grid, 4 * 64, SM3, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, q_tiles);
      } else {
        int grid = (int)std::min<long long>((pairs + 7) / 8, ctx->sm_count);
        if (T > 48 && T <= 56)   // ViT-B/32: 50 tokens = 7 key tiles
          k_attention_mma3<8, 1, 7><<<grid, 8 * 64, SM3, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, q_tiles);
        else
          k_attention_mma3<8, 1, 8><<<grid, 8 * 64, SM3, ctx->stream>>>(vs->qkv, vs->h, B, T, d.heads, W, scale, q_tiles);
      }
    }
    ctx->prof_end(PROF_ATTN, 4.0 * B * d.heads * (double)T * T * HD);
    HMSG_LAUNCH_CHECK();
    return HMSG_OK;
  };
  if (vs->ln_fold) {
    // ---- fp16 residual stream, LayerNorms folded into the consuming GEMMs: per block 4 GEMMs + attention, no LN kernels
    // vs->statsA: per-tile (sum, sumsq) parts written by the residual epilogues; vs->statsB: (mean, rstd) per row for the next folded LN
    float2* parts = vs->statsA;
    float2* ms = vs->statsB;
    const int P = W / 64;
    auto rowstats = [&](long long rows, long long stride) -> int32_t {
      ctx->prof_begin(PROF_ELTWISE);
      k_rowstats<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(parts, P, rows, stride, 1.0f / (float)W, 1e-5f, ms);
      ctx->prof_end(PROF_ELTWISE, (double)rows * P * 8);
      HMSG_LAUNCH_CHECK();
      return HMSG_OK;
    };
    k_embed_lnpre_f16<NV><<<(unsigned)((R * 32 + 255) / 256), 256, 0, ctx->stream>>>(patch, vs->cls, vs->pos, vs->lnpre_g, vs->lnpre_b, vs->xh, ms, R, T);
    HMSG_LAUNCH_CHECK();
    for (int l = 0; l < d.layers; l++) {
      const LayerW& L = vs->layers[l];
      const bool cls_only = cls_last && l == d.layers - 1;
      GemmExtra ex;
      if (!cls_only) {
        ex = GemmExtra(); ex.sres = L.sres_qkv; ex.stats_in = ms;
        if ((rc = gemm(ctx, vs, EPI_F16_LN, vs->xh, L.wqkv_f, (int)R, 3 * W, W, L.bqkv_f, vs->qkv, 3 * W, 0, &ex))) return rc;
      } else {   // last block: K, V for every token, Q for the class-token rows only (strided views)
        ex = GemmExtra(); ex.sres = L.sres_qkv + W; ex.stats_in = ms;
        if ((rc = gemm(ctx, vs, EPI_F16_LN, vs->xh, L.wqkv_f + (size_t)W * W, (int)R, 2 * W, W, L.bqkv_f + W, vs->qkv + W, 3 * W, 0, &ex))) return rc;
        ex = GemmExtra(); ex.sres = L.sres_qkv; ex.stats_in = ms; ex.stat_stride = T;
        if ((rc = gemm(ctx, vs, EPI_F16_LN, vs->xh, L.wqkv_f, B, W, W, L.bqkv_f, vs->qkv, T * 3 * W, (long long)T * W, &ex))) return rc;
      }
      if ((rc = attention(cls_only ? 1 : (1 << 30)))) return rc;
      if (cls_only) {
        ex = GemmExtra(); ex.resid = vs->xh; ex.ldr = (long long)T * W; ex.stats_out = parts; ex.stat_stride = T;
        if ((rc = gemm(ctx, vs, EPI_F16_RESID_STATS, vs->h, L.wo, B, W, W, L.bo, vs->xh, T * W, (long long)T * W, &ex))) return rc;
        if ((rc = rowstats(B, T))) return rc;
        ex = GemmExtra(); ex.sres = L.sres_fc; ex.stats_in = ms; ex.stat_stride = T;
        if ((rc = gemm(ctx, vs, EPI_F16_LN_GELU, vs->xh, L.wfc_f, B, d.mlp, W, L.bfc_f, vs->gbuf, d.mlp, (long long)T * W, &ex))) return rc;
        ex = GemmExtra(); ex.resid = vs->xh; ex.ldr = (long long)T * W;
        if ((rc = gemm(ctx, vs, EPI_F16_RESID_STATS, vs->gbuf, L.wproj, B, W, d.mlp, L.bproj, vs->xh, T * W, 0, &ex))) return rc;
        continue;
      }
      ex = GemmExtra(); ex.resid = vs->xh; ex.ldr = W; ex.stats_out = parts;
      if ((rc = gemm(ctx, vs, EPI_F16_RESID_STATS, vs->h, L.wo, (int)R, W, W, L.bo, vs->xh, W, 0, &ex))) return rc;
      if ((rc = rowstats(R, 1))) return rc;
      ex = GemmExtra(); ex.sres = L.sres_fc; ex.stats_in = ms;
      if ((rc = gemm(ctx, vs, EPI_F16_LN_GELU, vs->xh, L.wfc_f, (int)R, d.mlp, W, L.bfc_f, vs->gbuf, d.mlp, 0, &ex))) return rc;
      ex = GemmExtra(); ex.resid = vs->xh; ex.ldr = W; ex.stats_out = parts;
      if ((rc = gemm(ctx, vs, EPI_F16_RESID_STATS, vs->gbuf, L.wproj, (int)R, W, d.mlp, L.bproj, vs->xh, W, 0, &ex))) return rc;
      if ((rc = rowstats(R, 1))) return rc;
    }
    k_layernorm_h2h<NV><<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->xh, vs->lnpost_g, vs->lnpost_b, vs->pooled, B, T);
    HMSG_LAUNCH_CHECK();
    if ((rc = gemm(ctx, vs, EPI_F32_STORE, vs->pooled, vs->wout, B, d.out_dim, W, nullptr, vs->proj_out, d.out_dim))) return rc;
    k_l2norm_rows<<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->proj_out, dout, B, d.out_dim, normalize);
    HMSG_LAUNCH_CHECK();
    return HMSG_OK;
  }
  k_embed_lnpre<NV><<<(unsigned)((R * 32 + 255) / 256), 256, 0, ctx->stream>>>(patch, vs->cls, vs->pos, vs->lnpre_g, vs->lnpre_b, vs->x, R, T);
  HMSG_LAUNCH_CHECK();
  for (int l = 0; l < d.layers; l++) {
    const LayerW& L = vs->layers[l];
    ctx->prof_begin(PROF_ELTWISE);
    k_layernorm_f16<NV><<<(unsigned)((R * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->x, L.ln1_g, L.ln1_b, vs->h, R, 1);
    ctx->prof_end(PROF_ELTWISE, (double)R * W * 6);
    HMSG_LAUNCH_CHECK();
    // The tower's output is ln_post(x[:, 0]) @ proj: in the last block only the class-token row of every image is
    // consumed.  K and V are still needed for all tokens; Q, attention, out-proj, ln_2 and the MLP run on that row
    // alone (strided TMA views pick row b*T of h / x in place).  Same values as the full computation.
    const bool cls_only = cls_last && l == d.layers - 1;
    const int q_tiles = cls_only ? 1 : (1 << 30);
    if (!cls_only) {
      if ((rc = gemm(ctx, vs, EPI_F16_BIAS, vs->h, L.wqkv, (int)R, 3 * W, W, L.bqkv, vs->qkv, 3 * W))) return rc;
    } else {
      if ((rc = gemm(ctx, vs, EPI_F16_BIAS, vs->h, L.wqkv + (size_t)W * W, (int)R, 2 * W, W, L.bqkv + W, vs->qkv + W, 3 * W))) return rc;
      if ((rc = gemm(ctx, vs, EPI_F16_BIAS, vs->h, L.wqkv, B, W, W, L.bqkv, vs->qkv, T * 3 * W, (long long)T * W))) return rc;
    }
    if ((rc = attention(q_tiles))) return rc;
    if (cls_only) {
      if ((rc = gemm(ctx, vs, EPI_F32_RESIDUAL, vs->h, L.wo, B, W, W, L.bo, vs->x, T * W, (long long)T * W))) return rc;
      ctx->prof_begin(PROF_ELTWISE);
      k_layernorm_f16<NV><<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->x, L.ln2_g, L.ln2_b, vs->pooled, B, T);
      ctx->prof_end(PROF_ELTWISE, (double)B * W * 6);
      HMSG_LAUNCH_CHECK();
      if ((rc = gemm(ctx, vs, EPI_F16_BIAS_GELU, vs->pooled, L.wfc, B, d.mlp, W, L.bfc, vs->gbuf, d.mlp))) return rc;
      if ((rc = gemm(ctx, vs, EPI_F32_RESIDUAL, vs->gbuf, L.wproj, B, W, d.mlp, L.bproj, vs->x, T * W))) return rc;
      continue;
    }
    if ((rc = gemm(ctx, vs, EPI_F32_RESIDUAL, vs->h, L.wo, (int)R, W, W, L.bo, vs->x, W))) return rc;
    ctx->prof_begin(PROF_ELTWISE);
    k_layernorm_f16<NV><<<(unsigned)((R * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->x, L.ln2_g, L.ln2_b, vs->h, R, 1);
    ctx->prof_end(PROF_ELTWISE, (double)R * W * 6);
    HMSG_LAUNCH_CHECK();
    if ((rc = gemm(ctx, vs, EPI_F16_BIAS_GELU, vs->h, L.wfc, (int)R, d.mlp, W, L.bfc, vs->gbuf, d.mlp))) return rc;
    if ((rc = gemm(ctx, vs, EPI_F32_RESIDUAL, vs->gbuf, L.wproj, (int)R, W, d.mlp, L.bproj, vs->x, W))) return rc;
  }
  // ln_post on the class token of every image, projection, L2 normalisation
  k_layernorm_f16<NV><<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->x, vs->lnpost_g, vs->lnpost_b, vs->pooled, B, T);
  HMSG_LAUNCH_CHECK();
  if ((rc = gemm(ctx, vs, EPI_F32_STORE, vs->pooled, vs->wout, B, d.out_dim, W, nullptr, vs->proj_out, d.out_dim))) return rc;
  k_l2norm_rows<<<(unsigned)(((long long)B * 32 + 255) / 256), 256, 0, ctx->stream>>>(vs->proj_out, dout, B, d.out_dim, normalize);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

static int32_t forward_chunk(hmsg_ctx* ctx, VitState* vs, const float* dx, int B, float* dout, int normalize, bool patches_ready = false) {
  switch (vs->desc.width / 128) {
    case 6: return forward_chunk_t<6>(ctx, vs, dx, B, dout, normalize, patches_ready);
    case 8: return forward_chunk_t<8>(ctx, vs, dx, B, dout, normalize, patches_ready);
    case 10: return forward_chunk_t<10>(ctx, vs, dx, B, dout, normalize, patches_ready);
    case 2: return forward_chunk_t<2>(ctx, vs, dx, B, dout, normalize, patches_ready);
    case 4: return forward_chunk_t<4>(ctx, vs, dx, B, dout, normalize, patches_ready);
  }
  return ctx->fail(HMSG_ERR_ARG, "encoder: unsupported width");
}

// device-pointer entry used by the ingest pipeline (no staging)
int32_t vit_encode_device(hmsg_ctx* ctx, const float* dx, int B, float* dout, int normalize) {
  VitState* vs = ctx->vit;
  if (!vs) return ctx->fail(HMSG_ERR_STATE, "encoder: call hmsg_encoder_load first");
  int chunk_max = 2080;
  if (const char* e = getenv("HMSG_ENC_CHUNK")) chunk_max = std::max(1, atoi(e));
  int32_t rc = ensure_ws(ctx, vs, std::min(B, chunk_max));
  if (rc) return rc;
  size_t img = (size_t)3 * vs->desc.image * vs->desc.image;
  for (int b0 = 0; b0 < B; b0 += chunk_max) {
    int nb = std::min(chunk_max, B - b0);
    if ((rc = forward_chunk(ctx, vs, dx + (size_t)b0 * img, nb, dout + (size_t)b0 * vs->desc.out_dim, normalize))) return rc;
  }
  return HMSG_OK;
}

int32_t crops_run(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const int32_t* xywh, int32_t bbox_margin, int32_t on_device,
                  float** crops_dev_out, __half* a0, int P, int G, int Kpad);

// fused ingest entry: crops of a frame batch are resampled straight into the encoder's fp16 patch
// matrix (no fp32 crop tensor, no im2col pass) and encoded.  feats_out [n*(2M+1), out_dim] device.
extern "C" int32_t hmsg_encode_crops(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, int32_t M, const int32_t* xywh, int32_t bbox_margin,
                                     int32_t on_device, float* feats_out) {
  if (!ctx) return HMSG_ERR_ARG;
  VitState* vs = ctx->vit;
  if (!vs) return ctx->fail(HMSG_ERR_STATE, "hmsg_encode_crops: call hmsg_encoder_load first");
  if (!feats_out || !xywh) return ctx->fail(HMSG_ERR_ARG, "hmsg_encode_crops: null argument");
  if (vs->desc.image != 224) return ctx->fail(HMSG_ERR_ARG, "hmsg_encode_crops: the crop kernels emit 224x224 inputs");
  const int B = n * (2 * M + 1);
  int32_t rc;
  if (B > 4160 || vs->Kpad != vs->Kc) {   // large batches / padded K: unfused path
    float* crops = nullptr;
    if ((rc = crops_run(ctx, frame_begin, n, M, xywh, bbox_margin, on_device, &crops, nullptr, 0, 0, 0))) return rc;
    return vit_encode_device(ctx, crops, B, feats_out, 1);
  }
  if ((rc = ensure_ws(ctx, vs, B))) return rc;
  if ((rc = crops_run(ctx, frame_begin, n, M, xywh, bbox_margin, on_device, nullptr, vs->gbuf, vs->desc.patch, vs->G, vs->Kpad))) return rc;
  return forward_chunk(ctx, vs, nullptr, B, feats_out, 1, true);
}

extern "C" int32_t hmsg_encode_images(hmsg_ctx* ctx, const float* nchw, int32_t B, float* out, int32_t normalize, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  VitState* vs = ctx->vit;
  if (!vs) return ctx->fail(HMSG_ERR_STATE, "hmsg_encode_images: call hmsg_encoder_load first");
  if (!nchw || !out || B <= 0) return ctx->fail(HMSG_ERR_ARG, "hmsg_encode_images: bad argument");
  if (on_device) return vit_encode_device(ctx, nchw, B, out, normalize);
  size_t img = (size_t)3 * vs->desc.image * vs->desc.image;
  int32_t rc;
  if ((rc = ctx->reserve(&vs->in_stage, &vs->in_stage_bytes, (size_t)B * img * 4))) return rc;
  if ((rc = ctx->reserve(&vs->out_stage, &vs->out_stage_bytes, (size_t)B * vs->desc.out_dim * 4))) return rc;
  HMSG_CUDA(cudaMemcpyAsync(vs->in_stage, nchw, (size_t)B * img * 4, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = vit_encode_device(ctx, vs->in_stage, B, vs->out_stage, normalize))) return rc;
  HMSG_CUDA(cudaMemcpyAsync(out, vs->out_stage, (size_t)B * vs->desc.out_dim * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_gemm_f16_debug(hmsg_ctx* ctx, const void* A, const void* Wt, float* C, int32_t M, int32_t N, int32_t K) {
  if (!ctx) return HMSG_ERR_ARG;
  VitState tmp;
  int32_t rc = get_encoder(ctx, &tmp.encode);
  if (rc) return rc;
  tmp.desc.quick_gelu = 0;
  rc = gemm(ctx, &tmp, EPI_F32_STORE, (const __half*)A, (const __half*)Wt, M, N, K, nullptr, C, N);
  if (rc) return rc;
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}
