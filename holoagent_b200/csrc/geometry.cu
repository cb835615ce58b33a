// geometry.cu - A1..A4 of the HMSG build path on sm_100a:
//   unprojection, global bounds, voxel occupancy bitmap + popcount-rank index (a minimal
//   perfect hash of the occupied voxel grid), voxel accumulation, radius-outlier filter and
//   the exact pixel->node nearest-neighbour search.
// Reference: fsr_vln/memory/hmsg/dataloader/generic.py:74-138; graph/graph.py:339-364,:409.
//
// Data layout in HBM (all owned by hmsg_ctx):
//   depth u16 [F,H,W] | rgb u8 [F,H,W,3] | poses f64 [F,16]          resident frame store
//   bitmap u32 [nx*ny*nzp/32]   1 bit per voxel cell, k fastest, nzp = roundup(nz,32)
//   prefix u32 [nwords]         exclusive popcount prefix => rank(cell) = canonical index
//   vox_acc f64 [n_voxels,6]    sum (then mean) of xyz and rgb/255 ; vox_cnt u32
//   node_* compact copies of the voxels surviving the radius filter + nbitmap/nprefix
// Unprojected points are never written to HBM: every consumer (bounds, mark, accumulate,
// nearest-node) recomputes them from the 5 B/pixel inputs, so the geometry passes are
// read-only streams over depth/rgb.
#include "common.cuh"
#include <algorithm>
#include <cmath>

#define TPB 256

// ----------------------------------------------------------------------------- helpers
__device__ __forceinline__ double warp_min_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One block covers TPB*4 consecutive pixels of ONE frame (blockIdx.y = frame in launch).
struct FrameArgs {
  const uint16_t* depth;
  const uint8_t* rgb;
  const double* poses;
  long long frame0;   // first frame of this launch
  CamDesc cam;
  const double* frameK;   // per-frame (fx, fy, cx, cy) [cap,4] or nullptr: one K for the scene (dataloader/iphone.py:290-367 reads K per frame)
};

// sT[0..15] = pose of frame f, sT[16..19] = its intrinsics when the scene carries per-frame K
__device__ __forceinline__ void load_pose(const FrameArgs& a, long long f, double* sT) {
  if (threadIdx.x < 16) sT[threadIdx.x] = a.poses[f * 16 + threadIdx.x];
  else if (threadIdx.x < 20 && a.frameK) sT[threadIdx.x] = a.frameK[f * 4 + threadIdx.x - 16];
  __syncthreads();
}
__device__ __forceinline__ CamDesc frame_cam(const FrameArgs& a, const double* sT) {
  CamDesc c = a.cam;
  if (a.frameK) { c.fx = sT[16]; c.fy = sT[17]; c.cx = sT[18]; c.cy = sT[19]; }
  return c;
}

// ----------------------------------------------------------------------------- A1 dense
__global__ void __launch_bounds__(TPB) k_unproject_dense(FrameArgs a, double* xyz, double* rgbo, uint8_t* valid) {
  __shared__ double sT[20];
  long long f = a.frame0;
  load_pose(a, f, sT);
  const CamDesc cam = frame_cam(a, sT);
  int HW = a.cam.H * a.cam.W;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  unsigned short dep = a.depth[f * HW + p];
  int y = p / a.cam.W, x = p - y * a.cam.W;
  double wx = 0, wy = 0, wz = 0, r = 0, g = 0, b = 0;
  // generic.py:117  mask = depth_f32 > 0
  bool ok = __fdiv_rn((float)dep, a.cam.scale) > 0.0f;
  if (ok) {
    unproject_px(dep, x, y, cam, sT, wx, wy, wz);
    const uint8_t* c = a.rgb + (f * HW + p) * 3;
    r = __ddiv_rn((double)c[0], 255.0);
    g = __ddiv_rn((double)c[1], 255.0);
    b = __ddiv_rn((double)c[2], 255.0);
  }
  xyz[p * 3 + 0] = wx; xyz[p * 3 + 1] = wy; xyz[p * 3 + 2] = wz;
  rgbo[p * 3 + 0] = r; rgbo[p * 3 + 1] = g; rgbo[p * 3 + 2] = b;
  valid[p] = ok ? 1 : 0;
}

// ----------------------------------------------------------------------------- bounds
// grid = (blocks_per_frame, n_frames).  Thread handles 4 consecutive pixels (W % 4 == 0 not
// required: pixels are addressed linearly inside the frame).
__global__ void __launch_bounds__(TPB) k_bounds(FrameArgs a, long long* bounds) {
  __shared__ double sT[20];
  __shared__ double red[6][TPB / 32];
  long long f = a.frame0 + blockIdx.y;
  load_pose(a, f, sT);
  const CamDesc cam = frame_cam(a, sT);
  int HW = a.cam.H * a.cam.W;
  int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (p0 < HW) {
    unsigned short dv[4] = {0, 0, 0, 0};
    const uint16_t* dp = a.depth + f * HW + p0;
    if (p0 + 3 < HW && ((((size_t)dp) & 7) == 0)) {
      uint2 u = *reinterpret_cast<const uint2*>(dp);
      dv[0] = u.x & 0xffff; dv[1] = u.x >> 16; dv[2] = u.y & 0xffff; dv[3] = u.y >> 16;
    } else {
      for (int i = 0; i < 4; i++) if (p0 + i < HW) dv[i] = dp[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (__fdiv_rn((float)dv[i], a.cam.scale) > 0.0f) {
        int p = p0 + i;
        int y = p / a.cam.W, x = p - y * a.cam.W;
        double w[3];
        unproject_px(dv[i], x, y, cam, sT, w[0], w[1], w[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) { mn[k] = fmin(mn[k], w[k]); mx[k] = fmax(mx[k], w[k]); }
      }
    }
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double a0 = warp_min_d(mn[k]), a1 = warp_max_d(mx[k]);
    if (lane == 0) { red[k][wid] = a0; red[3 + k][wid] = a1; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int k = threadIdx.x;
    double v = red[k][0];
    for (int w = 1; w < TPB / 32; w++) v = (k < 3) ? fmin(v, red[k][w]) : fmax(v, red[k][w]);
    if (k < 3) { if (v != INFINITY) atomicMin(&bounds[k], d2ord(v)); }
    else       { if (v != -INFINITY) atomicMax(&bounds[k], d2ord(v)); }
  }
}

// ----------------------------------------------------------------------------- cell ids
__device__ __forceinline__ bool cell_of(const GridDesc& g, double wx, double wy, double wz, int& ci, int& cj, int& ck) {
  ci = cell_index(wx, g.vmin[0], g.vs, g.inv_vs);
  cj = cell_index(wy, g.vmin[1], g.vs, g.inv_vs);
  ck = cell_index(wz, g.vmin[2], g.vs, g.inv_vs);
  return ci >= 0 && cj >= 0 && ck >= 0 && ci < g.nx && cj < g.ny && ck < g.nz;
}
__device__ __forceinline__ long long cell_lin(const GridDesc& g, int ci, int cj, int ck) {
  return ((long long)ci * g.ny + cj) * g.nzp + ck;
}

__device__ __forceinline__ void load4_depth(const uint16_t* dp, int p0, int HW, unsigned short* dv) {
  dv[0] = dv[1] = dv[2] = dv[3] = 0;
  if (p0 + 3 < HW && ((((size_t)dp) & 7) == 0)) {
    uint2 u = *reinterpret_cast<const uint2*>(dp);
    dv[0] = u.x & 0xffff; dv[1] = u.x >> 16; dv[2] = u.y & 0xffff; dv[3] = u.y >> 16;
  } else {
    for (int i = 0; i < 4; i++) if (p0 + i < HW) dv[i] = dp[i];
  }
}

// pass 2a: mark occupancy.  Test-then-set: after the first few frames nearly every bit is
// already set, so the pass degenerates into cached bitmap reads (no atomics).
__global__ void __launch_bounds__(TPB) k_mark(FrameArgs a, GridDesc g, uint32_t* bitmap) {
  __shared__ double sT[20];
  long long f = a.frame0 + blockIdx.y;
  load_pose(a, f, sT);
  const CamDesc cam = frame_cam(a, sT);
  int HW = a.cam.H * a.cam.W;
  int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (p0 >= HW) return;
  unsigned short dv[4];
  load4_depth(a.depth + f * HW + p0, p0, HW, dv);
  long long lastw = -1; uint32_t pend = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (__fdiv_rn((float)dv[i], a.cam.scale) > 0.0f) {
      int p = p0 + i;
      int y = p / a.cam.W, x = p - y * a.cam.W;
      double wx, wy, wz;
      unproject_px(dv[i], x, y, cam, sT, wx, wy, wz);
      int ci, cj, ck;
      if (cell_of(g, wx, wy, wz, ci, cj, ck)) {
        long long lin = cell_lin(g, ci, cj, ck);
        long long w = lin >> 5; uint32_t bit = 1u << (lin & 31);
        if (w != lastw) {
          if (pend) { if ((bitmap[lastw] & pend) != pend) atomicOr(&bitmap[lastw], pend); }
          lastw = w; pend = 0;
        }
        pend |= bit;
      }
    }
  }
  if (pend) { if ((bitmap[lastw] & pend) != pend) atomicOr(&bitmap[lastw], pend); }
}

// ----------------------------------------------------------------------------- scan
// 3-phase exclusive scan of per-word popcounts. 1024 words per block.
__global__ void __launch_bounds__(TPB) k_popc_blocks(const uint32_t* bitmap, long long nwords, uint32_t* blocksums) {
  __shared__ uint32_t red[TPB / 32];
  long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) if (base + i < nwords) s += __popc(bitmap[base + i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < TPB / 32; w++) t += red[w];
    blocksums[blockIdx.x] = t;
  }
}
// single block: in-place exclusive scan of blocksums; total -> blocksums[nb]
__global__ void __launch_bounds__(1024) k_scan_blocksums(uint32_t* blocksums, long long nb) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (long long base = 0; base < nb; base += 1024) {
    long long i = base + threadIdx.x;
    uint32_t v = (i < nb) ? blocksums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t incl = sh[threadIdx.x];
    uint32_t c = carry;
    if (i < nb) blocksums[i] = c + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) blocksums[nb] = carry;
}
__global__ void __launch_bounds__(TPB) k_prefix_write(const uint32_t* bitmap, long long nwords, const uint32_t* blocksums, uint32_t* prefix) {
  __shared__ uint32_t wsum[TPB / 32];
  long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
  uint32_t c[4]; uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) { c[i] = (base + i < nwords) ? __popc(bitmap[base + i]) : 0; s += c[i]; }
  // inclusive warp scan of s
  uint32_t incl = s;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < wid; w++) woff += wsum[w];
  uint32_t run = blocksums[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int i = 0; i < 4; i++) { if (base + i < nwords) prefix[base + i] = run; run += c[i]; }
}

// ----------------------------------------------------------------------------- accumulate
// pass 2b: per valid pixel, rank = canonical voxel index; accumulate xyz, rgb/255 and count.
// A thread merges runs of equal rank among its 4 pixels before issuing atomics; then the
// warp merges runs of equal rank across neighbouring lanes with a segmented shuffle scan so
// that one lane per run issues the 7 reductions (neighbouring pixels share voxels heavily).
__global__ void __launch_bounds__(TPB) k_accumulate(FrameArgs a, GridDesc g, const uint32_t* bitmap, const uint32_t* prefix,
                                                    double* acc, uint32_t* cnt) {
  __shared__ double sT[20];
  long long f = a.frame0 + blockIdx.y;
  load_pose(a, f, sT);
  const CamDesc cam = frame_cam(a, sT);
  int HW = a.cam.H * a.cam.W;
  int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  unsigned short dv[4] = {0, 0, 0, 0};
  if (p0 < HW) load4_depth(a.depth + f * HW + p0, p0, HW, dv);
  long long cur = -1; double s[6] = {0, 0, 0, 0, 0, 0}; uint32_t n = 0;
  auto flush = [&]() {
    if (cur >= 0) {
      double* dst = acc + cur * 6;
#pragma unroll
      for (int k = 0; k < 6; k++) atomicAdd(dst + k, s[k]);
      atomicAdd(cnt + cur, n);
    }
  };
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (__fdiv_rn((float)dv[i], a.cam.scale) > 0.0f) {
      int p = p0 + i;
      int y = p / a.cam.W, x = p - y * a.cam.W;
      double w[3];
      unproject_px(dv[i], x, y, cam, sT, w[0], w[1], w[2]);
      int ci, cj, ck;
      if (cell_of(g, w[0], w[1], w[2], ci, cj, ck)) {
        long long lin = cell_lin(g, ci, cj, ck);
        long long wd = lin >> 5; int b = lin & 31;
        long long rank = (long long)prefix[wd] + __popc(bitmap[wd] & ((1u << b) - 1u));
        const uint8_t* c = a.rgb + (f * HW + p) * 3;
        double col[3] = {__ddiv_rn((double)c[0], 255.0), __ddiv_rn((double)c[1], 255.0), __ddiv_rn((double)c[2], 255.0)};
        if (rank != cur) { flush(); cur = rank; n = 0;
#pragma unroll
          for (int k = 0; k < 6; k++) s[k] = 0; }
#pragma unroll
        for (int k = 0; k < 3; k++) { s[k] += w[k]; s[3 + k] += col[k]; }
        n++;
      }
    }
  }
  flush();
}

__global__ void __launch_bounds__(TPB) k_finalize_voxels(double* acc, const uint32_t* cnt, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double c = (double)cnt[i];
#pragma unroll
  for (int k = 0; k < 6; k++) acc[i * 6 + k] = __ddiv_rn(acc[i * 6 + k], c);
}

// enumerate set bits -> ijk per rank
__global__ void __launch_bounds__(TPB) k_write_ijk(const uint32_t* bitmap, const uint32_t* prefix, GridDesc g, int32_t* ijk) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= g.nwords) return;
  uint32_t bits = bitmap[w];
  if (!bits) return;
  long long r = prefix[w];
  long long lin0 = w << 5;
  int k0 = (int)(lin0 % g.nzp);
  long long ij = lin0 / g.nzp;
  int j = (int)(ij % g.ny), i = (int)(ij / g.ny);
  while (bits) {
    int b = __ffs(bits) - 1;
    bits &= bits - 1;
    ijk[r * 3 + 0] = i; ijk[r * 3 + 1] = j; ijk[r * 3 + 2] = k0 + b;
    r++;
  }
}

// ----------------------------------------------------------------------------- A3 radius
// One warp per voxel: lanes sweep the (i,j) columns of the bounding cube, walk the bitmap
// words of each column, fetch candidate centroids by rank and count d2 < R^2 in float64
// with the accumulation order ((dx*dx+dy*dy)+dz*dz) (Open3D/nanoflann L2_Simple).
__global__ void __launch_bounds__(TPB) k_radius_count(GridDesc g, const uint32_t* bitmap, const uint32_t* prefix, const double* acc,
                                                      const int32_t* ijk, long long v_begin, long long v_end, double R, uint32_t* out) {
  long long v = v_begin + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (v >= v_end) return;
  double cx = acc[v * 6 + 0], cy = acc[v * 6 + 1], cz = acc[v * 6 + 2];
  int ci = ijk[v * 3 + 0], cj = ijk[v * 3 + 1], ck = ijk[v * 3 + 2];
  double R2 = __dmul_rn(R, R);
  int Rc = (int)ceil(R / g.vs) + 1;
  int side = 2 * Rc + 1;
  double lim = R2 * 1.000001 / (g.vs * g.vs);   // in cell units^2, with slack
  uint32_t count = 0;
  for (int t = lane; t < side * side; t += 32) {
    int di = t / side - Rc, dj = t % side - Rc;
    int ni = ci + di, nj = cj + dj;
    if (ni < 0 || nj < 0 || ni >= g.nx || nj >= g.ny) continue;
    int gi = max(abs(di) - 1, 0), gj = max(abs(dj) - 1, 0);
    double g2 = (double)(gi * gi + gj * gj);
    if (g2 > lim) continue;
    int kr = (int)floor(sqrt(lim - g2)) + 2;
    int k0 = max(ck - kr, 0), k1 = min(ck + kr, g.nz - 1);
    long long colbase = ((long long)ni * g.ny + nj) * g.nzp;
    for (int wk = (k0 >> 5); wk <= (k1 >> 5); wk++) {
      long long w = (colbase >> 5) + wk;
      uint32_t word = __ldg(&bitmap[w]);
      if (!word) continue;
      int lo = max(k0 - (wk << 5), 0), hi = min(k1 - (wk << 5), 31);
      uint32_t m = word & (0xffffffffu << lo) & (0xffffffffu >> (31 - hi));
      uint32_t pf = __ldg(&prefix[w]);
      while (m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        long long r = (long long)pf + __popc(word & ((1u << b) - 1u));
        const double* q = acc + r * 6;
        double d2 = sqdist3(q[0], q[1], q[2], cx, cy, cz);
        count += (d2 < R2) ? 1u : 0u;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  if (lane == 0) out[v] = count;
}

// keep flag -> cleared bitmap of kept voxels
__global__ void __launch_bounds__(TPB) k_node_bitmap(const uint32_t* bitmap, const uint32_t* prefix, long long nwords,
                                                     const uint32_t* rad_cnt, uint32_t nb_points, uint32_t* nbitmap) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint32_t bits = bitmap[w], outb = 0;
  long long r = prefix[w];
  while (bits) {
    int b = __ffs(bits) - 1;
    bits &= bits - 1;
    if (rad_cnt[r] > nb_points) outb |= 1u << b;
    r++;
  }
  nbitmap[w] = outb;
}
// compact voxel rows into node rows
__global__ void __launch_bounds__(TPB) k_compact_nodes(const uint32_t* bitmap, const uint32_t* prefix, const uint32_t* nbitmap,
                                                       const uint32_t* nprefix, long long nwords, const double* acc, const int32_t* ijk,
                                                       double* nxyz, double* nrgb, int32_t* nijk, int64_t* nvox) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint32_t bits = bitmap[w], nb = nbitmap[w];
  if (!nb) return;
  long long r = prefix[w], q = nprefix[w];
  while (bits) {
    int b = __ffs(bits) - 1;
    bits &= bits - 1;
    if (nb & (1u << b)) {
#pragma unroll
      for (int k = 0; k < 3; k++) { nxyz[q * 3 + k] = acc[r * 6 + k]; nrgb[q * 3 + k] = acc[r * 6 + 3 + k]; nijk[q * 3 + k] = ijk[r * 3 + k]; }
      nvox[q] = r;
      q++;
    }
    r++;
  }
}

// ----------------------------------------------------------------------------- A4 NN
// Exact nearest node by ring expansion over the occupancy grid.  Cells of Chebyshev ring r
// around the query's cell are visited; a cell is skipped when the distance from the query to
// the cell's box already exceeds the best distance; the search stops once the best distance
// is below r*vs (every unvisited cell is at least that far).  Ties -> lower node index.
// Far search (exact): Chebyshev ring expansion over the COARSE grid (4x4x4 blocks of cells).  A coarse
// cell is skipped when its bit is clear or when the distance from the query to its box already exceeds
// the best distance; occupied coarse cells are scanned cell by cell (4 bits of one bitmap word per
// column).  The search stops when the best distance is below the distance to the next coarse shell.
__device__ __noinline__ int nn_search_rings(const GridDesc& g, const uint32_t* __restrict__ bm, const uint32_t* __restrict__ pf,
                                           const uint32_t* __restrict__ cbm, const double* __restrict__ nodes, double px, double py, double pz,
                                           double& best_out) {
  const double fx = (px - g.vmin[0]) * g.inv_vs, fy = (py - g.vmin[1]) * g.inv_vs, fz = (pz - g.vmin[2]) * g.inv_vs;   // fine cell units
  const int bi = min(max((int)floor(fx * 0.25), 0), g.cnx - 1);
  const int bj = min(max((int)floor(fy * 0.25), 0), g.cny - 1);
  const int bk = min(max((int)floor(fz * 0.25), 0), g.cnz - 1);
  double best = INFINITY; int besti = -1;
  const double vs2 = g.vs * g.vs;
  const int maxr = max(g.cnx, max(g.cny, g.cnz));
  for (int r = 0; r <= maxr; r++) {
    if (besti >= 0 && r >= 1) {
      // a coarse cell of shell >= r is at least (r-1) coarse cells = 4(r-1) fine cells away (the query may sit anywhere in / outside its clamped base)
      double lbr = 4.0 * (double)(r - 1) * g.vs;
      if (best < lbr * lbr) break;
    }
    const int i0 = max(bi - r, 0), i1 = min(bi + r, g.cnx - 1);
    const int j0 = max(bj - r, 0), j1 = min(bj + r, g.cny - 1);
    for (int ci = i0; ci <= i1; ci++) {
      const double ax = fmax(0.0, fmax(4.0 * ci - fx, fx - 4.0 * (ci + 1)));
      const int adi = abs(ci - bi);
      for (int cj = j0; cj <= j1; cj++) {
        const double ay = fmax(0.0, fmax(4.0 * cj - fy, fy - 4.0 * (cj + 1)));
        const double axy = ax * ax + ay * ay;
        if (axy * vs2 * 0.999999999 > best) continue;
        const bool shell = (adi == r) || (abs(cj - bj) == r);
        const long long ccol = ((long long)ci * g.cny + cj) * g.cnzp;
        const int nk = shell ? (2 * r + 1) : 2;
        for (int t = 0; t < nk; t++) {
          const int ck = shell ? (bk - r + t) : (t == 0 ? bk - r : bk + r);
          if (ck < 0 || ck >= g.cnz) continue;
          if (!shell && r == 0 && t == 1) continue;
          const long long cl = ccol + ck;
          if (!((__ldg(&cbm[cl >> 5]) >> (cl & 31)) & 1u)) continue;
          const double az = fmax(0.0, fmax(4.0 * ck - fz, fz - 4.0 * (ck + 1)));
          if ((axy + az * az) * vs2 * 0.999999999 > best) continue;
          // scan the 4x4x4 fine cells of this block
          for (int ii = 4 * ci; ii < min(4 * ci + 4, g.nx); ii++) {
            const double bx = fmax(0.0, fmax((double)ii - fx, fx - (double)(ii + 1)));
            for (int jj = 4 * cj; jj < min(4 * cj + 4, g.ny); jj++) {
              const double by = fmax(0.0, fmax((double)jj - fy, fy - (double)(jj + 1)));
              const double bxy = bx * bx + by * by;
              if (bxy * vs2 * 0.999999999 > best) continue;
              const long long lin0 = ((long long)ii * g.ny + jj) * g.nzp + 4 * ck;     // 4-aligned: the 4 k-cells share one word
              const uint32_t word = __ldg(&bm[lin0 >> 5]);
              uint32_t nib = (word >> (lin0 & 31)) & 0xFu;
              if (!nib) continue;
              const uint32_t pre = __ldg(&pf[lin0 >> 5]);
              while (nib) {
                const int kb = __ffs(nib) - 1;
                nib &= nib - 1;
                const int b = (int)(lin0 & 31) + kb;
                const int idx = (int)(pre + __popc(word & ((1u << b) - 1u)));
                const double* q = nodes + (long long)idx * 3;
                const double d2 = sqdist3(q[0], q[1], q[2], px, py, pz);
                if (d2 < best || (d2 == best && idx < besti)) { best = d2; besti = idx; }
              }
            }
          }
        }
      }
    }
  }
  best_out = best;
  return besti;
}

// Fast path of the exact nearest-node search.  Almost every pixel finds its nearest centroid in
// its own voxel or one of the 26 neighbours, and most of those cells are empty on surfaces, so:
// (1) gather the 3x3x3 occupancy bits from 9 bitmap columns, (2) visit only occupied cells, own cell
// first, pruning with an fp32 lower bound of the point-to-cell-box distance (with slack), exact
// float64 distance for the survivors, (3) accept if the best distance is below one voxel edge (every
// cell outside the 3x3x3 block is at least that far); otherwise fall back to the generic ring search.
// Keeps all lanes of a warp on one short straight-line path (the generic search ran at 15.8/32 lanes).
__device__ __forceinline__ int nn_search_fast(const GridDesc& g, const uint32_t* __restrict__ bm, const uint32_t* __restrict__ pf,
                                              const double* __restrict__ nodes, double px, double py, double pz, double& best_out) {
  const double fxd = (px - g.vmin[0]) * g.inv_vs, fyd = (py - g.vmin[1]) * g.inv_vs, fzd = (pz - g.vmin[2]) * g.inv_vs;
  const int bi = (int)floor(fxd), bj = (int)floor(fyd), bk = (int)floor(fzd);
  if (bi < 0 || bj < 0 || bk < 0 || bi >= g.nx || bj >= g.ny || bk >= g.nz) return -2;
  const float ux = (float)(fxd - bi), uy = (float)(fyd - bj), uz = (float)(fzd - bk);   // position inside the own cell, [0,1)
  // per-axis gap (in cells) to the neighbour at offset -1 / 0 / +1
  auto gap = [](int dd, float u) { return dd < 0 ? u : (dd > 0 ? 1.f - u : 0.f); };
  const float vs2f = (float)(g.vs * g.vs) * 0.9999f;      // slack: the bound must never exceed the true distance
  double best = INFINITY; int besti = -1;
  float bestf = INFINITY;
  // visit order: own column first (dj = di = 0), then the rest
#pragma unroll 1
  for (int c = 0; c < 9; c++) {
    const int cc = (c == 0) ? 4 : (c <= 4 ? c - 1 : c);   // 4,0,1,2,3,5,6,7,8
    const int di = cc / 3 - 1, dj = cc % 3 - 1;
    const int ci = bi + di, cj = bj + dj;
    if (ci < 0 || cj < 0 || ci >= g.nx || cj >= g.ny) continue;
    const float gxi = gap(di, ux), gyj = gap(dj, uy);
    const float axy = gxi * gxi + gyj * gyj;
    if (axy * vs2f > bestf) continue;
    const long long colbase = ((long long)ci * g.ny + cj) * g.nzp;
    // bits for k = bk-1, bk, bk+1 (a column never crosses its own nzp-padded words, but bk-1/bk+1 may sit in the neighbouring word)
#pragma unroll
    for (int t = 0; t < 3; t++) {
      const int dk = (t == 0) ? 0 : (t == 1 ? -1 : 1);     // own k first
      const int ck = bk + dk;
      if (ck < 0 || ck >= g.nz) continue;
      const long long lin = colbase + ck;
      const uint32_t word = __ldg(&bm[lin >> 5]);
      const int b = (int)(lin & 31);
      if (!((word >> b) & 1u)) continue;
      const float gzk = gap(dk, uz);
      const float lb = (axy + gzk * gzk) * vs2f;
      if (lb > bestf) continue;
      const int idx = (int)(__ldg(&pf[lin >> 5]) + __popc(word & ((1u << b) - 1u)));
      const double* q = nodes + (long long)idx * 3;
      const double d2 = sqdist3(q[0], q[1], q[2], px, py, pz);
      if (d2 < best || (d2 == best && idx < besti)) { best = d2; besti = idx; bestf = (float)d2 * 1.0001f; }
    }
  }
  // every cell outside the 3x3x3 block is farther than one voxel edge minus nothing: >= (1 + min gap) * vs >= vs
  if (besti >= 0 && best < g.vs * g.vs) { best_out = best; return besti; }
  return -2;   // needs the far search
}

__device__ __forceinline__ int nn_search(const GridDesc& g, const uint32_t* __restrict__ bm, const uint32_t* __restrict__ pf,
                                         const uint32_t* __restrict__ cbm, const double* __restrict__ nodes, double px, double py, double pz,
                                         double& best_out) {
  int r = nn_search_fast(g, bm, pf, nodes, px, py, pz, best_out);
  if (r == -2) r = nn_search_rings(g, bm, pf, cbm, nodes, px, py, pz, best_out);
  return r;
}

// per-frame API kernel: idx int64 (-1 invalid), dist f64
__global__ void __launch_bounds__(TPB) k_pixel_to_node(FrameArgs a, GridDesc g, const uint32_t* bm, const uint32_t* pf, const uint32_t* cbm,
                                                       const double* nodes, int64_t* idx, double* dist) {
  __shared__ double sT[20];
  long long f = a.frame0;
  load_pose(a, f, sT);
  const CamDesc cam = frame_cam(a, sT);
  int HW = a.cam.H * a.cam.W;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  unsigned short dep = a.depth[f * HW + p];
  long long o = -1; double dd = 0.0;
  if (__fdiv_rn((float)dep, a.cam.scale) > 0.0f) {
    int y = p / a.cam.W, x = p - y * a.cam.W;
    double wx, wy, wz, best;
    unproject_px(dep, x, y, cam, sT, wx, wy, wz);
    o = nn_search(g, bm, pf, cbm, nodes, wx, wy, wz, best);
    dd = sqrt(best);
  }
  idx[p] = o;
  if (dist) dist[p] = dd;
}

__global__ void __launch_bounds__(TPB) k_points_to_node(const double* pts, long long n, GridDesc g, const uint32_t* bm, const uint32_t* pf,
                                                        const uint32_t* cbm, const double* nodes, int64_t* idx, double* dist) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double best;
  int o = nn_search(g, bm, pf, cbm, nodes, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], best);
  idx[i] = o;
  if (dist) dist[i] = sqrt(best);
}

// batch kernels: NN + last-writer-wins winner election.
// phase 1 (grid = (blocks, n_frames)): fast 3x3x3 path for every valid pixel; pixels that need the far
// search (own neighbourhood filtered out) are appended to a worklist so that phase 2 runs them densely
// instead of stalling 31 finished lanes per straggler.
__global__ void __launch_bounds__(TPB) k_nn_winner(FrameArgs a, GridDesc g, const uint32_t* bm, const uint32_t* pf, const double* nodes,
                                                   int32_t* pix_idx, unsigned long long* win, long long n_nodes, uint32_t epoch, int* far_list,
                                                   int* far_count) {
  __shared__ double sT[20];
  int fb = blockIdx.y;
  long long f = a.frame0 + fb;
  load_pose(a, f, sT);
  const CamDesc cam = frame_cam(a, sT);
  int HW = a.cam.H * a.cam.W;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  unsigned short dep = a.depth[f * HW + p];
  int o = -1;
  if (__fdiv_rn((float)dep, a.cam.scale) > 0.0f) {
    int y = p / a.cam.W, x = p - y * a.cam.W;
    double wx, wy, wz, best;
    unproject_px(dep, x, y, cam, sT, wx, wy, wz);
    o = nn_search_fast(g, bm, pf, nodes, wx, wy, wz, best);
    if (o >= 0) atomicMax(&win[(long long)fb * n_nodes + o], ((unsigned long long)epoch << 32) | (unsigned)p);
    else far_list[atomicAdd(far_count, 1)] = fb * HW + p;
  }
  pix_idx[(long long)fb * HW + p] = o;
}

__global__ void __launch_bounds__(TPB) k_nn_far(FrameArgs a, GridDesc g, const uint32_t* bm, const uint32_t* pf, const uint32_t* cbm,
                                                const double* nodes, int32_t* pix_idx, unsigned long long* win, long long n_nodes, uint32_t epoch,
                                                const int* far_list, const int* far_count) {
  int n = *far_count;
  int HW = a.cam.H * a.cam.W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int e = far_list[i];
    int fb = e / HW, p = e - fb * HW;
    long long f = a.frame0 + fb;
    double T[16];
#pragma unroll
    for (int k = 0; k < 16; k++) T[k] = a.poses[f * 16 + k];
    unsigned short dep = a.depth[f * HW + p];
    int y = p / a.cam.W, x = p - y * a.cam.W;
    double wx, wy, wz, best;
    CamDesc cam = a.cam;
    if (a.frameK) { cam.fx = a.frameK[f * 4]; cam.fy = a.frameK[f * 4 + 1]; cam.cx = a.frameK[f * 4 + 2]; cam.cy = a.frameK[f * 4 + 3]; }
    unproject_px(dep, x, y, cam, T, wx, wy, wz);
    int o = nn_search_rings(g, bm, pf, cbm, nodes, wx, wy, wz, best);
    if (o >= 0) atomicMax(&win[(long long)fb * n_nodes + o], ((unsigned long long)epoch << 32) | (unsigned)p);
    pix_idx[(long long)fb * HW + p] = o;
  }
}

// coarse occupancy: one thread per fine bitmap word (32 k-cells of one column) -> up to 8 coarse bits
__global__ void __launch_bounds__(TPB) k_coarse_mark(const uint32_t* __restrict__ nbitmap, GridDesc g, uint32_t* cbitmap) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= g.nwords) return;
  uint32_t word = nbitmap[w];
  if (!word) return;
  long long lin0 = w << 5;
  int k0 = (int)(lin0 % g.nzp);
  long long ij = lin0 / g.nzp;
  int j = (int)(ij % g.ny), i = (int)(ij / g.ny);
  for (int nb = 0; nb < 8; nb++) {
    if ((word >> (4 * nb)) & 0xFu) {
      long long cl = ((long long)(i >> 2) * g.cny + (j >> 2)) * g.cnzp + ((k0 >> 2) + nb);
      atomicOr(&cbitmap[cl >> 5], 1u << (cl & 31));
    }
  }
}

// ======================================================================================
// host side
// ======================================================================================
static FrameArgs frame_args(hmsg_ctx* ctx, long long frame0) {
  FrameArgs a;
  a.depth = ctx->depth; a.rgb = ctx->rgb; a.poses = ctx->poses; a.frame0 = frame0; a.cam = ctx->cam; a.frameK = ctx->frameK;
  return a;
}

extern "C" int32_t hmsg_scene_begin(hmsg_ctx* ctx, int32_t H, int32_t W, const double K[9], float depth_scale,
                                    double voxel_size, int64_t frame_capacity) {
  if (!ctx) return HMSG_ERR_ARG;
  if (H <= 0 || W <= 0 || !K || depth_scale <= 0 || voxel_size <= 0 || frame_capacity <= 0)
    return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_begin: bad argument");
  HMSG_CUDA(cudaSetDevice(ctx->device));
  free_dev(ctx->depth); free_dev(ctx->rgb); free_dev(ctx->poses); free_dev(ctx->frameK);
  ctx->cam.H = H; ctx->cam.W = W; ctx->cam.fx = K[0]; ctx->cam.fy = K[4]; ctx->cam.cx = K[2]; ctx->cam.cy = K[5];
  ctx->cam.scale = depth_scale; ctx->vs = voxel_size;
  ctx->cap = frame_capacity; ctx->nframes = 0;
  if (!ctx->uploads.empty()) HMSG_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  ctx->uploads.clear(); ctx->upload_events_used = 0;
  size_t hw = (size_t)H * W;
  HMSG_CUDA(cudaMalloc((void**)&ctx->depth, hw * 2 * frame_capacity));
  HMSG_CUDA(cudaMalloc((void**)&ctx->rgb, hw * 3 * frame_capacity));
  HMSG_CUDA(cudaMalloc((void**)&ctx->poses, 16 * sizeof(double) * frame_capacity));
  if (!ctx->d_bounds) HMSG_CUDA(cudaMalloc((void**)&ctx->d_bounds, 6 * sizeof(long long)));
  ctx->voxels_built = false; ctx->nodes_built = false; ctx->n_voxels = 0; ctx->n_nodes = 0;
  ctx->batch_begin = -1;
  return HMSG_OK;
}

// per-frame intrinsics (dataloader/iphone.py:290-367: `camera_matrix = np.array(self.frames[image_id - 1]["K"])` inside create_pcd)
extern "C" int32_t hmsg_scene_set_intrinsics(hmsg_ctx* ctx, int64_t frame_begin, int32_t n, const double* K) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->depth) return ctx->fail(HMSG_ERR_STATE, "hmsg_scene_set_intrinsics: call hmsg_scene_begin first");
  if (!K || n <= 0 || frame_begin < 0 || frame_begin + n > ctx->cap) return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_set_intrinsics: bad argument");
  if (!ctx->frameK) {   // every frame starts with the scene's K
    HMSG_CUDA(cudaMalloc((void**)&ctx->frameK, (size_t)ctx->cap * 32));
    std::vector<double> init((size_t)ctx->cap * 4);
    for (int64_t f = 0; f < ctx->cap; f++) { init[f * 4] = ctx->cam.fx; init[f * 4 + 1] = ctx->cam.fy; init[f * 4 + 2] = ctx->cam.cx; init[f * 4 + 3] = ctx->cam.cy; }
    HMSG_CUDA(cudaMemcpy(ctx->frameK, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
  }
  std::vector<double> k4((size_t)n * 4);
  for (int i = 0; i < n; i++) { k4[i * 4] = K[i * 9]; k4[i * 4 + 1] = K[i * 9 + 4]; k4[i * 4 + 2] = K[i * 9 + 2]; k4[i * 4 + 3] = K[i * 9 + 5]; }
  HMSG_CUDA(cudaMemcpyAsync(ctx->frameK + frame_begin * 4, k4.data(), k4.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->voxels_built = false; ctx->nodes_built = false;
  return HMSG_OK;
}

extern "C" int32_t hmsg_scene_add_frames(hmsg_ctx* ctx, const uint16_t* depth, const uint8_t* rgb, const double* poses,
                                         int32_t n, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->depth) return ctx->fail(HMSG_ERR_STATE, "hmsg_scene_add_frames: call hmsg_scene_begin first");
  if (!depth || !rgb || !poses || n < 0) return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_add_frames: bad argument");
  if (ctx->nframes + n > ctx->cap) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_scene_add_frames: frame capacity exceeded");
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  HMSG_CUDA(cudaMemcpyAsync(ctx->depth + hw * ctx->nframes, depth, hw * 2 * n, kind, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(ctx->rgb + hw * 3 * ctx->nframes, rgb, hw * 3 * n, kind, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(ctx->poses + 16 * ctx->nframes, poses, 16 * sizeof(double) * n, kind, ctx->stream));
  if (!on_device) HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->nframes += n;
  ctx->voxels_built = false; ctx->nodes_built = false;
  return HMSG_OK;
}

extern "C" int32_t hmsg_scene_put_frames(hmsg_ctx* ctx, int64_t frame_begin, const uint16_t* depth, const uint8_t* rgb, const double* poses,
                                         int32_t n, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->depth) return ctx->fail(HMSG_ERR_STATE, "hmsg_scene_put_frames: call hmsg_scene_begin first");
  if (!depth || !rgb || !poses || n < 0 || frame_begin < 0) return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_put_frames: bad argument");
  if (frame_begin + n > ctx->cap) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_scene_put_frames: frame capacity exceeded");
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  // Host frames go over PCIe on the copy stream so that the upload of later batches overlaps the kernels of earlier
  // ones; an event per call lets every consumer wait for exactly the frames it reads (hmsg_ctx::wait_frames).  The
  // first upload of a scene is ordered after everything the compute stream still reads from the old frames.
  cudaStream_t st = on_device ? ctx->stream : ctx->copy_stream;
  if (!on_device && ctx->uploads.empty()) {
    HMSG_CUDA(cudaEventRecord(ctx->sync_event, ctx->stream));
    HMSG_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->sync_event, 0));
  }
  HMSG_CUDA(cudaMemcpyAsync(ctx->depth + hw * frame_begin, depth, hw * 2 * n, kind, st));
  HMSG_CUDA(cudaMemcpyAsync(ctx->rgb + hw * 3 * frame_begin, rgb, hw * 3 * n, kind, st));
  HMSG_CUDA(cudaMemcpyAsync(ctx->poses + 16 * frame_begin, poses, 16 * sizeof(double) * n, kind, st));
  if (!on_device) {
    if (ctx->upload_events_used == ctx->upload_events.size()) {
      cudaEvent_t e;
      HMSG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->upload_events.push_back(e);
    }
    cudaEvent_t ev = ctx->upload_events[ctx->upload_events_used++];
    HMSG_CUDA(cudaEventRecord(ev, ctx->copy_stream));
    ctx->uploads.push_back(UploadRec{frame_begin, n, ev, false});
  }
  if (frame_begin + n > ctx->nframes) ctx->nframes = frame_begin + n;
  ctx->voxels_built = false; ctx->nodes_built = false;
  return HMSG_OK;
}

extern "C" int32_t hmsg_scene_put_rgb(hmsg_ctx* ctx, int64_t frame_begin, const uint8_t* rgb, int32_t n, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->rgb) return ctx->fail(HMSG_ERR_STATE, "hmsg_scene_put_rgb: call hmsg_scene_begin first");
  if (!rgb || n < 0 || frame_begin < 0) return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_put_rgb: bad argument");
  if (frame_begin + n > ctx->nframes) return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_put_rgb: frames not stored yet");
  // ordered on the compute stream after every kernel that still reads the old colours (the voxel accumulation); the frames'
  // own uploads must have landed first
  ctx->wait_frames(frame_begin, n);
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  HMSG_CUDA(cudaMemcpyAsync(ctx->rgb + hw * 3 * frame_begin, rgb, hw * 3 * n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_scene_set_num_frames(hmsg_ctx* ctx, int64_t n) {
  if (!ctx) return HMSG_ERR_ARG;
  if (n < 0 || n > ctx->cap) return ctx->fail(HMSG_ERR_ARG, "hmsg_scene_set_num_frames: out of range");
  ctx->nframes = n;
  return HMSG_OK;
}

extern "C" int32_t hmsg_scene_reset_frames(hmsg_ctx* ctx) {
  if (!ctx) return HMSG_ERR_ARG;
  ctx->nframes = 0; ctx->voxels_built = false; ctx->nodes_built = false; ctx->batch_begin = -1;
  // pending uploads of the old scene: nothing may still be in flight when their slots are reused
  if (!ctx->uploads.empty()) HMSG_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  ctx->uploads.clear();
  ctx->upload_events_used = 0;
  return HMSG_OK;
}

extern "C" int64_t hmsg_scene_num_frames(const hmsg_ctx* ctx) { return ctx ? ctx->nframes : -1; }

extern "C" int32_t hmsg_unproject_frame(hmsg_ctx* ctx, int64_t frame, double* xyz, double* rgb, uint8_t* valid) {
  if (!ctx) return HMSG_ERR_ARG;
  if (frame < 0 || frame >= ctx->nframes) return ctx->fail(HMSG_ERR_ARG, "hmsg_unproject_frame: frame out of range");
  if (!xyz || !rgb || !valid) return ctx->fail(HMSG_ERR_ARG, "hmsg_unproject_frame: null output");
  ctx->wait_frames(frame, 1);
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  size_t need = hw * (48 + 1) + 64;
  int32_t rc = ctx->reserve((char**)&ctx->scratch, &ctx->scratch_bytes, need);
  if (rc) return rc;
  double* dx = (double*)ctx->scratch; double* dc = dx + hw * 3; uint8_t* dvld = (uint8_t*)(dc + hw * 3);
  k_unproject_dense<<<(unsigned)((hw + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(frame_args(ctx, frame), dx, dc, dvld);
  HMSG_LAUNCH_CHECK();
  HMSG_CUDA(cudaMemcpyAsync(xyz, dx, hw * 24, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(rgb, dc, hw * 24, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(valid, dvld, hw, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

static int32_t run_scan(hmsg_ctx* ctx, const uint32_t* bitmap, long long nwords, uint32_t* prefix, int64_t* total) {
  long long nb = (nwords + 1023) / 1024;
  k_popc_blocks<<<(unsigned)nb, TPB, 0, ctx->stream>>>(bitmap, nwords, ctx->blocksums);
  HMSG_LAUNCH_CHECK();
  k_scan_blocksums<<<1, 1024, 0, ctx->stream>>>(ctx->blocksums, nb);
  HMSG_LAUNCH_CHECK();
  k_prefix_write<<<(unsigned)nb, TPB, 0, ctx->stream>>>(bitmap, nwords, ctx->blocksums, prefix);
  HMSG_LAUNCH_CHECK();
  uint32_t t = 0;
  HMSG_CUDA(cudaMemcpyAsync(&t, ctx->blocksums + nb, 4, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  *total = t;
  return HMSG_OK;
}

// launch a per-frame-block kernel over all frames in chunks of <= 65535 frames (gridDim.y)
template <typename F>
static int32_t for_frame_chunks(hmsg_ctx* ctx, F&& launch) {
  int HW = ctx->cam.H * ctx->cam.W;
  ctx->wait_frames(0, ctx->nframes);
  unsigned bpf = (unsigned)((HW + TPB * 4 - 1) / (TPB * 4));
  for (int64_t f0 = 0; f0 < ctx->nframes; f0 += 32768) {
    unsigned nf = (unsigned)std::min<int64_t>(32768, ctx->nframes - f0);
    launch(dim3(bpf, nf), f0);
    HMSG_LAUNCH_CHECK();
  }
  return HMSG_OK;
}

// launch a per-frame-block kernel over the frame range [f_begin, f_begin+n) in chunks (gridDim.y limit)
template <typename F>
static int32_t for_frame_range(hmsg_ctx* ctx, int64_t f_begin, int64_t n, F&& launch) {
  int HW = ctx->cam.H * ctx->cam.W;
  ctx->wait_frames(f_begin, n);
  unsigned bpf = (unsigned)((HW + TPB * 4 - 1) / (TPB * 4));
  for (int64_t f0 = f_begin; f0 < f_begin + n; f0 += 32768) {
    unsigned nf = (unsigned)std::min<int64_t>(32768, f_begin + n - f0);
    launch(dim3(bpf, nf), f0);
    HMSG_LAUNCH_CHECK();
  }
  return HMSG_OK;
}

static bool range_ok(hmsg_ctx* ctx, int64_t fb, int64_t n) { return fb >= 0 && n >= 0 && fb + n <= ctx->nframes; }

// ---- staged voxel build (each stage works on a frame range so that ranks can split the frames and
// merge with tiny collectives: min/max of 6 doubles, OR of the bitmap, sum of the accumulators)
extern "C" int32_t hmsg_voxel_bounds(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames, double minmax[6]) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!range_ok(ctx, frame_begin, n_frames) || !minmax) return ctx->fail(HMSG_ERR_ARG, "hmsg_voxel_bounds: bad frame range");
  HMSG_CUDA(cudaSetDevice(ctx->device));
  long long init[6];
  {
    double pinf = INFINITY, ninf = -INFINITY;
    long long a, b; memcpy(&a, &pinf, 8); memcpy(&b, &ninf, 8);
    long long op = a, on = b ^ 0x7FFFFFFFFFFFFFFFLL;
    init[0] = init[1] = init[2] = op; init[3] = init[4] = init[5] = on;
  }
  ctx->prof_begin(PROF_GEOM);
  HMSG_CUDA(cudaMemcpyAsync(ctx->d_bounds, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  int32_t rc = for_frame_range(ctx, frame_begin, n_frames, [&](dim3 grid, int64_t f0) {
    k_bounds<<<grid, TPB, 0, ctx->stream>>>(frame_args(ctx, f0), ctx->d_bounds);
  });
  if (rc) return rc;
  ctx->prof_end(PROF_GEOM, (double)n_frames * ctx->cam.H * ctx->cam.W * 2.0);
  long long hb[6];
  HMSG_CUDA(cudaMemcpyAsync(hb, ctx->d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 6; k++) minmax[k] = ord2d_host(hb[k]);
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxel_grid_set(hmsg_ctx* ctx, const double minmax[6]) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!minmax) return ctx->fail(HMSG_ERR_ARG, "hmsg_voxel_grid_set: null bounds");
  for (int k = 0; k < 3; k++) { ctx->min_bound[k] = minmax[k]; ctx->max_bound[k] = minmax[3 + k]; }
  if (!(ctx->min_bound[0] <= ctx->max_bound[0])) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_build: no valid depth pixel in any frame");
  GridDesc& g = ctx->grid;
  g.vs = ctx->vs;
  g.inv_vs = 1.0 / ctx->vs;
  int dims[3];
  for (int k = 0; k < 3; k++) {
    g.vmin[k] = ctx->min_bound[k] - ctx->vs * 0.5;                      // Open3D: min_bound - voxel_size*0.5
    double ext = std::floor((ctx->max_bound[k] - g.vmin[k]) / ctx->vs);
    if (ext > 2.0e6) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_voxel_build: scene extent exceeds 2^21 voxels per axis");
    dims[k] = (int)ext + 1;
  }
  g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2]; g.nzp = (g.nz + 31) & ~31;
  long long ncells = (long long)g.nx * g.ny * g.nzp;
  if (ncells > (1LL << 37)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_voxel_build: dense occupancy index would exceed 16 GiB");
  g.nwords = ncells >> 5;
  g.cnx = (g.nx + 3) / 4; g.cny = (g.ny + 3) / 4; g.cnz = (g.nz + 3) / 4; g.cnzp = (g.cnz + 31) & ~31;
  g.cnwords = ((long long)g.cnx * g.cny * g.cnzp) >> 5;
  if ((size_t)std::max<long long>(g.cnwords, 1) > ctx->cbitmap_cap_words) {
    free_dev(ctx->cbitmap);
    HMSG_CUDA(cudaMalloc((void**)&ctx->cbitmap, (size_t)std::max<long long>(g.cnwords, 1) * 4));
    ctx->cbitmap_cap_words = (size_t)std::max<long long>(g.cnwords, 1);
  }
  long long nb = (g.nwords + 1023) / 1024;
  if ((size_t)g.nwords > ctx->bitmap_cap_words) {
    free_dev(ctx->bitmap); free_dev(ctx->prefix); free_dev(ctx->blocksums); free_dev(ctx->nbitmap); free_dev(ctx->nprefix);
    HMSG_CUDA(cudaMalloc((void**)&ctx->bitmap, g.nwords * 4));
    HMSG_CUDA(cudaMalloc((void**)&ctx->prefix, g.nwords * 4));
    HMSG_CUDA(cudaMalloc((void**)&ctx->nbitmap, g.nwords * 4));
    HMSG_CUDA(cudaMalloc((void**)&ctx->nprefix, g.nwords * 4));
    HMSG_CUDA(cudaMalloc((void**)&ctx->blocksums, (nb + 1) * 4));
    ctx->bitmap_cap_words = (size_t)g.nwords;
  }
  HMSG_CUDA(cudaMemsetAsync(ctx->bitmap, 0, g.nwords * 4, ctx->stream));
  ctx->voxels_built = false; ctx->nodes_built = false; ctx->grid_set = true;
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxel_mark(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->grid_set) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_mark: call hmsg_voxel_grid_set first");
  if (!range_ok(ctx, frame_begin, n_frames)) return ctx->fail(HMSG_ERR_ARG, "hmsg_voxel_mark: bad frame range");
  ctx->prof_begin(PROF_GEOM);
  int32_t rc = for_frame_range(ctx, frame_begin, n_frames, [&](dim3 grid, int64_t f0) {
    k_mark<<<grid, TPB, 0, ctx->stream>>>(frame_args(ctx, f0), ctx->grid, ctx->bitmap);
  });
  ctx->prof_end(PROF_GEOM, (double)n_frames * ctx->cam.H * ctx->cam.W * 2.0);
  return rc;
}

__global__ void __launch_bounds__(TPB) k_bitmap_or(const uint32_t* __restrict__ gathered, int world, long long nwords, uint32_t* __restrict__ bitmap) {
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  uint32_t v = 0;
  for (int r = 0; r < world; r++) v |= gathered[(long long)r * nwords + w];
  bitmap[w] = v;
}

extern "C" int32_t hmsg_voxel_bitmap(hmsg_ctx* ctx, uint32_t** bitmap_dev, int64_t* nwords) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->grid_set) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_bitmap: call hmsg_voxel_grid_set first");
  if (bitmap_dev) *bitmap_dev = ctx->bitmap;
  if (nwords) *nwords = ctx->grid.nwords;
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxel_bitmap_or(hmsg_ctx* ctx, const uint32_t* gathered, int32_t world) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->grid_set || !gathered || world < 1) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_bitmap_or: bad state/argument");
  long long nw = ctx->grid.nwords;
  k_bitmap_or<<<(unsigned)((nw + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(gathered, world, nw, ctx->bitmap);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxel_scan(hmsg_ctx* ctx, int64_t* n_voxels) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->grid_set) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_scan: call hmsg_voxel_grid_set first");
  int32_t rc = run_scan(ctx, ctx->bitmap, ctx->grid.nwords, ctx->prefix, &ctx->n_voxels);
  if (rc) return rc;
  size_t nv = (size_t)std::max<int64_t>(ctx->n_voxels, 1);
  if (nv > ctx->voxel_cap) {
    free_dev(ctx->vox_acc); free_dev(ctx->vox_cnt); free_dev(ctx->vox_ijk); free_dev(ctx->rad_cnt);
    HMSG_CUDA(cudaMalloc((void**)&ctx->vox_acc, nv * 48));
    HMSG_CUDA(cudaMalloc((void**)&ctx->vox_cnt, nv * 4));
    HMSG_CUDA(cudaMalloc((void**)&ctx->vox_ijk, nv * 12));
    HMSG_CUDA(cudaMalloc((void**)&ctx->rad_cnt, nv * 4));
    ctx->voxel_cap = nv;
  }
  HMSG_CUDA(cudaMemsetAsync(ctx->vox_acc, 0, nv * 48, ctx->stream));
  HMSG_CUDA(cudaMemsetAsync(ctx->vox_cnt, 0, nv * 4, ctx->stream));
  if (n_voxels) *n_voxels = ctx->n_voxels;
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxel_accumulate(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->grid_set || !ctx->vox_acc) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_accumulate: call hmsg_voxel_scan first");
  if (!range_ok(ctx, frame_begin, n_frames)) return ctx->fail(HMSG_ERR_ARG, "hmsg_voxel_accumulate: bad frame range");
  ctx->prof_begin(PROF_GEOM);
  int32_t rc = for_frame_range(ctx, frame_begin, n_frames, [&](dim3 grid, int64_t f0) {
    k_accumulate<<<grid, TPB, 0, ctx->stream>>>(frame_args(ctx, f0), ctx->grid, ctx->bitmap, ctx->prefix, ctx->vox_acc, ctx->vox_cnt);
  });
  ctx->prof_end(PROF_GEOM, (double)n_frames * ctx->cam.H * ctx->cam.W * 5.0);
  return rc;
}

extern "C" int32_t hmsg_voxel_acc(hmsg_ctx* ctx, double** acc_dev, uint32_t** cnt_dev, int64_t* n_voxels) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->vox_acc) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_acc: call hmsg_voxel_scan first");
  if (acc_dev) *acc_dev = ctx->vox_acc;
  if (cnt_dev) *cnt_dev = ctx->vox_cnt;
  if (n_voxels) *n_voxels = ctx->n_voxels;
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxel_finalize(hmsg_ctx* ctx) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->grid_set || !ctx->vox_acc) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_finalize: call hmsg_voxel_scan first");
  size_t nv = (size_t)std::max<int64_t>(ctx->n_voxels, 1);
  k_finalize_voxels<<<(unsigned)((nv + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->vox_acc, ctx->vox_cnt, ctx->n_voxels);
  HMSG_LAUNCH_CHECK();
  k_write_ijk<<<(unsigned)((ctx->grid.nwords + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->bitmap, ctx->prefix, ctx->grid, ctx->vox_ijk);
  HMSG_LAUNCH_CHECK();
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->voxels_built = true; ctx->nodes_built = false;
  return HMSG_OK;
}

// single-GPU composition of the stages over all stored frames
extern "C" int32_t hmsg_voxel_build(hmsg_ctx* ctx, int64_t* n_voxels, double* min_bound_out) {
  if (!ctx) return HMSG_ERR_ARG;
  if (ctx->nframes <= 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxel_build: no frames");
  double mm[6];
  int32_t rc;
  if ((rc = hmsg_voxel_bounds(ctx, 0, ctx->nframes, mm))) return rc;       // graph.py:344-348: keys are relative to the GLOBAL min bound (H3)
  if ((rc = hmsg_voxel_grid_set(ctx, mm))) return rc;
  if ((rc = hmsg_voxel_mark(ctx, 0, ctx->nframes))) return rc;
  if ((rc = hmsg_voxel_scan(ctx, nullptr))) return rc;
  if ((rc = hmsg_voxel_accumulate(ctx, 0, ctx->nframes))) return rc;
  if ((rc = hmsg_voxel_finalize(ctx))) return rc;
  if (n_voxels) *n_voxels = ctx->n_voxels;
  if (min_bound_out) for (int k = 0; k < 3; k++) min_bound_out[k] = ctx->min_bound[k];
  return HMSG_OK;
}

// strided D2H helper: copy columns [c0,c0+3) of a [n,6] f64 table into a packed [n,3] host array
static int32_t read_cols(hmsg_ctx* ctx, const double* acc, int64_t n, int c0, double* out) {
  HMSG_CUDA(cudaMemcpy2DAsync(out, 24, acc + c0, 48, 24, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_voxels_read(hmsg_ctx* ctx, double* xyz, double* rgb, int32_t* ijk, uint32_t* count) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->voxels_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_voxels_read: call hmsg_voxel_build first");
  int64_t n = ctx->n_voxels;
  if (n == 0) return HMSG_OK;
  int32_t rc;
  if (xyz && (rc = read_cols(ctx, ctx->vox_acc, n, 0, xyz))) return rc;
  if (rgb && (rc = read_cols(ctx, ctx->vox_acc, n, 3, rgb))) return rc;
  if (ijk) HMSG_CUDA(cudaMemcpyAsync(ijk, ctx->vox_ijk, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (count) HMSG_CUDA(cudaMemcpyAsync(count, ctx->vox_cnt, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

// neighbour counts of the voxels [v_begin, v_begin + n): ranks can split the voxel table (hmsg_radius_filter_sharded)
int32_t geometry_radius_count(hmsg_ctx* ctx, double radius, int64_t v_begin, int64_t n) {
  if (!ctx->voxels_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_radius_filter: call hmsg_voxel_build first");
  if (!(radius > 0) || v_begin < 0 || n < 0 || v_begin + n > ctx->n_voxels) return ctx->fail(HMSG_ERR_ARG, "hmsg_radius_filter: bad argument");
  if (n == 0) return HMSG_OK;
  ctx->prof_begin(PROF_GEOM);
  long long threads = n * 32;
  k_radius_count<<<(unsigned)((threads + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->grid, ctx->bitmap, ctx->prefix, ctx->vox_acc, ctx->vox_ijk, v_begin,
                                                                                  v_begin + n, radius, ctx->rad_cnt);
  ctx->prof_end(PROF_GEOM, 0.0);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

// keep iff count > nb_points (self included, H5) -> node bitmap / prefix / coarse bitmap / compact node table
int32_t geometry_radius_finish(hmsg_ctx* ctx, int32_t nb_points, int64_t* n_nodes) {
  if (!ctx->voxels_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_radius_filter: call hmsg_voxel_build first");
  if (nb_points < 0) return ctx->fail(HMSG_ERR_ARG, "hmsg_radius_filter: bad argument");
  const GridDesc& g = ctx->grid;
  k_node_bitmap<<<(unsigned)((g.nwords + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->bitmap, ctx->prefix, g.nwords, ctx->rad_cnt,
                                                                                 (uint32_t)nb_points, ctx->nbitmap);
  HMSG_LAUNCH_CHECK();
  int32_t rc = run_scan(ctx, ctx->nbitmap, g.nwords, ctx->nprefix, &ctx->n_nodes);
  if (rc) return rc;
  HMSG_CUDA(cudaMemsetAsync(ctx->cbitmap, 0, (size_t)std::max<long long>(g.cnwords, 1) * 4, ctx->stream));
  k_coarse_mark<<<(unsigned)((g.nwords + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->nbitmap, g, ctx->cbitmap);
  HMSG_LAUNCH_CHECK();
  size_t nn = (size_t)std::max<int64_t>(ctx->n_nodes, 1);
  if (nn > ctx->node_cap) {
    free_dev(ctx->node_xyz); free_dev(ctx->node_rgb); free_dev(ctx->node_ijk); free_dev(ctx->node_vox);
    HMSG_CUDA(cudaMalloc((void**)&ctx->node_xyz, nn * 24));
    HMSG_CUDA(cudaMalloc((void**)&ctx->node_rgb, nn * 24));
    HMSG_CUDA(cudaMalloc((void**)&ctx->node_ijk, nn * 12));
    HMSG_CUDA(cudaMalloc((void**)&ctx->node_vox, nn * 8));
    ctx->node_cap = nn;
  }
  k_compact_nodes<<<(unsigned)((g.nwords + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(ctx->bitmap, ctx->prefix, ctx->nbitmap, ctx->nprefix, g.nwords,
                                                                                   ctx->vox_acc, ctx->vox_ijk, ctx->node_xyz, ctx->node_rgb,
                                                                                   ctx->node_ijk, ctx->node_vox);
  HMSG_LAUNCH_CHECK();
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->nodes_built = true;
  // feature state refers to node indices: invalidate
  ctx->d = 0;   // features_begin must be called again (buffers are kept and re-zeroed there)
  if (n_nodes) *n_nodes = ctx->n_nodes;
  return HMSG_OK;
}

extern "C" int32_t hmsg_radius_filter(hmsg_ctx* ctx, int32_t nb_points, double radius, int64_t* n_nodes) {
  if (!ctx) return HMSG_ERR_ARG;
  // pcd_denoise_dbscan(eps=0.01,min_points=100) (graph.py:352) is the identity for voxel_size >= 0.02
  int32_t rc = geometry_radius_count(ctx, radius, 0, ctx->n_voxels);
  if (rc) return rc;
  return geometry_radius_finish(ctx, nb_points, n_nodes);
}

extern "C" int32_t hmsg_radius_counts_read(hmsg_ctx* ctx, uint32_t* counts) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_radius_counts_read: call hmsg_radius_filter first");
  if (ctx->n_voxels) HMSG_CUDA(cudaMemcpy(counts, ctx->rad_cnt, ctx->n_voxels * 4, cudaMemcpyDeviceToHost));
  return HMSG_OK;
}

extern "C" int32_t hmsg_nodes_read(hmsg_ctx* ctx, double* xyz, double* rgb, int32_t* ijk, int64_t* voxel_index) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_nodes_read: call hmsg_radius_filter first");
  int64_t n = ctx->n_nodes;
  if (n == 0) return HMSG_OK;
  if (xyz) HMSG_CUDA(cudaMemcpyAsync(xyz, ctx->node_xyz, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (rgb) HMSG_CUDA(cudaMemcpyAsync(rgb, ctx->node_rgb, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (ijk) HMSG_CUDA(cudaMemcpyAsync(ijk, ctx->node_ijk, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  if (voxel_index) HMSG_CUDA(cudaMemcpyAsync(voxel_index, ctx->node_vox, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" int64_t hmsg_num_nodes(const hmsg_ctx* ctx) { return (ctx && ctx->nodes_built) ? ctx->n_nodes : -1; }

extern "C" int32_t hmsg_pixel_to_node(hmsg_ctx* ctx, int64_t frame, int64_t* idx, double* dist) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_pixel_to_node: call hmsg_radius_filter first");
  if (frame < 0 || frame >= ctx->nframes || !idx) return ctx->fail(HMSG_ERR_ARG, "hmsg_pixel_to_node: bad argument");
  if (ctx->n_nodes == 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_pixel_to_node: node table is empty");
  size_t hw = (size_t)ctx->cam.H * ctx->cam.W;
  int32_t rc = ctx->reserve((char**)&ctx->scratch, &ctx->scratch_bytes, hw * 16);
  if (rc) return rc;
  int64_t* di = (int64_t*)ctx->scratch; double* dd = (double*)(di + hw);
  ctx->wait_frames(frame, 1);
  k_pixel_to_node<<<(unsigned)((hw + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(frame_args(ctx, frame), ctx->grid, ctx->nbitmap, ctx->nprefix,
                                                                              ctx->cbitmap, ctx->node_xyz, di, dist ? dd : nullptr);
  HMSG_LAUNCH_CHECK();
  HMSG_CUDA(cudaMemcpyAsync(idx, di, hw * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (dist) HMSG_CUDA(cudaMemcpyAsync(dist, dd, hw * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_points_to_node(hmsg_ctx* ctx, const double* xyz, int64_t n, int64_t* idx, double* dist) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_points_to_node: call hmsg_radius_filter first");
  if (!xyz || !idx || n < 0) return ctx->fail(HMSG_ERR_ARG, "hmsg_points_to_node: bad argument");
  if (n == 0) return HMSG_OK;
  if (ctx->n_nodes == 0) return ctx->fail(HMSG_ERR_STATE, "hmsg_points_to_node: node table is empty");
  int32_t rc = ctx->reserve((char**)&ctx->scratch, &ctx->scratch_bytes, (size_t)n * 40);
  if (rc) return rc;
  double* dp = (double*)ctx->scratch; int64_t* di = (int64_t*)(dp + n * 3); double* dd = (double*)(di + n);
  HMSG_CUDA(cudaMemcpyAsync(dp, xyz, n * 24, cudaMemcpyHostToDevice, ctx->stream));
  k_points_to_node<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(dp, n, ctx->grid, ctx->nbitmap, ctx->nprefix, ctx->cbitmap,
                                                                             ctx->node_xyz, di, dist ? dd : nullptr);
  HMSG_LAUNCH_CHECK();
  HMSG_CUDA(cudaMemcpyAsync(idx, di, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (dist) HMSG_CUDA(cudaMemcpyAsync(dist, dd, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

// used by objects.cu (N2): device points -> nearest node index + euclidean distance, all on device
int32_t geometry_points_to_node_dev(hmsg_ctx* ctx, const double* d_pts, long long n, int64_t* d_idx, double* d_dist) {
  if (!ctx->nodes_built || ctx->n_nodes == 0) return ctx->fail(HMSG_ERR_STATE, "points_to_node: node table missing or empty");
  if (n <= 0) return HMSG_OK;
  k_points_to_node<<<(unsigned)((n + TPB - 1) / TPB), TPB, 0, ctx->stream>>>(d_pts, n, ctx->grid, ctx->nbitmap, ctx->nprefix, ctx->cbitmap, ctx->node_xyz,
                                                                             d_idx, d_dist);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}

// used by features.cu
int32_t geometry_nn_winner(hmsg_ctx* ctx, int64_t frame_begin, int n_frames) {
  int HW = ctx->cam.H * ctx->cam.W;
  dim3 grid((HW + TPB - 1) / TPB, n_frames);
  int32_t rc = ctx->reserve(&ctx->far_list, &ctx->far_list_bytes, (size_t)n_frames * HW * 4);
  if (rc) return rc;
  if (!ctx->far_count) HMSG_CUDA(cudaMalloc((void**)&ctx->far_count, 4));
  ctx->wait_frames(frame_begin, n_frames);
  HMSG_CUDA(cudaMemsetAsync(ctx->far_count, 0, 4, ctx->stream));
  ctx->prof_begin(PROF_NN);
  k_nn_winner<<<grid, TPB, 0, ctx->stream>>>(frame_args(ctx, frame_begin), ctx->grid, ctx->nbitmap, ctx->nprefix, ctx->node_xyz, ctx->pix_idx,
                                             ctx->win, ctx->n_nodes, ctx->epoch, ctx->far_list, ctx->far_count);
  HMSG_LAUNCH_CHECK();
  k_nn_far<<<ctx->sm_count * 4, TPB, 0, ctx->stream>>>(frame_args(ctx, frame_begin), ctx->grid, ctx->nbitmap, ctx->nprefix, ctx->cbitmap, ctx->node_xyz,
                                                       ctx->pix_idx, ctx->win, ctx->n_nodes, ctx->epoch, ctx->far_list, ctx->far_count);
  ctx->prof_end(PROF_NN, (double)HW * n_frames * 14.0);
  HMSG_LAUNCH_CHECK();
  return HMSG_OK;
}
