// objects.cu - N1: 3-D mask (point-set) merging across frames -> object instances.
// Reference: fsr_vln/memory/hmsg/graph/graph.py:424-448 (seq_merge call + small-mask removal) and
// fsr_vln/memory/hmsg/utils/graph_utils.py:
//   :1015-1038 seq_merge            global = frame0 masks; for every further frame
//                                   global = merge_3d_masks(global + frame masks); one more merge at the end
//   :919-956   merge_3d_masks       AABB-IoU gate -> overlap ratio -> connected components -> concat -> DBSCAN denoise
//   :883-916   compute_3d_bbox_iou  float64
//   :620-662   find_overlapping_ratio_faiss   float32 exact-L2 nearest neighbour both ways, count D < (1.5 r)^2
//   :665-679   merge_point_clouds_list        `+=` concat in list order, pcd_denoise_dbscan(eps 0.1, min_points 10)
//   :827-880   pcd_denoise_dbscan   Open3D ClusterDBSCAN, keep the largest cluster unless it has < 5 points
//
// The reference does this with O(F * masks^2) Python loops around faiss / Open3D calls.  Here every
// mask of the current list lives in one device pool (float64 xyz + rgb, ragged offsets); one merge step is
//   AABBs -> all-pairs IoU gate -> (mask, cell) sort + hash of the float32 points -> one block per gated
//   pair direction counting points with a neighbour in the other mask -> lock-free union-find over the
//   edges -> concat by component -> (component, cell) sort + hash of the float64 points -> DBSCAN
//   (neighbour counts, union-find over core points, border = lowest cluster) -> largest cluster -> compaction.
// "NN distance < r^2" is evaluated as "some point of the other mask within r" with exactly the float32
// arithmetic of an exact-L2 scan ((dx*dx + dy*dy) + dz*dz, individually rounded), so the counts equal the
// brute-force ones; DBSCAN uses the float64 d2 < eps^2 rule of nanoflann (self included).  Only ragged
// metadata (offsets, component roots: a few ints per mask) crosses to the host between the two phases.
#include "common.cuh"
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#define OTPB 256
static const unsigned long long EMPTY_KEY = ~0ull;

struct ObjState {
  double th = 0.75, radius = 0.05, iou_thresh = 0.05;
  int frames_added = 0;
  bool finished = false;
  // current global list: pool A
  double* a_xyz = nullptr; size_t a_xyz_bytes = 0;
  double* a_rgb = nullptr; size_t a_rgb_bytes = 0;
  std::vector<long long> a_off{0};
  // work list / concat pool B, result staging C
  double* b_xyz = nullptr; size_t b_xyz_bytes = 0;
  double* b_rgb = nullptr; size_t b_rgb_bytes = 0;
  double* c_xyz = nullptr; size_t c_xyz_bytes = 0;
  double* c_rgb = nullptr; size_t c_rgb_bytes = 0;
  long long* d_off = nullptr; size_t d_off_bytes = 0;       // offsets of the list being merged [n+1]
  long long* d_coff = nullptr; size_t d_coff_bytes = 0;     // component offsets [nc+1]
  long long* d_dst = nullptr; size_t d_dst_bytes = 0;       // destination offset of every mask in the concat pool
  int* d_comp_of = nullptr; size_t d_comp_of_bytes = 0;     // component id of every mask
  double* d_lo = nullptr; size_t d_lo_bytes = 0;            // [n,3] / [n,3] AABBs, then [6] global
  double* d_hi = nullptr; size_t d_hi_bytes = 0;
  double* d_glob = nullptr;                                 // 3 doubles: pool minimum
  int2* d_pairs = nullptr; size_t d_pairs_bytes = 0;
  int* d_pair_cnt = nullptr; size_t d_pair_cnt_bytes = 0;   // [2*npairs] hit counts
  int* d_counters = nullptr;                                // [4]: npairs, error flag
  int* d_parent = nullptr; size_t d_parent_bytes = 0;       // mask-level union-find [n]
  // per-point scratch
  unsigned long long* keys = nullptr; size_t keys_bytes = 0;        // 2 x npts (in / out)
  int* pidx = nullptr; size_t pidx_bytes = 0;                        // 2 x npts
  float4* spts = nullptr; size_t spts_bytes = 0;                     // sorted float32 points (+ original index)
  unsigned char* sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
  unsigned long long* hkeys = nullptr; size_t hkeys_bytes = 0;      // hash: key -> first sorted position
  int* hvals = nullptr; size_t hvals_bytes = 0;
  int* pt_comp = nullptr; size_t pt_comp_bytes = 0;                  // component of every concat point
  int* uf = nullptr; size_t uf_bytes = 0;                            // point-level union-find
  unsigned char* core = nullptr; size_t core_bytes = 0;
  int* label = nullptr; size_t label_bytes = 0;
  int* csize = nullptr; size_t csize_bytes = 0;
  int* cfirst = nullptr; size_t cfirst_bytes = 0;
  unsigned long long* best = nullptr; size_t best_bytes = 0;        // per component
  int* keep = nullptr; size_t keep_bytes = 0;
  int* keep_scan = nullptr; size_t keep_scan_bytes = 0;
  // host staging of frame masks
  double* in_xyz = nullptr; size_t in_xyz_bytes = 0;
  double* in_rgb = nullptr; size_t in_rgb_bytes = 0;
  long long stat_pairs = 0, stat_steps = 0;
};

// ------------------------------------------------------------------------------------------------
// lock-free union-find: larger root is hooked under the smaller one, so a root is the minimum index
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
  while (true) {
    int p = ((volatile int*)parent)[x];
    if (p == x) return x;
    int gp = ((volatile int*)parent)[p];
    if (gp != p) parent[x] = gp;     // path halving: always an ancestor, benign race
    x = p;
  }
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}

__device__ __forceinline__ unsigned long long hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}
__device__ __forceinline__ int hash_lookup(const unsigned long long* __restrict__ hk, const int* __restrict__ hv, unsigned long long hmask,
                                           unsigned long long key) {
  unsigned long long h = hash64(key) & hmask;
  while (true) {
    unsigned long long k = hk[h];
    if (k == key) return hv[h];
    if (k == EMPTY_KEY) return -1;
    h = (h + 1) & hmask;
  }
}

__device__ __forceinline__ int mask_of_point(const long long* __restrict__ off, int n, long long p) {   // off[m] <= p < off[m+1]
  int lo = 0, hi = n;
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= p) lo = mid; else hi = mid; }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// phase 1: AABBs, gate, overlap, components
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mask_aabb(const double* __restrict__ xyz, const long long* __restrict__ off, double* __restrict__ lo,
                                                   double* __restrict__ hi) {
  int m = blockIdx.x;
  long long b = off[m], e = off[m + 1];
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long long p = b + threadIdx.x; p < e; p += blockDim.x)
    for (int a = 0; a < 3; a++) { double v = xyz[p * 3 + a]; mn[a] = fmin(mn[a], v); mx[a] = fmax(mx[a], v); }
  __shared__ double s[2][3][128];
  for (int a = 0; a < 3; a++) { s[0][a][threadIdx.x] = mn[a]; s[1][a][threadIdx.x] = mx[a]; }
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int a = 0; a < 3; a++) {
        s[0][a][threadIdx.x] = fmin(s[0][a][threadIdx.x], s[0][a][threadIdx.x + o]);
        s[1][a][threadIdx.x] = fmax(s[1][a][threadIdx.x], s[1][a][threadIdx.x + o]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 3) {
    bool empty = e <= b;                              // Open3D: bounds of an empty cloud are (0,0,0)
    lo[m * 3 + threadIdx.x] = empty ? 0.0 : s[0][threadIdx.x][0];
    hi[m * 3 + threadIdx.x] = empty ? 0.0 : s[1][threadIdx.x][0];
  }
}

__global__ void k_global_min(const double* __restrict__ lo, const long long* __restrict__ off, int n, double* __restrict__ glob) {
  __shared__ double s[3][256];
  double mn[3] = {INFINITY, INFINITY, INFINITY};
  for (int m = threadIdx.x; m < n; m += blockDim.x)
    if (off[m + 1] > off[m]) for (int a = 0; a < 3; a++) mn[a] = fmin(mn[a], lo[m * 3 + a]);
  for (int a = 0; a < 3; a++) s[a][threadIdx.x] = mn[a];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) for (int a = 0; a < 3; a++) s[a][threadIdx.x] = fmin(s[a][threadIdx.x], s[a][threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x < 3) glob[threadIdx.x] = s[threadIdx.x][0] - 1.0;   // margin: float32-rounded copies of the minimum stay above the base
}

// compute_3d_bbox_iou(aa_bb[i], aa_bb[j]) > iou_thresh, i < j  (graph_utils.py:939-942)
__global__ void __launch_bounds__(OTPB) k_gate_pairs(const double* __restrict__ lo, const double* __restrict__ hi, int n, double iou_thresh,
                                                     int2* __restrict__ pairs, int* __restrict__ counters) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * n) return;
  int i = (int)(t / n), j = (int)(t % n);
  if (i >= j) return;
  double sz[3], e1[3], e2[3];
  for (int a = 0; a < 3; a++) {
    double omin = fmax(lo[i * 3 + a], lo[j * 3 + a]);
    double omax = fmin(hi[i * 3 + a], hi[j * 3 + a]);
    sz[a] = fmax(__dsub_rn(omax, omin), 0.0);
    e1[a] = __dsub_rn(hi[i * 3 + a], lo[i * 3 + a]);
    e2[a] = __dsub_rn(hi[j * 3 + a], lo[j * 3 + a]);
  }
  double ov = __dmul_rn(__dmul_rn(sz[0], sz[1]), sz[2]);
  double v1 = __dmul_rn(__dmul_rn(e1[0], e1[1]), e1[2]);
  double v2 = __dmul_rn(__dmul_rn(e2[0], e2[1]), e2[2]);
  double iou = __ddiv_rn(ov, __dsub_rn(__dadd_rn(v1, v2), ov));
  if (iou > iou_thresh) {                              // NaN (0/0) compares false, as in numpy
    int slot = atomicAdd(&counters[0], 1);
    pairs[slot] = make_int2(i, j);
  }
}

// sort key of a point: (mask or component) << 42 | cx << 28 | cy << 14 | cz, cells of size h relative to the pool minimum
template <bool F32>
__global__ void __launch_bounds__(OTPB) k_point_keys(const double* __restrict__ xyz, long long npts, const long long* __restrict__ off, int n,
                                                     const int* __restrict__ group_of_point, const double* __restrict__ glob, double h,
                                                     unsigned long long* __restrict__ keys, int* __restrict__ idx, int* __restrict__ counters) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int g = group_of_point ? group_of_point[p] : mask_of_point(off, n, p);
  unsigned long long key = (unsigned long long)g << 42;
  for (int a = 0; a < 3; a++) {
    double v = xyz[p * 3 + a];
    if (F32) v = (double)(float)v;
    long long c = (long long)floor((v - glob[a]) / h);
    if (c < 0 || c > 16380) { atomicExch(&counters[1], 1); c = 0; }
    key |= (unsigned long long)(c + 1) << (28 - 14 * a);         // +1: neighbour cell -1 stays non-negative
  }
  keys[p] = key;
  idx[p] = (int)p;
}

__global__ void __launch_bounds__(OTPB) k_hash_clear(unsigned long long* __restrict__ hk, unsigned long long size) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < size) hk[i] = EMPTY_KEY;
}

// run heads of the sorted key array -> hash(key) = first position; also gathers the float32 points
__global__ void __launch_bounds__(OTPB) k_hash_build(const unsigned long long* __restrict__ skeys, const int* __restrict__ sidx, long long npts,
                                                     const double* __restrict__ xyz, float4* __restrict__ spts, unsigned long long* __restrict__ hk,
                                                     int* __restrict__ hv, unsigned long long hmask) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  unsigned long long key = skeys[p];
  if (spts) {
    int o = sidx[p];
    spts[p] = make_float4((float)xyz[(long long)o * 3], (float)xyz[(long long)o * 3 + 1], (float)xyz[(long long)o * 3 + 2], __int_as_float(o));
  }
  if (p == 0 || skeys[p - 1] != key) {
    unsigned long long h = hash64(key) & hmask;
    while (true) {
      unsigned long long old = atomicCAS(&hk[h], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { hv[h] = (int)p; break; }
      h = (h + 1) & hmask;
    }
  }
}

// one block per (gated pair, direction): how many points of src have a point of dst with float32 d2 < r2
__global__ void __launch_bounds__(128) k_overlap(const int2* __restrict__ pairs, const long long* __restrict__ off, const double* __restrict__ xyz,
                                                 const double* __restrict__ glob, double h, float r2, const unsigned long long* __restrict__ skeys,
                                                 const float4* __restrict__ spts, const unsigned long long* __restrict__ hk,
                                                 const int* __restrict__ hv, unsigned long long hmask, long long npts, int* __restrict__ out) {
  int pr = blockIdx.x >> 1, dir = blockIdx.x & 1;
  int2 ij = pairs[pr];
  int src = dir ? ij.y : ij.x, dst = dir ? ij.x : ij.y;
  long long b = off[src], e = off[src + 1];
  int cnt = 0;
  if (off[dst + 1] > off[dst]) {
    for (long long p = b + threadIdx.x; p < e; p += blockDim.x) {
      float x = (float)xyz[p * 3], y = (float)xyz[p * 3 + 1], z = (float)xyz[p * 3 + 2];
      long long c[3];
      c[0] = (long long)floor(((double)x - glob[0]) / h) + 1;
      c[1] = (long long)floor(((double)y - glob[1]) / h) + 1;
      c[2] = (long long)floor(((double)z - glob[2]) / h) + 1;
      bool hit = false;
      for (int dx = -1; dx <= 1 && !hit; dx++)
        for (int dy = -1; dy <= 1 && !hit; dy++)
          for (int dz = -1; dz <= 1 && !hit; dz++) {
            unsigned long long key = ((unsigned long long)dst << 42) | ((unsigned long long)(c[0] + dx) << 28) | ((unsigned long long)(c[1] + dy) << 14) |
                                     (unsigned long long)(c[2] + dz);
            int s = hash_lookup(hk, hv, hmask, key);
            if (s < 0) continue;
            for (long long q = s; q < npts && skeys[q] == key; q++) {
              float4 o = spts[q];
              float ex = __fsub_rn(x, o.x), ey = __fsub_rn(y, o.y), ez = __fsub_rn(z, o.z);
              float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
              if (d2 < r2) { hit = true; break; }
            }
          }
      cnt += hit ? 1 : 0;
    }
  }
  __shared__ int s[128];
  s[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

__global__ void k_iota(int* __restrict__ a, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int)i;
}

// overlap_matrix[i, j] > overlap_threshold -> union(i, j)   (graph_utils.py:947-948)
__global__ void __launch_bounds__(OTPB) k_mask_edges(const int2* __restrict__ pairs, const int* __restrict__ cnt, int npairs,
                                                     const long long* __restrict__ off, double th, int* __restrict__ parent) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npairs) return;
  int2 ij = pairs[t];
  long long n1 = off[ij.x + 1] - off[ij.x], n2 = off[ij.y + 1] - off[ij.y];
  if (n1 == 0 || n2 == 0) return;                      // find_overlapping_ratio_faiss returns 0
  double r1 = __ddiv_rn((double)cnt[2 * t], (double)n1), r2 = __ddiv_rn((double)cnt[2 * t + 1], (double)n2);
  if (fmax(r1, r2) > th) uf_union(parent, ij.x, ij.y);
}

__global__ void k_flatten(int* __restrict__ parent, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = uf_find(parent, (int)i);
}

// ------------------------------------------------------------------------------------------------
// phase 2: concat by component, DBSCAN, largest cluster, compaction
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OTPB) k_concat(const double* __restrict__ xyz, const double* __restrict__ rgb, long long npts,
                                                 const long long* __restrict__ off, int n, const long long* __restrict__ dst,
                                                 const int* __restrict__ comp_of, double* __restrict__ oxyz, double* __restrict__ orgb,
                                                 int* __restrict__ pt_comp) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int m = mask_of_point(off, n, p);
  long long q = dst[m] + (p - off[m]);
  for (int a = 0; a < 3; a++) { oxyz[q * 3 + a] = xyz[p * 3 + a]; orgb[q * 3 + a] = rgb[p * 3 + a]; }
  pt_comp[q] = comp_of[m];
}

// neighbour scan of point p (original index) over the 27 cells of its component; F(q, d2) for every q with d2 < eps2 (self included)
template <typename Fn>
__device__ __forceinline__ void db_neighbours(long long p, int comp, const double* __restrict__ xyz, const double* __restrict__ glob, double h,
                                              double eps2, const unsigned long long* __restrict__ skeys, const int* __restrict__ sidx,
                                              const unsigned long long* __restrict__ hk, const int* __restrict__ hv, unsigned long long hmask,
                                              long long npts, Fn fn) {
  double x = xyz[p * 3], y = xyz[p * 3 + 1], z = xyz[p * 3 + 2];
  long long c0 = (long long)floor((x - glob[0]) / h) + 1, c1 = (long long)floor((y - glob[1]) / h) + 1, c2 = (long long)floor((z - glob[2]) / h) + 1;
  for (int dx = -1; dx <= 1; dx++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dz = -1; dz <= 1; dz++) {
        unsigned long long key = ((unsigned long long)comp << 42) | ((unsigned long long)(c0 + dx) << 28) | ((unsigned long long)(c1 + dy) << 14) |
                                 (unsigned long long)(c2 + dz);
        int s = hash_lookup(hk, hv, hmask, key);
        if (s < 0) continue;
        for (long long q = s; q < npts && skeys[q] == key; q++) {
          int o = sidx[q];
          double d2 = sqdist3(xyz[(long long)o * 3], xyz[(long long)o * 3 + 1], xyz[(long long)o * 3 + 2], x, y, z);
          if (d2 < eps2) fn(o);
        }
      }
}

#define DB_ARGS const double *__restrict__ xyz, const int *__restrict__ pt_comp, const double *__restrict__ glob, double h, double eps2, \
                const unsigned long long *__restrict__ skeys, const int *__restrict__ sidx, const unsigned long long *__restrict__ hk,      \
                const int *__restrict__ hv, unsigned long long hmask, long long npts
#define DB_PASS xyz, glob, h, eps2, skeys, sidx, hk, hv, hmask, npts

__global__ void __launch_bounds__(OTPB) k_db_core(DB_ARGS, int min_points, unsigned char* __restrict__ core) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  long long p = sidx[t];                                   // walk in sorted order: neighbouring threads share cells
  int cnt = 0;
  db_neighbours(p, pt_comp[p], DB_PASS, [&](int) { cnt++; });
  core[p] = cnt >= min_points ? 1 : 0;
}

__global__ void __launch_bounds__(OTPB) k_db_union(DB_ARGS, const unsigned char* __restrict__ core, int* __restrict__ uf) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  long long p = sidx[t];
  if (!core[p]) return;
  db_neighbours(p, pt_comp[p], DB_PASS, [&](int q) { if (q < p && core[q]) uf_union(uf, (int)p, q); });
}

// label = cluster root (lowest core index of the cluster); border points take the lowest-numbered cluster
// among their core neighbours (clusters are numbered by their lowest core index); noise = -1
__global__ void __launch_bounds__(OTPB) k_db_label(DB_ARGS, const unsigned char* __restrict__ core, int* __restrict__ uf, int* __restrict__ label,
                                                   int* __restrict__ csize, int* __restrict__ cfirst) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  long long p = sidx[t];
  int lab;
  if (core[p]) {
    lab = uf_find(uf, (int)p);
  } else {
    int best = 0x7fffffff;
    db_neighbours(p, pt_comp[p], DB_PASS, [&](int q) { if (core[q]) best = min(best, uf_find(uf, q)); });
    lab = best == 0x7fffffff ? -1 : best;
  }
  label[p] = lab;
  if (lab >= 0) { atomicAdd(&csize[lab], 1); atomicMin(&cfirst[lab], (int)p); }
}

// Counter(labels).most_common(1): largest size, ties -> the label met first in array order
__global__ void __launch_bounds__(OTPB) k_db_best(const int* __restrict__ label, const int* __restrict__ csize, const int* __restrict__ cfirst,
                                                  const int* __restrict__ pt_comp, const long long* __restrict__ coff, long long npts,
                                                  unsigned long long* __restrict__ best) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  if (label[p] != (int)p) return;                          // cluster roots only
  int c = pt_comp[p];
  unsigned int rel = (unsigned int)(cfirst[p] - coff[c]);
  unsigned long long v = ((unsigned long long)(unsigned int)csize[p] << 32) | (unsigned long long)(0xFFFFFFFFu - rel);
  atomicMax(&best[c], v);
}

__global__ void __launch_bounds__(OTPB) k_db_keep(const int* __restrict__ label, const int* __restrict__ pt_comp, const long long* __restrict__ coff,
                                                  const unsigned long long* __restrict__ best, long long npts, int* __restrict__ keep) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int c = pt_comp[p];
  unsigned long long b = best[c];
  int k = 1;
  if (b != 0 && (b >> 32) >= 5) {                         // a cluster exists and is not "too small": keep only it
    long long first = coff[c] + (long long)(0xFFFFFFFFu - (unsigned int)(b & 0xFFFFFFFFu));
    k = label[p] == label[first] ? 1 : 0;
  }
  keep[p] = k;
}

__global__ void __launch_bounds__(OTPB) k_compact(const double* __restrict__ xyz, const double* __restrict__ rgb, const int* __restrict__ keep,
                                                  const int* __restrict__ scan, long long npts, double* __restrict__ oxyz,
                                                  double* __restrict__ orgb) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts || !keep[p]) return;
  long long q = scan[p];
  for (int a = 0; a < 3; a++) { oxyz[q * 3 + a] = xyz[p * 3 + a]; orgb[q * 3 + a] = rgb[p * 3 + a]; }
}

__global__ void k_new_offsets(const int* __restrict__ scan, const int* __restrict__ keep, const long long* __restrict__ coff, int nc, long long npts,
                              long long* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nc) return;
  long long p = coff[c];
  out[c] = p < npts ? scan[p] : (npts ? scan[npts - 1] + keep[npts - 1] : 0);
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n, int tpb = OTPB) { return (unsigned)((n + tpb - 1) / tpb); }

static int32_t sort_and_hash(hmsg_ctx* ctx, ObjState* st, const double* xyz, long long npts, bool with_f32, unsigned long long* hmask_out) {
  int32_t rc;
  size_t tmp = 0;
  unsigned long long* kin = st->keys; unsigned long long* kout = st->keys + npts;
  int* iin = st->pidx; int* iout = st->pidx + npts;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, iin, iout, (int)npts, 0, 63, ctx->stream);
  if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
  HMSG_CUDA(cub::DeviceRadixSort::SortPairs(st->sort_tmp, tmp, kin, kout, iin, iout, (int)npts, 0, 63, ctx->stream));
  unsigned long long hsize = 1024;
  while (hsize < (unsigned long long)npts * 2) hsize <<= 1;
  if ((rc = ctx->reserve(&st->hkeys, &st->hkeys_bytes, hsize * 8))) return rc;
  if ((rc = ctx->reserve(&st->hvals, &st->hvals_bytes, hsize * 4))) return rc;
  k_hash_clear<<<blocks_for((long long)hsize), OTPB, 0, ctx->stream>>>(st->hkeys, hsize);
  if (with_f32 && (rc = ctx->reserve(&st->spts, &st->spts_bytes, (size_t)npts * 16))) return rc;
  k_hash_build<<<blocks_for(npts), OTPB, 0, ctx->stream>>>(kout, iout, npts, xyz, with_f32 ? st->spts : nullptr, st->hkeys, st->hvals, hsize - 1);
  HMSG_LAUNCH_CHECK();
  *hmask_out = hsize - 1;
  return HMSG_OK;
}

// merge_3d_masks over the list (xyz, rgb, off[n+1]) living in pool B; result -> pool A / st->a_off
static int32_t merge_list(hmsg_ctx* ctx, ObjState* st, const std::vector<long long>& off) {
  int32_t rc;
  int n = (int)off.size() - 1;
  long long npts = off[n];
  st->stat_steps++;
  if (n == 0) { st->a_off.assign(1, 0); return HMSG_OK; }
  if (n >= (1 << 21)) return ctx->fail(HMSG_ERR_CAPACITY, "objects: more than 2^21 masks in one merge");
  if (npts >= (1LL << 31)) return ctx->fail(HMSG_ERR_CAPACITY, "objects: more than 2^31 mask points in one merge");
  if ((rc = ctx->reserve(&st->d_off, &st->d_off_bytes, (size_t)(n + 1) * 8))) return rc;
  HMSG_CUDA(cudaMemcpyAsync(st->d_off, off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = ctx->reserve(&st->d_lo, &st->d_lo_bytes, (size_t)n * 24 + 64))) return rc;
  if ((rc = ctx->reserve(&st->d_hi, &st->d_hi_bytes, (size_t)n * 24))) return rc;
  if (!st->d_counters) { HMSG_CUDA(cudaMalloc((void**)&st->d_counters, 16)); HMSG_CUDA(cudaMalloc((void**)&st->d_glob, 24)); }
  HMSG_CUDA(cudaMemsetAsync(st->d_counters, 0, 16, ctx->stream));
  std::vector<int> parent(n);
  for (int i = 0; i < n; i++) parent[i] = i;
  int npairs = 0;
  if (n > 1 && npts > 0) {
    size_t maxpairs = (size_t)n * (n - 1) / 2;
    if ((rc = ctx->reserve(&st->d_pairs, &st->d_pairs_bytes, maxpairs * 8))) return rc;
    k_mask_aabb<<<n, 128, 0, ctx->stream>>>(st->b_xyz, st->d_off, st->d_lo, st->d_hi);
    k_global_min<<<1, 256, 0, ctx->stream>>>(st->d_lo, st->d_off, n, st->d_glob);
    k_gate_pairs<<<blocks_for((long long)n * n), OTPB, 0, ctx->stream>>>(st->d_lo, st->d_hi, n, st->iou_thresh, st->d_pairs, st->d_counters);
    HMSG_LAUNCH_CHECK();
    HMSG_CUDA(cudaMemcpyAsync(&npairs, st->d_counters, 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
    st->stat_pairs += npairs;
  }
  if (npairs > 0) {
    double r = 1.5 * st->radius;                       // graph_utils.py:943
    float r2 = (float)(r * r);                         // `D < radius**2` on a float32 array: compared in float32
    double h = r * 1.01;
    if ((rc = ctx->reserve(&st->keys, &st->keys_bytes, (size_t)npts * 16))) return rc;
    if ((rc = ctx->reserve(&st->pidx, &st->pidx_bytes, (size_t)npts * 8))) return rc;
    k_point_keys<true><<<blocks_for(npts), OTPB, 0, ctx->stream>>>(st->b_xyz, npts, st->d_off, n, nullptr, st->d_glob, h, st->keys, st->pidx, st->d_counters);
    unsigned long long hmask;
    if ((rc = sort_and_hash(ctx, st, st->b_xyz, npts, true, &hmask))) return rc;
    if ((rc = ctx->reserve(&st->d_pair_cnt, &st->d_pair_cnt_bytes, (size_t)npairs * 8))) return rc;
    k_overlap<<<2 * npairs, 128, 0, ctx->stream>>>(st->d_pairs, st->d_off, st->b_xyz, st->d_glob, h, r2, st->keys + npts, st->spts, st->hkeys, st->hvals,
                                                   hmask, npts, st->d_pair_cnt);
    if ((rc = ctx->reserve(&st->d_parent, &st->d_parent_bytes, (size_t)n * 4))) return rc;
    k_iota<<<blocks_for(n), OTPB, 0, ctx->stream>>>(st->d_parent, n);
    k_mask_edges<<<blocks_for(npairs), OTPB, 0, ctx->stream>>>(st->d_pairs, st->d_pair_cnt, npairs, st->d_off, st->th, st->d_parent);
    k_flatten<<<blocks_for(n), OTPB, 0, ctx->stream>>>(st->d_parent, n);
    HMSG_LAUNCH_CHECK();
    int flag[2];
    HMSG_CUDA(cudaMemcpyAsync(parent.data(), st->d_parent, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(flag, st->d_counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag[1]) return ctx->fail(HMSG_ERR_CAPACITY, "objects: mask points span more than 2^14 search cells per axis");
  }
  // components in order of their lowest mask index (scipy connected_components labelling), members in index order
  std::vector<int> comp_of(n), root_comp(n, -1);
  int nc = 0;
  for (int i = 0; i < n; i++) if (parent[i] == i) root_comp[i] = nc++;
  std::vector<long long> coff(nc + 1, 0), dst(n);
  for (int i = 0; i < n; i++) { comp_of[i] = root_comp[parent[i]]; coff[comp_of[i] + 1] += off[i + 1] - off[i]; }
  for (int c = 0; c < nc; c++) coff[c + 1] += coff[c];
  {
    std::vector<long long> fill(coff.begin(), coff.end() - 1);
    for (int i = 0; i < n; i++) { dst[i] = fill[comp_of[i]]; fill[comp_of[i]] += off[i + 1] - off[i]; }
  }
  if (npts == 0) { st->a_off.assign(nc + 1, 0); return HMSG_OK; }
  // concat into pool C
  if ((rc = ctx->reserve(&st->c_xyz, &st->c_xyz_bytes, (size_t)npts * 24))) return rc;
  if ((rc = ctx->reserve(&st->c_rgb, &st->c_rgb_bytes, (size_t)npts * 24))) return rc;
  if ((rc = ctx->reserve(&st->d_coff, &st->d_coff_bytes, (size_t)(nc + 2) * 8 * 2))) return rc;
  if ((rc = ctx->reserve(&st->d_dst, &st->d_dst_bytes, (size_t)n * 8))) return rc;
  if ((rc = ctx->reserve(&st->d_comp_of, &st->d_comp_of_bytes, (size_t)n * 4))) return rc;
  if ((rc = ctx->reserve(&st->pt_comp, &st->pt_comp_bytes, (size_t)npts * 4))) return rc;
  HMSG_CUDA(cudaMemcpyAsync(st->d_coff, coff.data(), (size_t)(nc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(st->d_dst, dst.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(st->d_comp_of, comp_of.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  k_concat<<<blocks_for(npts), OTPB, 0, ctx->stream>>>(st->b_xyz, st->b_rgb, npts, st->d_off, n, st->d_dst, st->d_comp_of, st->c_xyz, st->c_rgb, st->pt_comp);
  HMSG_LAUNCH_CHECK();
  // DBSCAN(eps 0.1, min_points 10) per component  (graph_utils.py:678)
  const double eps = 0.1; const int min_points = 10;
  double h = eps * 1.0001;
  if (npairs == 0) k_global_min<<<1, 256, 0, ctx->stream>>>(st->d_lo, st->d_off, n, st->d_glob);      // pool minimum (AABBs exist when n > 1)
  if (n == 1) { k_mask_aabb<<<1, 128, 0, ctx->stream>>>(st->b_xyz, st->d_off, st->d_lo, st->d_hi); k_global_min<<<1, 256, 0, ctx->stream>>>(st->d_lo, st->d_off, n, st->d_glob); }
  if ((rc = ctx->reserve(&st->keys, &st->keys_bytes, (size_t)npts * 16))) return rc;
  if ((rc = ctx->reserve(&st->pidx, &st->pidx_bytes, (size_t)npts * 8))) return rc;
  HMSG_CUDA(cudaMemsetAsync(st->d_counters, 0, 16, ctx->stream));
  k_point_keys<false><<<blocks_for(npts), OTPB, 0, ctx->stream>>>(st->c_xyz, npts, nullptr, 0, st->pt_comp, st->d_glob, h, st->keys, st->pidx, st->d_counters);
  unsigned long long hmask;
  if ((rc = sort_and_hash(ctx, st, st->c_xyz, npts, false, &hmask))) return rc;
  if ((rc = ctx->reserve(&st->core, &st->core_bytes, (size_t)npts))) return rc;
  if ((rc = ctx->reserve(&st->uf, &st->uf_bytes, (size_t)npts * 4))) return rc;
  if ((rc = ctx->reserve(&st->label, &st->label_bytes, (size_t)npts * 4))) return rc;
  if ((rc = ctx->reserve(&st->csize, &st->csize_bytes, (size_t)npts * 4))) return rc;
  if ((rc = ctx->reserve(&st->cfirst, &st->cfirst_bytes, (size_t)npts * 4))) return rc;
  if ((rc = ctx->reserve(&st->best, &st->best_bytes, (size_t)nc * 8))) return rc;
  if ((rc = ctx->reserve(&st->keep, &st->keep_bytes, (size_t)npts * 4))) return rc;
  if ((rc = ctx->reserve(&st->keep_scan, &st->keep_scan_bytes, (size_t)npts * 4))) return rc;
  const unsigned long long* skeys = st->keys + npts;
  const int* sidx = st->pidx + npts;
  unsigned gb = blocks_for(npts);
  k_db_core<<<gb, OTPB, 0, ctx->stream>>>(st->c_xyz, st->pt_comp, st->d_glob, h, eps * eps, skeys, sidx, st->hkeys, st->hvals, hmask, npts, min_points, st->core);
  k_iota<<<gb, OTPB, 0, ctx->stream>>>(st->uf, npts);
  k_db_union<<<gb, OTPB, 0, ctx->stream>>>(st->c_xyz, st->pt_comp, st->d_glob, h, eps * eps, skeys, sidx, st->hkeys, st->hvals, hmask, npts, st->core, st->uf);
  HMSG_CUDA(cudaMemsetAsync(st->csize, 0, (size_t)npts * 4, ctx->stream));
  HMSG_CUDA(cudaMemsetAsync(st->cfirst, 0x7f, (size_t)npts * 4, ctx->stream));
  HMSG_CUDA(cudaMemsetAsync(st->best, 0, (size_t)nc * 8, ctx->stream));
  k_db_label<<<gb, OTPB, 0, ctx->stream>>>(st->c_xyz, st->pt_comp, st->d_glob, h, eps * eps, skeys, sidx, st->hkeys, st->hvals, hmask, npts, st->core, st->uf,
                                          st->label, st->csize, st->cfirst);
  k_db_best<<<gb, OTPB, 0, ctx->stream>>>(st->label, st->csize, st->cfirst, st->pt_comp, st->d_coff, npts, st->best);
  k_db_keep<<<gb, OTPB, 0, ctx->stream>>>(st->label, st->pt_comp, st->d_coff, st->best, npts, st->keep);
  HMSG_LAUNCH_CHECK();
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, st->keep, st->keep_scan, (int)npts, ctx->stream);
  if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
  HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->sort_tmp, tmp, st->keep, st->keep_scan, (int)npts, ctx->stream));
  if ((rc = ctx->reserve(&st->a_xyz, &st->a_xyz_bytes, (size_t)npts * 24))) return rc;
  if ((rc = ctx->reserve(&st->a_rgb, &st->a_rgb_bytes, (size_t)npts * 24))) return rc;
  k_compact<<<gb, OTPB, 0, ctx->stream>>>(st->c_xyz, st->c_rgb, st->keep, st->keep_scan, npts, st->a_xyz, st->a_rgb);
  long long* d_newoff = st->d_coff + (nc + 2);
  k_new_offsets<<<blocks_for(nc + 1), OTPB, 0, ctx->stream>>>(st->keep_scan, st->keep, st->d_coff, nc, npts, d_newoff);
  HMSG_LAUNCH_CHECK();
  st->a_off.resize(nc + 1);
  int flag[2];
  HMSG_CUDA(cudaMemcpyAsync(st->a_off.data(), d_newoff, (size_t)(nc + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(flag, st->d_counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flag[1]) return ctx->fail(HMSG_ERR_CAPACITY, "objects: mask points span more than 2^14 DBSCAN cells per axis");
  return HMSG_OK;
}

int32_t objects_destroy(hmsg_ctx* ctx) {
  ObjState* st = ctx->obj;
  if (!st) return HMSG_OK;
  free_dev(st->a_xyz); free_dev(st->a_rgb); free_dev(st->b_xyz); free_dev(st->b_rgb); free_dev(st->c_xyz); free_dev(st->c_rgb);
  free_dev(st->d_off); free_dev(st->d_coff); free_dev(st->d_dst); free_dev(st->d_comp_of); free_dev(st->d_lo); free_dev(st->d_hi); free_dev(st->d_glob);
  free_dev(st->d_pairs); free_dev(st->d_pair_cnt); free_dev(st->d_counters); free_dev(st->d_parent); free_dev(st->keys); free_dev(st->pidx);
  free_dev(st->spts); free_dev(st->sort_tmp); free_dev(st->hkeys); free_dev(st->hvals); free_dev(st->pt_comp); free_dev(st->uf); free_dev(st->core);
  free_dev(st->label); free_dev(st->csize); free_dev(st->cfirst); free_dev(st->best); free_dev(st->keep); free_dev(st->keep_scan);
  free_dev(st->in_xyz); free_dev(st->in_rgb);
  delete st;
  ctx->obj = nullptr;
  return HMSG_OK;
}

extern "C" int32_t hmsg_objects_begin(hmsg_ctx* ctx, double overlap_thresh, double down_size, double iou_thresh) {
  if (!ctx) return HMSG_ERR_ARG;
  if (!(down_size > 0)) return ctx->fail(HMSG_ERR_ARG, "hmsg_objects_begin: down_size must be positive");
  if (!ctx->obj) ctx->obj = new ObjState();
  ObjState* st = ctx->obj;
  st->th = overlap_thresh; st->radius = down_size; st->iou_thresh = iou_thresh;
  st->frames_added = 0; st->finished = false;
  st->a_off.assign(1, 0);
  st->stat_pairs = st->stat_steps = 0;
  return HMSG_OK;
}

// global + frame masks -> pool B (global first, list order of `global_masks + frames_pcd[i]`)
static int32_t stage_list(hmsg_ctx* ctx, ObjState* st, int32_t n_masks, const int64_t* offsets, const double* xyz, const double* rgb, int on_device,
                          std::vector<long long>& off) {
  int32_t rc;
  long long g_pts = st->a_off.back(), f_pts = n_masks ? offsets[n_masks] : 0;
  if ((rc = ctx->reserve(&st->b_xyz, &st->b_xyz_bytes, (size_t)std::max<long long>(g_pts + f_pts, 1) * 24))) return rc;
  if ((rc = ctx->reserve(&st->b_rgb, &st->b_rgb_bytes, (size_t)std::max<long long>(g_pts + f_pts, 1) * 24))) return rc;
  if (g_pts) {
    HMSG_CUDA(cudaMemcpyAsync(st->b_xyz, st->a_xyz, (size_t)g_pts * 24, cudaMemcpyDeviceToDevice, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(st->b_rgb, st->a_rgb, (size_t)g_pts * 24, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (f_pts) {
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    HMSG_CUDA(cudaMemcpyAsync(st->b_xyz + g_pts * 3, xyz, (size_t)f_pts * 24, kind, ctx->stream));
    if (rgb) HMSG_CUDA(cudaMemcpyAsync(st->b_rgb + g_pts * 3, rgb, (size_t)f_pts * 24, kind, ctx->stream));
    else HMSG_CUDA(cudaMemsetAsync(st->b_rgb + g_pts * 3, 0, (size_t)f_pts * 24, ctx->stream));
  }
  off = st->a_off;
  for (int m = 0; m < n_masks; m++) off.push_back(g_pts + offsets[m + 1]);
  return HMSG_OK;
}

extern "C" int32_t hmsg_objects_add_masks(hmsg_ctx* ctx, int32_t n_masks, const int64_t* offsets, const double* xyz, const double* rgb,
                                          int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  ObjState* st = ctx->obj;
  if (!st || st->finished) return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_add_masks: call hmsg_objects_begin first");
  if (n_masks < 0 || (n_masks > 0 && (!offsets || offsets[0] != 0)) || (n_masks > 0 && offsets[n_masks] > 0 && !xyz))
    return ctx->fail(HMSG_ERR_ARG, "hmsg_objects_add_masks: bad argument (offsets[0] must be 0; offsets are host memory)");
  std::vector<long long> off;
  int32_t rc = stage_list(ctx, st, n_masks, offsets, xyz, rgb, on_device, off);
  if (rc) return rc;
  if (st->frames_added == 0) {          // global_masks = frames_pcd[0]: taken as is (graph_utils.py:1021)
    std::swap(st->a_xyz, st->b_xyz); std::swap(st->a_xyz_bytes, st->b_xyz_bytes);
    std::swap(st->a_rgb, st->b_rgb); std::swap(st->a_rgb_bytes, st->b_rgb_bytes);
    st->a_off = off;
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));     // host staging buffers may be reused by the caller
  } else {
    if ((rc = merge_list(ctx, st, off))) return rc;
  }
  st->frames_added++;
  return HMSG_OK;
}

extern "C" int32_t hmsg_objects_finish(hmsg_ctx* ctx, int32_t min_points, int64_t* n_objects, int64_t* n_points) {
  if (!ctx) return HMSG_ERR_ARG;
  ObjState* st = ctx->obj;
  if (!st || st->finished) return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_finish: call hmsg_objects_begin / add first");
  std::vector<long long> off;
  int32_t rc = stage_list(ctx, st, 0, nullptr, nullptr, nullptr, 1, off);       // "apply one more merge" (graph_utils.py:1032-1037)
  if (rc) return rc;
  if ((rc = merge_list(ctx, st, off))) return rc;
  // graph.py:444-448: drop masks that are empty or have < min_points points (list order kept)
  int n = (int)st->a_off.size() - 1;
  std::vector<long long> noff{0};
  long long npts = st->a_off.back();
  if ((rc = ctx->reserve(&st->b_xyz, &st->b_xyz_bytes, (size_t)std::max<long long>(npts, 1) * 24))) return rc;
  if ((rc = ctx->reserve(&st->b_rgb, &st->b_rgb_bytes, (size_t)std::max<long long>(npts, 1) * 24))) return rc;
  for (int m = 0; m < n; m++) {
    long long c = st->a_off[m + 1] - st->a_off[m];
    if (c == 0 || c < min_points) continue;
    HMSG_CUDA(cudaMemcpyAsync(st->b_xyz + noff.back() * 3, st->a_xyz + st->a_off[m] * 3, (size_t)c * 24, cudaMemcpyDeviceToDevice, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(st->b_rgb + noff.back() * 3, st->a_rgb + st->a_off[m] * 3, (size_t)c * 24, cudaMemcpyDeviceToDevice, ctx->stream));
    noff.push_back(noff.back() + c);
  }
  std::swap(st->a_xyz, st->b_xyz); std::swap(st->a_xyz_bytes, st->b_xyz_bytes);
  std::swap(st->a_rgb, st->b_rgb); std::swap(st->a_rgb_bytes, st->b_rgb_bytes);
  st->a_off = noff;
  st->finished = true;
  if (n_objects) *n_objects = (int64_t)noff.size() - 1;
  if (n_points) *n_points = noff.back();
  return HMSG_OK;
}

extern "C" int32_t hmsg_objects_read(hmsg_ctx* ctx, int64_t* offsets, double* xyz, double* rgb) {
  if (!ctx) return HMSG_ERR_ARG;
  ObjState* st = ctx->obj;
  if (!st) return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_read: no object list");
  int n = (int)st->a_off.size() - 1;
  if (offsets) for (int m = 0; m <= n; m++) offsets[m] = st->a_off[m];
  long long npts = st->a_off.back();
  if (xyz && npts) HMSG_CUDA(cudaMemcpyAsync(xyz, st->a_xyz, (size_t)npts * 24, cudaMemcpyDeviceToHost, ctx->stream));
  if (rgb && npts) HMSG_CUDA(cudaMemcpyAsync(rgb, st->a_rgb, (size_t)npts * 24, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  return HMSG_OK;
}

extern "C" int32_t hmsg_objects_count(hmsg_ctx* ctx, int64_t* n_masks, int64_t* n_points, int64_t* gated_pairs) {
  if (!ctx) return HMSG_ERR_ARG;
  ObjState* st = ctx->obj;
  if (!st) return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_count: no object list");
  if (n_masks) *n_masks = (int64_t)st->a_off.size() - 1;
  if (n_points) *n_points = st->a_off.back();
  if (gated_pairs) *gated_pairs = st->stat_pairs;
  return HMSG_OK;
}
