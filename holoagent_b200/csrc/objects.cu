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

struct ObjFeatScratch {
  long long* d_idx = nullptr; size_t d_idx_bytes = 0;
  double* d_dist = nullptr; size_t d_dist_bytes = 0;
  double* vox = nullptr; size_t vox_bytes = 0;
  int* heads = nullptr; size_t heads_bytes = 0;
  int* hscan = nullptr; size_t hscan_bytes = 0;
  int* valid = nullptr; size_t valid_bytes = 0;
  int* vscan = nullptr; size_t vscan_bytes = 0;
  int* row_node = nullptr; size_t row_node_bytes = 0;
  long long* d_voff = nullptr; size_t d_voff_bytes = 0;
  float* X = nullptr; size_t X_bytes = 0;
  float* Xn = nullptr; size_t Xn_bytes = 0;
  uint32_t* adj = nullptr; size_t adj_bytes = 0;
  float* ff_stage = nullptr; size_t ff_stage_bytes = 0;
  float* out_stage = nullptr; size_t out_stage_bytes = 0;
};

struct ObjState {
  ObjFeatScratch of;
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
  ObjFeatScratch& S = st->of;
  free_dev(S.d_idx); free_dev(S.d_dist); free_dev(S.vox); free_dev(S.heads); free_dev(S.hscan); free_dev(S.valid); free_dev(S.vscan);
  free_dev(S.row_node); free_dev(S.d_voff); free_dev(S.X); free_dev(S.Xn); free_dev(S.adj); free_dev(S.ff_stage); free_dev(S.out_stage);
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

// =================================================================================================
// N2: per-object feature = largest cosine-DBSCAN cluster mean of its node features
// Reference: fsr_vln/memory/hmsg/graph/graph.py:451-488 and utils/graph_utils.py:682-728
//   for mask_3d in self.mask_pcds:
//     mask_3d = mask_3d.voxel_down_sample(voxel_size)          -> (object, cell) stable sort, run means in input order
//     dist, idx = tree_pcd.query(points, k=1); valid = dist <= 0.8 -> node NN on the occupancy bitmap (A4 machinery)
//     feats = np.nan_to_num(self.full_feats_array[idx[valid]])
//     feats = feats_denoise_dbscan(feats, eps=0.01, min_points=100)
//       sklearn DBSCAN(metric="cosine"): neighbours = rows with 1 - <x/|x|, y/|y|> <= eps (float32), core iff
//       >= min_points of them (self included), clusters = components of core points numbered by their first
//       core row, border rows take the lowest-numbered adjacent cluster; Counter.most_common(1) -> largest
//       cluster (ties: first met); result = mean of its rows, or the mean of all rows when nothing clusters.
// Per object: rows are gathered + normalised (warp per row), the n x n thresholded similarity is built once
// as a bit matrix by a 64x64-tile fp32 kernel (GEMM-shaped but it must be fp32: the threshold sits 1e-2 from
// 1.0), then core / union-find / border / sizes run on the bit rows.  Objects with fewer rows than
// min_points cannot have a core point and skip the n^2 work.
// =================================================================================================

__global__ void __launch_bounds__(OTPB) k_obj_vox_keys(const double* __restrict__ xyz, long long npts, const long long* __restrict__ off, int n,
                                                       const double* __restrict__ lo, double vs, unsigned long long* __restrict__ keys,
                                                       int* __restrict__ idx, int* __restrict__ counters) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  int o = mask_of_point(off, n, p);
  unsigned long long key = (unsigned long long)o << 42;
  for (int a = 0; a < 3; a++) {
    double vmin = __dsub_rn(lo[o * 3 + a], __dmul_rn(vs, 0.5));       // Open3D voxel_min_bound = min_bound - voxel_size / 2
    long long c = (long long)floor(cell_coord(xyz[p * 3 + a], vmin, vs));
    if (c < 0 || c > 16383) { atomicExch(&counters[1], 1); c = 0; }
    key |= (unsigned long long)c << (28 - 14 * a);
  }
  keys[p] = key;
  idx[p] = (int)p;
}

__global__ void __launch_bounds__(OTPB) k_run_heads(const unsigned long long* __restrict__ skeys, long long npts, int* __restrict__ heads) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < npts) heads[p] = (p == 0 || skeys[p - 1] != skeys[p]) ? 1 : 0;
}

// one thread per run: sequential float64 sum in input order (stable sort), then / count  (Open3D AccumulatedPoint)
__global__ void __launch_bounds__(OTPB) k_run_means(const unsigned long long* __restrict__ skeys, const int* __restrict__ sidx,
                                                    const int* __restrict__ heads, const int* __restrict__ hscan, long long npts,
                                                    const double* __restrict__ xyz, double* __restrict__ vox) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts || !heads[p]) return;
  unsigned long long key = skeys[p];
  double sx = 0, sy = 0, sz = 0; int c = 0;
  for (long long q = p; q < npts && skeys[q] == key; q++) {
    long long o = sidx[q];
    sx = __dadd_rn(sx, xyz[o * 3]); sy = __dadd_rn(sy, xyz[o * 3 + 1]); sz = __dadd_rn(sz, xyz[o * 3 + 2]);
    c++;
  }
  long long v = hscan[p];
  vox[v * 3] = __ddiv_rn(sx, (double)c); vox[v * 3 + 1] = __ddiv_rn(sy, (double)c); vox[v * 3 + 2] = __ddiv_rn(sz, (double)c);
}

__global__ void k_pick_offsets(const int* __restrict__ scan, const int* __restrict__ flag, const long long* __restrict__ off, int n, long long total,
                               long long* __restrict__ out) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o > n) return;
  long long p = off[o];
  out[o] = p < total ? scan[p] : (total ? scan[total - 1] + flag[total - 1] : 0);
}

__global__ void __launch_bounds__(OTPB) k_valid_rows(const double* __restrict__ dist, long long nv, double max_dist, int* __restrict__ valid) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv) valid[v] = dist[v] <= max_dist ? 1 : 0;
}

__global__ void __launch_bounds__(OTPB) k_compact_rows(const long long* __restrict__ idx, const int* __restrict__ valid, const int* __restrict__ vscan,
                                                       long long nv, int* __restrict__ row_node) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv && valid[v]) row_node[vscan[v]] = (int)idx[v];
}

// warp per row: X = nan_to_num(full_feats[node]); Xn = X / (|X| or 1)   (sklearn normalize: zero norms -> 1)
__global__ void __launch_bounds__(OTPB) k_gather_norm(const float* __restrict__ full, const int* __restrict__ row_node, int n, int d,
                                                      float* __restrict__ X, float* __restrict__ Xn) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  const float* src = full + (long long)row_node[r] * d;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) {
    float v = src[j];
    if (isnan(v)) v = 0.f; else if (isinf(v)) v = v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;     // np.nan_to_num
    X[(long long)r * d + j] = v;
    ss = fmaf(v, v, ss);
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  float nrm = sqrtf(ss);
  if (nrm == 0.f) nrm = 1.f;
  for (int j = lane; j < d; j += 32) Xn[(long long)r * d + j] = __fdiv_rn(X[(long long)r * d + j], nrm);
}

// adj[i][j] = (1 - <Xn_i, Xn_j> <= eps), 64x64 tile per block, 4x4 per thread, fp32
__global__ void __launch_bounds__(256) k_cos_adj(const float* __restrict__ Xn, int n, int d, float eps, uint32_t* __restrict__ adj, int W) {
  __shared__ float sa[32][65], sb[32][65];
  __shared__ uint32_t bits[64][2];
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  float acc[4][4] = {};
  if (threadIdx.x < 128) bits[threadIdx.x >> 1][threadIdx.x & 1] = 0u;
  for (int k0 = 0; k0 < d; k0 += 32) {
    for (int e = threadIdx.x; e < 64 * 32; e += 256) {
      int rr = e >> 5, kk = e & 31;
      sa[kk][rr] = (r0 + rr < n && k0 + kk < d) ? Xn[(long long)(r0 + rr) * d + k0 + kk] : 0.f;
      sb[kk][rr] = (c0 + rr < n && k0 + kk < d) ? Xn[(long long)(c0 + rr) * d + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = sa[kk][ty * 4 + i]; b[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int col = c0 + tx * 4 + j;
      float dist = __fsub_rn(1.0f, acc[i][j]);
      if (col < n && dist <= eps) m |= 1u << (((tx & 7) * 4) + j);
    }
    if (m) atomicOr(&bits[ty * 4 + i][tx >> 3], m);
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    int rr = threadIdx.x >> 1, w = threadIdx.x & 1;
    if (r0 + rr < n) adj[(long long)(r0 + rr) * W + blockIdx.x * 2 + w] = bits[rr][w];
  }
}

__global__ void __launch_bounds__(OTPB) k_adj_core(const uint32_t* __restrict__ adj, int n, int W, int min_points, unsigned char* __restrict__ core) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  int c = 0;
  for (int w = lane; w < W; w += 32) c += __popc(adj[(long long)r * W + w]);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) core[r] = c >= min_points ? 1 : 0;
}

__global__ void __launch_bounds__(OTPB) k_adj_union(const uint32_t* __restrict__ adj, int n, int W, const unsigned char* __restrict__ core,
                                                    int* __restrict__ uf) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n || !core[r]) return;
  for (int w = lane; w * 32 < r; w += 32) {
    uint32_t b = adj[(long long)r * W + w];
    while (b) {
      int j = w * 32 + __ffs(b) - 1;
      b &= b - 1;
      if (j < r && core[j]) uf_union(uf, r, j);
    }
  }
}

__global__ void __launch_bounds__(OTPB) k_adj_label(const uint32_t* __restrict__ adj, int n, int W, const unsigned char* __restrict__ core,
                                                    int* __restrict__ uf, int* __restrict__ label, int* __restrict__ csize, int* __restrict__ cfirst) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  int lab;
  if (core[r]) {
    lab = uf_find(uf, r);
  } else {
    int best = 0x7fffffff;
    for (int w = lane; w < W; w += 32) {
      uint32_t b = adj[(long long)r * W + w];
      while (b) {
        int j = w * 32 + __ffs(b) - 1;
        b &= b - 1;
        if (core[j]) best = min(best, uf_find(uf, j));
      }
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    lab = best == 0x7fffffff ? -1 : best;
  }
  if (lane == 0) {
    label[r] = lab;
    if (lab >= 0) { atomicAdd(&csize[lab], 1); atomicMin(&cfirst[lab], r); }
  }
}

__global__ void __launch_bounds__(OTPB) k_obj_best(const int* __restrict__ label, const int* __restrict__ csize, const int* __restrict__ cfirst, int n,
                                                   unsigned long long* __restrict__ best) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || label[r] != r) return;
  atomicMax(best, ((unsigned long long)(unsigned int)csize[r] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)cfirst[r]));
}

// out[j] = mean over the selected rows (largest cluster, or all rows when best == 0), rows added in order in float32
__global__ void __launch_bounds__(OTPB) k_mean_rows(const float* __restrict__ X, int n, int d, const int* __restrict__ label,
                                                    const unsigned long long* __restrict__ best, float* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  unsigned long long b = best ? *best : 0ull;
  int want = -2;
  if (b) want = label[0xFFFFFFFFu - (unsigned int)(b & 0xFFFFFFFFu)];
  float s = 0.f; int c = 0;
  for (int r = 0; r < n; r++)
    if (want == -2 || label[r] == want) { s = __fadd_rn(s, X[(long long)r * d + j]); c++; }
  out[j] = __fdiv_rn(s, (float)c);
}

extern "C" int32_t hmsg_object_feats(hmsg_ctx* ctx, const float* full_feats, int32_t d, double voxel_size, double max_dist, float eps,
                                     int32_t min_points, float* out, int32_t on_device) {
  if (!ctx) return HMSG_ERR_ARG;
  ObjState* st = ctx->obj;
  if (!st || !st->finished) return ctx->fail(HMSG_ERR_STATE, "hmsg_object_feats: call hmsg_objects_finish first");
  if (!full_feats || !out || d <= 0 || !(voxel_size > 0) || min_points < 1) return ctx->fail(HMSG_ERR_ARG, "hmsg_object_feats: bad argument");
  if (!ctx->nodes_built) return ctx->fail(HMSG_ERR_STATE, "hmsg_object_feats: no node table (hmsg_radius_filter)");
  ObjFeatScratch& S = st->of;
  int32_t rc;
  int n_obj = (int)st->a_off.size() - 1;
  long long npts = st->a_off.back();
  if (n_obj == 0) return HMSG_OK;
  const float* dfull = full_feats; float* dout = out;
  if (!on_device) {
    if ((rc = ctx->reserve(&S.ff_stage, &S.ff_stage_bytes, (size_t)std::max<int64_t>(ctx->n_nodes, 1) * d * 4))) return rc;
    if ((rc = ctx->reserve(&S.out_stage, &S.out_stage_bytes, (size_t)n_obj * d * 4))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(S.ff_stage, full_feats, (size_t)ctx->n_nodes * d * 4, cudaMemcpyHostToDevice, ctx->stream));
    dfull = S.ff_stage; dout = S.out_stage;
  }
  HMSG_CUDA(cudaMemsetAsync(dout, 0, (size_t)n_obj * d * 4, ctx->stream));      // objects without rows: np.zeros((1, dim))
  std::vector<long long> roff(n_obj + 1, 0);
  if (npts > 0) {
    // ---- voxel_down_sample of every object at once
    if ((rc = ctx->reserve(&st->d_off, &st->d_off_bytes, (size_t)(n_obj + 1) * 8))) return rc;
    HMSG_CUDA(cudaMemcpyAsync(st->d_off, st->a_off.data(), (size_t)(n_obj + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = ctx->reserve(&st->d_lo, &st->d_lo_bytes, (size_t)n_obj * 24 + 64))) return rc;
    if ((rc = ctx->reserve(&st->d_hi, &st->d_hi_bytes, (size_t)n_obj * 24))) return rc;
    if ((rc = ctx->reserve(&st->keys, &st->keys_bytes, (size_t)npts * 16))) return rc;
    if ((rc = ctx->reserve(&st->pidx, &st->pidx_bytes, (size_t)npts * 8))) return rc;
    if (!st->d_counters) { HMSG_CUDA(cudaMalloc((void**)&st->d_counters, 16)); HMSG_CUDA(cudaMalloc((void**)&st->d_glob, 24)); }
    HMSG_CUDA(cudaMemsetAsync(st->d_counters, 0, 16, ctx->stream));
    k_mask_aabb<<<n_obj, 128, 0, ctx->stream>>>(st->a_xyz, st->d_off, st->d_lo, st->d_hi);
    k_obj_vox_keys<<<blocks_for(npts), OTPB, 0, ctx->stream>>>(st->a_xyz, npts, st->d_off, n_obj, st->d_lo, voxel_size, st->keys, st->pidx, st->d_counters);
    size_t tmp = 0;
    unsigned long long* kout = st->keys + npts; int* iout = st->pidx + npts;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, st->keys, kout, st->pidx, iout, (int)npts, 0, 63, ctx->stream);
    if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
    HMSG_CUDA(cub::DeviceRadixSort::SortPairs(st->sort_tmp, tmp, st->keys, kout, st->pidx, iout, (int)npts, 0, 63, ctx->stream));
    if ((rc = ctx->reserve(&S.heads, &S.heads_bytes, (size_t)npts * 4))) return rc;
    if ((rc = ctx->reserve(&S.hscan, &S.hscan_bytes, (size_t)npts * 4))) return rc;
    k_run_heads<<<blocks_for(npts), OTPB, 0, ctx->stream>>>(kout, npts, S.heads);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, S.heads, S.hscan, (int)npts, ctx->stream);
    if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
    HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->sort_tmp, tmp, S.heads, S.hscan, (int)npts, ctx->stream));
    if ((rc = ctx->reserve(&S.d_voff, &S.d_voff_bytes, (size_t)(n_obj + 1) * 8 * 2))) return rc;
    k_pick_offsets<<<blocks_for(n_obj + 1), OTPB, 0, ctx->stream>>>(S.hscan, S.heads, st->d_off, n_obj, npts, S.d_voff);
    HMSG_LAUNCH_CHECK();
    std::vector<long long> voff(n_obj + 1);
    int flag[2];
    HMSG_CUDA(cudaMemcpyAsync(voff.data(), S.d_voff, (size_t)(n_obj + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaMemcpyAsync(flag, st->d_counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag[1]) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_object_feats: an object spans more than 2^14 voxels per axis");
    long long nv = voff[n_obj];
    if ((rc = ctx->reserve(&S.vox, &S.vox_bytes, (size_t)nv * 24))) return rc;
    if ((rc = ctx->reserve(&S.d_idx, &S.d_idx_bytes, (size_t)nv * 8))) return rc;
    if ((rc = ctx->reserve(&S.d_dist, &S.d_dist_bytes, (size_t)nv * 8))) return rc;
    if ((rc = ctx->reserve(&S.valid, &S.valid_bytes, (size_t)nv * 4))) return rc;
    if ((rc = ctx->reserve(&S.vscan, &S.vscan_bytes, (size_t)nv * 4))) return rc;
    if ((rc = ctx->reserve(&S.row_node, &S.row_node_bytes, (size_t)nv * 4))) return rc;
    k_run_means<<<blocks_for(npts), OTPB, 0, ctx->stream>>>(kout, iout, S.heads, S.hscan, npts, st->a_xyz, S.vox);
    HMSG_LAUNCH_CHECK();
    // ---- nearest node + distance gate (graph.py:459-473)
    if ((rc = geometry_points_to_node_dev(ctx, S.vox, nv, (int64_t*)S.d_idx, S.d_dist))) return rc;
    k_valid_rows<<<blocks_for(nv), OTPB, 0, ctx->stream>>>(S.d_dist, nv, max_dist, S.valid);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, S.valid, S.vscan, (int)nv, ctx->stream);
    if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
    HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->sort_tmp, tmp, S.valid, S.vscan, (int)nv, ctx->stream));
    k_compact_rows<<<blocks_for(nv), OTPB, 0, ctx->stream>>>(S.d_idx, S.valid, S.vscan, nv, S.row_node);
    long long* d_roff = S.d_voff + (n_obj + 1);
    HMSG_CUDA(cudaMemcpyAsync(S.d_voff, voff.data(), (size_t)(n_obj + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_pick_offsets<<<blocks_for(n_obj + 1), OTPB, 0, ctx->stream>>>(S.vscan, S.valid, S.d_voff, n_obj, nv, d_roff);
    HMSG_LAUNCH_CHECK();
    HMSG_CUDA(cudaMemcpyAsync(roff.data(), d_roff, (size_t)(n_obj + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // ---- per object: gather, cosine DBSCAN, cluster mean
  long long nmax = 0;
  for (int o = 0; o < n_obj; o++) nmax = std::max(nmax, roff[o + 1] - roff[o]);
  if (nmax == 0) goto done;
  if (nmax > 65536) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_object_feats: an object has more than 65536 node rows (bit-matrix DBSCAN limit)");
  {
    int Wmax = (int)(((nmax + 63) / 64) * 2);
    if ((rc = ctx->reserve(&S.X, &S.X_bytes, (size_t)nmax * d * 4))) return rc;
    if ((rc = ctx->reserve(&S.Xn, &S.Xn_bytes, (size_t)nmax * d * 4))) return rc;
    if (nmax >= min_points && (rc = ctx->reserve(&S.adj, &S.adj_bytes, (size_t)nmax * Wmax * 4))) return rc;
    if ((rc = ctx->reserve(&st->core, &st->core_bytes, (size_t)nmax))) return rc;
    if ((rc = ctx->reserve(&st->uf, &st->uf_bytes, (size_t)nmax * 4))) return rc;
    if ((rc = ctx->reserve(&st->label, &st->label_bytes, (size_t)nmax * 4))) return rc;
    if ((rc = ctx->reserve(&st->csize, &st->csize_bytes, (size_t)nmax * 4))) return rc;
    if ((rc = ctx->reserve(&st->cfirst, &st->cfirst_bytes, (size_t)nmax * 4))) return rc;
    if ((rc = ctx->reserve(&st->best, &st->best_bytes, 8))) return rc;
    for (int o = 0; o < n_obj; o++) {
      int n = (int)(roff[o + 1] - roff[o]);
      if (n == 0) continue;
      unsigned wb = blocks_for((long long)n * 32);
      k_gather_norm<<<wb, OTPB, 0, ctx->stream>>>(dfull, S.row_node + roff[o], n, d, S.X, S.Xn);
      HMSG_CUDA(cudaMemsetAsync(st->best, 0, 8, ctx->stream));
      if (n >= min_points) {
        int W = ((n + 63) / 64) * 2;
        dim3 tg((n + 63) / 64, (n + 63) / 64);
        k_cos_adj<<<tg, 256, 0, ctx->stream>>>(S.Xn, n, d, eps, S.adj, W);
        k_adj_core<<<wb, OTPB, 0, ctx->stream>>>(S.adj, n, W, min_points, st->core);
        k_iota<<<blocks_for(n), OTPB, 0, ctx->stream>>>(st->uf, n);
        k_adj_union<<<wb, OTPB, 0, ctx->stream>>>(S.adj, n, W, st->core, st->uf);
        HMSG_CUDA(cudaMemsetAsync(st->csize, 0, (size_t)n * 4, ctx->stream));
        HMSG_CUDA(cudaMemsetAsync(st->cfirst, 0x7f, (size_t)n * 4, ctx->stream));
        k_adj_label<<<wb, OTPB, 0, ctx->stream>>>(S.adj, n, W, st->core, st->uf, st->label, st->csize, st->cfirst);
        k_obj_best<<<blocks_for(n), OTPB, 0, ctx->stream>>>(st->label, st->csize, st->cfirst, n, st->best);
      }
      k_mean_rows<<<blocks_for(d), OTPB, 0, ctx->stream>>>(S.X, n, d, st->label, st->best, dout + (size_t)o * d);
      HMSG_LAUNCH_CHECK();
    }
  }
done:
  if (!on_device) {
    HMSG_CUDA(cudaMemcpyAsync(out, dout, (size_t)n_obj * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HMSG_OK;
}

// =================================================================================================
// A7 -> N1 chained on the device: the 3-D masks of one frame (generic.py:141-190 create_3d_masks) go
// straight into the merge without visiting the host.
//   pcd_masked = create_pcd(mask, depth, pose, mask_img=True, filter_distance)   -> the frame's pixel->node map
//   pcd_masked = pcd[indices]                  one node centroid PER MASK PIXEL (multiplicity kept)
//   pcd_mask.voxel_down_sample(down_size)      relative to the mask's own min bound
// Entries (mask, pixel) are ordered by a first radix sort, keyed by (mask, cell) and ordered by a second
// STABLE sort, so each voxel's float64 sum runs over its pixels in row-major order exactly like Open3D's
// sequential accumulation: the mask point sets are bit-identical to the reference's, not just close.
// filter_distance: a mask whose mean depth exceeds it yields an empty cloud (generic.py:126-127; the mean is
// taken in float64 here, float32 pairwise in numpy - they differ only within 1e-7 of the threshold).
// =================================================================================================
__global__ void __launch_bounds__(OTPB) k_fm_stats(const int32_t* __restrict__ pidx, const uint32_t* __restrict__ mbits, const uint16_t* __restrict__ depth,
                                                   float scale, int HW, int MW, const double* __restrict__ nodes, int* __restrict__ cnt,
                                                   double* __restrict__ dsum, long long* __restrict__ mb) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  int n = pidx[p];
  if (n < 0) return;
  double df = (double)__fdiv_rn((float)depth[p], scale);
  for (int w = 0; w < MW; w++) {
    uint32_t bits = mbits[(long long)p * MW + w];
    while (bits) {
      int m = __ffs(bits) - 1 + 32 * w;
      bits &= bits - 1;
      atomicAdd(&cnt[m], 1);
      atomicAdd(&dsum[m], df);
      for (int k = 0; k < 3; k++) atomicMin(&mb[m * 3 + k], d2ord(nodes[(long long)n * 3 + k]));
    }
  }
}

__global__ void __launch_bounds__(OTPB) k_fm_entries(const int32_t* __restrict__ pidx, const uint32_t* __restrict__ mbits, int HW, int MW,
                                                     const long long* __restrict__ eoff, const unsigned char* __restrict__ keepm, int* __restrict__ cursor,
                                                     unsigned long long* __restrict__ ekey) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW || pidx[p] < 0) return;
  for (int w = 0; w < MW; w++) {
    uint32_t bits = mbits[(long long)p * MW + w];
    while (bits) {
      int m = __ffs(bits) - 1 + 32 * w;
      bits &= bits - 1;
      if (!keepm[m]) continue;
      long long pos = eoff[m] + atomicAdd(&cursor[m], 1);
      ekey[pos] = ((unsigned long long)m << 32) | (unsigned int)p;
    }
  }
}

__global__ void __launch_bounds__(OTPB) k_fm_cellkeys(const unsigned long long* __restrict__ ekey, long long E, const int32_t* __restrict__ pidx,
                                                      const double* __restrict__ nodes, const long long* __restrict__ mb, double vs,
                                                      unsigned long long* __restrict__ keys, int* __restrict__ vals, int* __restrict__ counters) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  unsigned long long ek = ekey[e];
  int m = (int)(ek >> 32), p = (int)(ek & 0xffffffffu);
  int n = pidx[p];
  unsigned long long key = (unsigned long long)m << 42;
  for (int k = 0; k < 3; k++) {
    long long o = mb[m * 3 + k];
    long long bb = o >= 0 ? o : (o ^ 0x7FFFFFFFFFFFFFFFLL);
    double vmin = __dsub_rn(__longlong_as_double(bb), __dmul_rn(vs, 0.5));
    long long c = (long long)floor(cell_coord(nodes[(long long)n * 3 + k], vmin, vs));
    if (c < 0 || c > 16383) { atomicExch(&counters[1], 1); c = 0; }
    key |= (unsigned long long)c << (28 - 14 * k);
  }
  keys[e] = key;
  vals[e] = n;
}

__global__ void __launch_bounds__(OTPB) k_fm_means(const unsigned long long* __restrict__ skeys, const int* __restrict__ snode, const int* __restrict__ heads,
                                                   const int* __restrict__ hscan, long long E, const double* __restrict__ nodes,
                                                   const double* __restrict__ nrgb, double* __restrict__ oxyz, double* __restrict__ orgb) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || !heads[e]) return;
  unsigned long long key = skeys[e];
  double s[6] = {0, 0, 0, 0, 0, 0}; int c = 0;
  for (long long q = e; q < E && skeys[q] == key; q++) {
    long long n = snode[q];
    for (int k = 0; k < 3; k++) { s[k] = __dadd_rn(s[k], nodes[n * 3 + k]); s[3 + k] = __dadd_rn(s[3 + k], nrgb[n * 3 + k]); }
    c++;
  }
  long long v = hscan[e];
  for (int k = 0; k < 3; k++) { oxyz[v * 3 + k] = __ddiv_rn(s[k], (double)c); orgb[v * 3 + k] = __ddiv_rn(s[3 + k], (double)c); }
}

extern "C" int32_t hmsg_objects_add_frame(hmsg_ctx* ctx, int64_t frame, double down_size, double filter_distance) {
  if (!ctx) return HMSG_ERR_ARG;
  ObjState* st = ctx->obj;
  if (!st || st->finished) return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_add_frame: call hmsg_objects_begin first");
  if (ctx->batch_begin < 0 || frame < ctx->batch_begin || frame >= ctx->batch_begin + ctx->batch_n)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_add_frame: frame is not in the current mask batch (hmsg_masks_*)");
  if (!(down_size > 0)) return ctx->fail(HMSG_ERR_ARG, "hmsg_objects_add_frame: bad down_size");
  int32_t rc;
  if ((rc = features_ensure_pix_idx(ctx))) return rc;
  ObjFeatScratch& S = st->of;
  int M = ctx->batch_M, MW = ctx->batch_MW, HW = ctx->cam.H * ctx->cam.W;
  int fb = (int)(frame - ctx->batch_begin);
  const int32_t* pidx = ctx->pix_idx + (size_t)fb * HW;
  const uint32_t* mbits = ctx->maskbits + (size_t)fb * HW * MW;
  const uint16_t* depth = ctx->depth + (size_t)frame * HW;
  // per-mask scratch: cnt[M] int | cursor[M] int | dsum[M] double | mb[3M] ll | eoff[M+1] ll | keep[M] u8
  size_t need = (size_t)M * 4 * 2 + (size_t)M * 8 + (size_t)M * 24 + (size_t)(M + 1) * 8 + M + 64;
  if ((rc = ctx->reserve((char**)&ctx->scratch, &ctx->scratch_bytes, need))) return rc;
  char* base = (char*)ctx->scratch;
  double* dsum = (double*)base;
  long long* mb = (long long*)(dsum + M);
  long long* d_eoff = mb + 3 * M;
  int* cnt = (int*)(d_eoff + M + 1);
  int* cursor = cnt + M;
  unsigned char* d_keep = (unsigned char*)(cursor + M);
  std::vector<long long> init(3 * M);
  { double pinf = INFINITY; long long a; memcpy(&a, &pinf, 8); for (auto& v : init) v = a; }
  HMSG_CUDA(cudaMemsetAsync(base, 0, need, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(mb, init.data(), (size_t)M * 24, cudaMemcpyHostToDevice, ctx->stream));
  ctx->wait_frames(frame, 1);
  k_fm_stats<<<blocks_for(HW), OTPB, 0, ctx->stream>>>(pidx, mbits, depth, ctx->cam.scale, HW, MW, ctx->node_xyz, cnt, dsum, mb);
  HMSG_LAUNCH_CHECK();
  std::vector<int> hcnt(M); std::vector<double> hsum(M);
  HMSG_CUDA(cudaMemcpyAsync(hcnt.data(), cnt, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(hsum.data(), dsum, (size_t)M * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<long long> eoff(M + 1, 0); std::vector<unsigned char> keepm(M, 0);
  const int Mr = ctx->batch_counts[fb];                                            // real masks of this frame; padded slots are not list entries
  for (int m = 0; m < M; m++) {
    bool keep = m < Mr && hcnt[m] > 0 && !(hsum[m] / (double)hcnt[m] > filter_distance);     // `if Z.mean() > filter_distance: return empty`
    keepm[m] = keep ? 1 : 0;
    eoff[m + 1] = eoff[m] + (keep ? hcnt[m] : 0);
  }
  long long E = eoff[M];
  std::vector<int64_t> voff(M + 1, 0);
  if (E == 0) return hmsg_objects_add_masks(ctx, Mr, voff.data(), nullptr, nullptr, 1);
  if (E >= (1LL << 31)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_objects_add_frame: too many mask pixels");
  HMSG_CUDA(cudaMemcpyAsync(d_eoff, eoff.data(), (size_t)(M + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(d_keep, keepm.data(), (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = ctx->reserve(&st->keys, &st->keys_bytes, (size_t)E * 16))) return rc;
  if ((rc = ctx->reserve(&st->pidx, &st->pidx_bytes, (size_t)E * 8))) return rc;
  if (!st->d_counters) { HMSG_CUDA(cudaMalloc((void**)&st->d_counters, 16)); HMSG_CUDA(cudaMalloc((void**)&st->d_glob, 24)); }
  HMSG_CUDA(cudaMemsetAsync(st->d_counters, 0, 16, ctx->stream));
  unsigned long long* k0 = st->keys; unsigned long long* k1 = st->keys + E;
  int* v0 = st->pidx; int* v1 = st->pidx + E;
  k_fm_entries<<<blocks_for(HW), OTPB, 0, ctx->stream>>>(pidx, mbits, HW, MW, d_eoff, d_keep, cursor, k0);
  size_t tmp = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp, k0, k1, (int)E, 0, 43, ctx->stream);
  cub::DeviceRadixSort::SortPairs(nullptr, tmp2, k0, k1, v0, v1, (int)E, 0, 53, ctx->stream);
  if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, std::max(tmp, tmp2)))) return rc;
  HMSG_CUDA(cub::DeviceRadixSort::SortKeys(st->sort_tmp, tmp, k0, k1, (int)E, 0, 43, ctx->stream));                 // (mask, pixel) order
  k_fm_cellkeys<<<blocks_for(E), OTPB, 0, ctx->stream>>>(k1, E, pidx, ctx->node_xyz, mb, down_size, k0, v0, st->d_counters);
  HMSG_CUDA(cub::DeviceRadixSort::SortPairs(st->sort_tmp, tmp2, k0, k1, v0, v1, (int)E, 0, 53, ctx->stream));        // stable: (mask, cell), pixels in order
  if ((rc = ctx->reserve(&S.heads, &S.heads_bytes, (size_t)E * 4))) return rc;
  if ((rc = ctx->reserve(&S.hscan, &S.hscan_bytes, (size_t)E * 4))) return rc;
  k_run_heads<<<blocks_for(E), OTPB, 0, ctx->stream>>>(k1, E, S.heads);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, S.heads, S.hscan, (int)E, ctx->stream);
  if ((rc = ctx->reserve(&st->sort_tmp, &st->sort_tmp_bytes, tmp))) return rc;
  HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->sort_tmp, tmp, S.heads, S.hscan, (int)E, ctx->stream));
  if ((rc = ctx->reserve(&S.d_voff, &S.d_voff_bytes, (size_t)(M + 1) * 8 * 2))) return rc;
  k_pick_offsets<<<blocks_for(M + 1), OTPB, 0, ctx->stream>>>(S.hscan, S.heads, d_eoff, M, E, S.d_voff);
  HMSG_LAUNCH_CHECK();
  int flag[2];
  std::vector<long long> hv(M + 1);
  HMSG_CUDA(cudaMemcpyAsync(hv.data(), S.d_voff, (size_t)(M + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(flag, st->d_counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (flag[1]) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_objects_add_frame: a mask spans more than 2^14 voxels per axis");
  long long nv = hv[M];
  if ((rc = ctx->reserve(&st->in_xyz, &st->in_xyz_bytes, (size_t)nv * 24))) return rc;
  if ((rc = ctx->reserve(&st->in_rgb, &st->in_rgb_bytes, (size_t)nv * 24))) return rc;
  k_fm_means<<<blocks_for(E), OTPB, 0, ctx->stream>>>(k1, v1, S.heads, S.hscan, E, ctx->node_xyz, ctx->node_rgb, st->in_xyz, st->in_rgb);
  HMSG_LAUNCH_CHECK();
  for (int m = 0; m <= M; m++) voff[m] = hv[m];
  return hmsg_objects_add_masks(ctx, Mr, voff.data(), st->in_xyz, st->in_rgb, 1);   // slots >= Mr are empty: offsets[Mr] is the total
}

// the frame masks staged by the last hmsg_objects_add_frame are not kept; this reads the CURRENT global list instead
