// masks3d.cu - A7 for a whole frame batch: RGBDDataset.create_3d_masks (fsr_vln/memory/hmsg/dataloader/
// generic.py:140-190) as called once per frame by Graph.create_feature_map (graph/graph.py:391-402).
//
// Reference, per mask of a frame:
//   pcd_masked = create_pcd(mask, depth, pose, mask_img=True, filter_distance)   points of the mask pixels with depth > 0
//   dist, indices = full_pcd_tree.query(pcd_masked, k=1)                         == the frame's pixel -> node map (A4)
//   pcd_mask.points = pcd[indices]; .colors = colors[indices]                    one node centroid PER MASK PIXEL
//   pcd_mask = pcd_mask.voxel_down_sample(down_size)                             relative to the mask's own min bound
//
// A mask pixel contributes its node's centroid, so a (mask, voxel) mean is sum_n k_n * c_n / sum_n k_n over the
// distinct nodes n whose centroid falls in the voxel, k_n = number of mask pixels mapped to n.  The batch path
// therefore never touches per-pixel point data:
//   k_m3d_scan      one thread per pixel of the batch; a warp's 32 pixels are grouped by (mask, node) with
//                   match_any / ballot, the first lane of a group adds its pixel count to a (job, node) hash
//                   table in HBM (job = frame-in-batch * M + mask) - ~1 atomic per 10 mask pixels - and folds the
//                   node centroid into the job's min bound; the job's pixel count and integer depth sum (for
//                   the `Z.mean() > filter_distance` test of generic.py:126) ride on the same ballots
//   k_m3d_place     (job, node) entry -> voxel key floor((c - (min - vs/2)) / vs) with the reference's float64 ops
//   radix sort by (job, i, j, k)  => the canonical ascending-key output order of the oracle (H2)
//   k_m3d_means     one thread per (job, voxel) run: the <= 8 nodes of the run are visited in ascending node order
//                   (deterministic), sum += k_n * c_n, mean = sum / count
// Open3D adds the centroid once per pixel in row-major order; k_n * c_n summed by node differs from that by
// reordering of float64 additions only (<= 1e-15 relative): voxel keys and counts are exact, means agree to 1e-12
// (tests/test_gpu_masks3d.py).  The per-frame path hmsg_objects_add_frame keeps the bit-identical ordered sums.
// Only per-batch scalars (entry count, overflow flags) visit the host; results stay in an HBM store that
// hmsg_objects_merge_stored feeds to the N1 merge (frames_pcd of graph.py:399).
#include "common.cuh"
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#define MTPB 256
static const unsigned long long M3D_EMPTY = ~0ull;

struct M3dBatchRec {
  int64_t frame_begin = 0;
  int n_frames = 0, M = 0;
  std::vector<int> counts;          // real masks per frame (ragged; <= M)
  double* xyz = nullptr;            // [cap_pts,3]
  double* rgb = nullptr;
  int32_t* ijk = nullptr;
  long long* d_off = nullptr;       // [n_frames*M + 1] device offsets into xyz/rgb/ijk
  std::vector<long long> h_off;     // lazily copied
  long long cap_pts = 0;
};

struct M3dChunk { char* base = nullptr; size_t cap = 0, used = 0; };

struct M3dState {
  std::vector<M3dChunk> chunks;
  size_t cur = 0;
  std::vector<M3dBatchRec> batches;
  M3dBatchRec scratch;
  bool scratch_valid = false;
  double scratch_down = 0, scratch_filter = 0;
  // work buffers (grow-only)
  unsigned long long* hkeys = nullptr; size_t hkeys_bytes = 0;
  uint32_t* hcnt = nullptr; size_t hcnt_bytes = 0;
  uint32_t* entry_slot = nullptr; size_t entry_slot_bytes = 0;
  size_t hcap = 0;
  int* counters = nullptr;                                   // [0] entries, [1] table overflow, [2] cell range overflow
  char* jobs = nullptr; size_t jobs_bytes = 0;               // cnt int | dsum ull | mn ll[3] | keep u8
  unsigned long long* skeys = nullptr; size_t skeys_bytes = 0;   // 2 x entries
  int* svals = nullptr; size_t svals_bytes = 0;               // 2 x entries
  int* heads = nullptr; size_t heads_bytes = 0;
  int* hscan = nullptr; size_t hscan_bytes = 0;
  unsigned char* tmp = nullptr; size_t tmp_bytes = 0;
  char* sout = nullptr; size_t sout_bytes = 0;                // scratch-mode outputs
};

__device__ __forceinline__ unsigned long long m3d_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

__global__ void __launch_bounds__(MTPB) k_m3d_init(int n_jobs, int* cnt, int* nent, unsigned long long* dsum, long long* mn, unsigned long long* hkeys,
                                                   uint32_t* hcnt, unsigned long long hsize, int* counters) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hsize) { hkeys[i] = M3D_EMPTY; hcnt[i] = 0u; }
  if (i <= (unsigned long long)n_jobs) nent[i] = 0;
  if (i < (unsigned long long)n_jobs) {
    cnt[i] = 0; dsum[i] = 0ull;
    const long long pinf = 0x7FF0000000000000LL;             // d2ord(+inf)
    mn[i * 3] = pinf; mn[i * 3 + 1] = pinf; mn[i * 3 + 2] = pinf;
  }
  if (i < 4) counters[i] = 0;
}

// one thread per pixel; grid = (n_frames, ceil(HW / MTPB)): the FRAME is the fast block index, so blocks resident at the same
// time work on different frames - their per-job atomics (count, depth sum, min bound) go to n_frames x M different addresses
// instead of queueing on the 32 counters of one frame (raster order over one frame at a time was measured 2x slower)
__global__ void __launch_bounds__(MTPB) k_m3d_scan(const int32_t* __restrict__ pix_idx, const uint32_t* __restrict__ maskbits,
                                                   const uint16_t* __restrict__ depth, int HW, int M, int MW, const int32_t* __restrict__ mask_cnt,
                                                   const double* __restrict__ nodes, unsigned long long* __restrict__ hkeys,
                                                   uint32_t* __restrict__ hcnt, unsigned long long hmask, uint32_t* __restrict__ entry_slot,
                                                   uint32_t entry_cap, int* __restrict__ counters, int* __restrict__ job_cnt, int* __restrict__ job_nent,
                                                   unsigned long long* __restrict__ job_dsum, long long* __restrict__ job_mn) {
  const int fb = blockIdx.x;
  const int p = blockIdx.y * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const long long gp = (long long)fb * HW + p;
  int n = (p < HW) ? pix_idx[gp] : -1;
  const unsigned dep = (n >= 0) ? (unsigned)depth[gp] : 0u;
  const unsigned grp = __match_any_sync(0xffffffffu, n);
  const int mreal = mask_cnt[fb];
  for (int w = 0; w < MW; w++) {
    uint32_t bits = (n >= 0) ? __ldg(&maskbits[gp * MW + w]) : 0u;
    const int left = mreal - w * 32;                          // padded mask slots never produce jobs
    if (left < 32) bits &= (left <= 0) ? 0u : ((1u << left) - 1u);
    uint32_t uni = __reduce_or_sync(0xffffffffu, bits);
    while (uni) {
      const int ml = __ffs(uni) - 1;
      uni &= uni - 1;
      const bool has = (bits >> ml) & 1u;
      const unsigned b = __ballot_sync(0xffffffffu, has);
      if (has) {
        const int job = fb * M + w * 32 + ml;
        const unsigned dsum = __reduce_add_sync(b, dep);
        if ((b & lt) == 0u) {
          atomicAdd(&job_cnt[job], __popc(b));
          atomicAdd(&job_dsum[job], (unsigned long long)dsum);
        }
        const unsigned mine = b & grp;
        if ((mine & lt) == 0u) {                              // first lane of this (mask, node) group
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const long long o = d2ord(nodes[(long long)n * 3 + k]);
            if (o < job_mn[job * 3 + k]) atomicMin(&job_mn[job * 3 + k], o);
          }
          const unsigned long long key = ((unsigned long long)job << 32) | (unsigned)n;
          unsigned long long h = m3d_hash(key) & hmask;
          bool done = false;
          for (unsigned long long probe = 0; probe <= hmask && !done; probe++) {
            unsigned long long cur = hkeys[h];
            if (cur == M3D_EMPTY) {
              cur = atomicCAS(&hkeys[h], M3D_EMPTY, key);
              if (cur == M3D_EMPTY) cur = key;          // claimed; the used slots are collected afterwards by k_m3d_compact
            }
            if (cur == key) { atomicAdd(&hcnt[h], (uint32_t)__popc(mine)); done = true; }
            else h = (h + 1) & hmask;
          }
          if (!done) atomicExch(&counters[1], 1);
        }
      }
    }
  }
}

// used slots -> entry list.  One atomicAdd per 1024 slots (block-aggregated); the order of the list is irrelevant
// (a sort by voxel key follows and every voxel visits its nodes in ascending id).
__global__ void __launch_bounds__(MTPB) k_m3d_compact(const unsigned long long* __restrict__ hkeys, unsigned long long hsize,
                                                      uint32_t* __restrict__ entry_slot, uint32_t entry_cap, int* __restrict__ counters) {
  __shared__ int s_warp[MTPB / 32];
  __shared__ int s_base;
  const unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool used[4]; int c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { used[k] = (i0 + k < hsize) && hkeys[i0 + k] != M3D_EMPTY; c += used[k] ? 1 : 0; }
  int incl = c;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < MTPB / 32; w++) { const int v = s_warp[w]; s_warp[w] = t; t += v; }
    s_base = t ? atomicAdd(&counters[0], t) : 0;
  }
  __syncthreads();
  int pos = s_base + s_warp[warp] + incl - c;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (used[k]) { if ((uint32_t)pos < entry_cap) entry_slot[pos] = (uint32_t)(i0 + k); else atomicExch(&counters[1], 1); pos++; }
}

// keep[job] = mask has pixels and not (mean depth > filter_distance)   (generic.py:126-127; a mask without valid
// pixels yields an empty cloud).  mean = (sum of uint16 depths / count) / scale in float64; the reference takes the
// float32 mean of depth/scale - they differ only within 1e-7 relative of the threshold.
__global__ void k_m3d_jobs(int n_jobs, const int* __restrict__ cnt, const unsigned long long* __restrict__ dsum, float scale, double filter_distance,
                           unsigned char* __restrict__ keep) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_jobs) return;
  bool k = cnt[j] > 0;
  if (k) {
    double mean = ((double)dsum[j] / (double)cnt[j]) / (double)scale;
    if (mean > filter_distance) k = false;
  }
  keep[j] = k ? 1 : 0;
}

// entry e -> sort key job << 42 | i << 28 | j << 14 | k
__global__ void __launch_bounds__(MTPB) k_m3d_place(int n_entries, const uint32_t* __restrict__ entry_slot, const unsigned long long* __restrict__ hkeys,
                                                    const long long* __restrict__ job_mn, const double* __restrict__ nodes, double vs,
                                                    unsigned long long* __restrict__ skeys, int* __restrict__ svals, int* __restrict__ counters) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const unsigned long long hk = hkeys[entry_slot[e]];
  const int job = (int)(hk >> 32);
  const long long n = (long long)(hk & 0xffffffffull);
  unsigned long long key = (unsigned long long)job << 42;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const long long o = job_mn[job * 3 + k];
    const long long bb = o >= 0 ? o : (o ^ 0x7FFFFFFFFFFFFFFFLL);
    const double vmin = __dsub_rn(__longlong_as_double(bb), __dmul_rn(vs, 0.5));     // Open3D: min_bound - voxel_size * 0.5
    long long c = (long long)floor(cell_coord(nodes[n * 3 + k], vmin, vs));
    if (c < 0 || c > 16383) { atomicExch(&counters[2], 1); c = 0; }
    key |= (unsigned long long)c << (28 - 14 * k);
  }
  skeys[e] = key;
  svals[e] = e;
}

// filtered masks (generic.py:126-127) keep their bucket but start no voxel
__global__ void __launch_bounds__(MTPB) k_m3d_heads(const unsigned long long* __restrict__ skeys, int n, const unsigned char* __restrict__ keep,
                                                    int* __restrict__ heads) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const unsigned long long k = skeys[e];
  heads[e] = (keep[(int)(k >> 42)] && (e == 0 || skeys[e - 1] != k)) ? 1 : 0;
}

__global__ void __launch_bounds__(MTPB) k_m3d_means(const unsigned long long* __restrict__ skeys, const int* __restrict__ svals, const int* __restrict__ heads,
                                                    const int* __restrict__ hscan, int n, const uint32_t* __restrict__ entry_slot,
                                                    const unsigned long long* __restrict__ hkeys, const uint32_t* __restrict__ hcnt,
                                                    const double* __restrict__ nodes, const double* __restrict__ nrgb, double* __restrict__ oxyz,
                                                    double* __restrict__ orgb, int32_t* __restrict__ oijk) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !heads[e]) return;
  const unsigned long long key = skeys[e];
  int q1 = e + 1;
  while (q1 < n && skeys[q1] == key) q1++;
  double s[6] = {0, 0, 0, 0, 0, 0};
  unsigned long long total = 0;
  long long last = -1;
  for (int it = e; it < q1; it++) {                            // selection by ascending node id: deterministic for any run length
    long long best = 0x7FFFFFFFFFFFFFFFLL; uint32_t bc = 0;
    for (int q = e; q < q1; q++) {
      const uint32_t slot = entry_slot[svals[q]];
      const long long nd = (long long)(hkeys[slot] & 0xffffffffull);
      if (nd > last && nd < best) { best = nd; bc = hcnt[slot]; }
    }
    last = best;
    const double kd = (double)bc;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      s[k] = __dadd_rn(s[k], __dmul_rn(kd, nodes[best * 3 + k]));
      s[3 + k] = __dadd_rn(s[3 + k], __dmul_rn(kd, nrgb[best * 3 + k]));
    }
    total += bc;
  }
  const long long v = hscan[e];
  const double t = (double)total;
#pragma unroll
  for (int k = 0; k < 3; k++) { oxyz[v * 3 + k] = __ddiv_rn(s[k], t); orgb[v * 3 + k] = __ddiv_rn(s[3 + k], t); }
  oijk[v * 3] = (int32_t)((key >> 28) & 0x3FFF); oijk[v * 3 + 1] = (int32_t)((key >> 14) & 0x3FFF); oijk[v * 3 + 2] = (int32_t)(key & 0x3FFF);
}

// off[j] = number of voxels of jobs < j  (lower bound of j << 42 in the sorted keys -> heads before it)
__global__ void k_m3d_offsets(const unsigned long long* __restrict__ skeys, const int* __restrict__ heads, const int* __restrict__ hscan, int n,
                              int n_jobs, long long* __restrict__ off) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_jobs) return;
  const unsigned long long target = (unsigned long long)j << 42;
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (skeys[mid] < target) lo = mid + 1; else hi = mid; }
  off[j] = (lo < n) ? (long long)hscan[lo] : (n > 0 ? (long long)hscan[n - 1] + heads[n - 1] : 0);
}


// =================================================================================================
// Per-job on-chip path (opt-in, hmsg_set_option("mask3d_path", 0)): ONE thread block per (frame, mask).  The block scans the mask's pixel bounding box,
// counts pixels per node in a shared-memory hash table, derives the voxel keys, sorts them with a shared-memory bitonic
// sort and reduces every voxel run - nothing but the final voxels touches HBM.  Two launches per batch: a count pass
// (voxels per job -> exclusive scan -> exact, contiguous output offsets) and a write pass.
// Jobs whose node set does not fit on chip (> JOB_SC distinct nodes) raise a flag and the batch is redone by the global-hash
// path above.  On the BASELINE workload (rectangles up to 256 x 256 px seeing walls at 5 - 10 m: up to ~20 k nodes per mask) that
// happens in every batch, which is why the global path is the default; this path serves workloads of small masks.
// =================================================================================================
constexpr int JOB_THREADS = 512;
constexpr int JOB_HC = 16384;            // hash slots (node, count)
constexpr int JOB_SC = 8192;             // distinct nodes of one mask sorted on chip (bitonic: a power of two)
constexpr int JOB_SMEM = JOB_HC * 8 + JOB_SC * 8 + 4096;   // hk + hc | sort keys | reductions  (196 KB: one block per SM)

template <bool WRITE>
__global__ void __launch_bounds__(JOB_THREADS, 1) k_m3d_job(const int32_t* __restrict__ pix_idx, const uint32_t* __restrict__ maskbits,
                                                            const uint16_t* __restrict__ depth, int HW, int W, int M, int MW,
                                                            const int32_t* __restrict__ mask_cnt, const int32_t* __restrict__ mask_rect,
                                                            const double* __restrict__ nodes, const double* __restrict__ nrgb, double vs, float scale,
                                                            double filter_distance, int* __restrict__ job_cells, const long long* __restrict__ job_off,
                                                            int* __restrict__ counters, double* __restrict__ oxyz, double* __restrict__ orgb,
                                                            int32_t* __restrict__ oijk) {
  extern __shared__ __align__(16) unsigned char jsm[];
  int* hk = reinterpret_cast<int*>(jsm);                                   // [JOB_HC] node id or -1
  uint32_t* hc = reinterpret_cast<uint32_t*>(jsm + JOB_HC * 4);            // [JOB_HC] pixel count
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(jsm + JOB_HC * 8);   // [<= JOB_HC] sort keys
  double* redd = reinterpret_cast<double*>(jsm + JOB_HC * 8 + JOB_SC * 8);  // [3][16] min bound partials
  unsigned long long* redu = reinterpret_cast<unsigned long long*>(redd + 48);         // [16] depth sums
  int* redi = reinterpret_cast<int*>(redu + 16);                           // [16] counts, then scan partials [512 + 8]
  const int job = blockIdx.x;
  const int fb = job / M, m = job - fb * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = JOB_THREADS / 32;
  if (WRITE && job_off[job + 1] == job_off[job]) return;                  // empty / filtered / padded (block-uniform)
  const int32_t* rc = mask_rect + (long long)job * 4;
  const int x0 = rc[0], y0 = rc[1], x1 = rc[2], y1 = rc[3];
  if (m >= mask_cnt[fb] || x1 <= x0 || y1 <= y0) { if (!WRITE && tid == 0) job_cells[job] = 0; return; }
  for (int i = tid; i < JOB_HC; i += JOB_THREADS) { hk[i] = -1; hc[i] = 0u; }
  __shared__ int s_nent, s_over;
  if (tid == 0) { s_nent = 0; s_over = 0; }
  __syncthreads();
  // ---- phase 1: scan the bounding box, 32 consecutive pixels of a row per warp step
  const int w = x1 - x0, cpr = (w + 31) >> 5, nchunks = cpr * (y1 - y0);
  const int mw = m >> 5, mb = m & 31;
  const unsigned lt = (1u << lane) - 1u;
  int cnt = 0; unsigned long long dsum = 0ull;
  for (int c = warp; c < nchunks; c += NW) {
    const int y = y0 + c / cpr, x = x0 + (c % cpr) * 32 + lane;
    int n = -1; unsigned dep = 0;
    if (x < x1) {
      const long long gp = (long long)fb * HW + (long long)y * W + x;
      if ((__ldg(&maskbits[gp * MW + mw]) >> mb) & 1u) {
        n = pix_idx[gp];
        if (n >= 0) dep = depth[gp];
      }
    }
    const unsigned act = __ballot_sync(0xffffffffu, n >= 0);
    if (!act) continue;
    const unsigned grp = __match_any_sync(0xffffffffu, n);
    cnt += __popc(act);
    dsum += __reduce_add_sync(0xffffffffu, dep);
    if (n >= 0 && (grp & lt) == 0u) {                                    // first lane of a node group
      unsigned h = ((unsigned)n * 2654435761u) >> 18;                    // 14 bits
      bool done = false;
      for (int probe = 0; probe < 256 && !done; probe++) {           // a long probe chain = a table that is filling up: give up early
        int cur = hk[h];
        if (cur == -1) cur = atomicCAS(&hk[h], -1, n);
        if (cur == -1 || cur == n) { atomicAdd(&hc[h], (uint32_t)__popc(grp)); done = true; }
        else h = (h + 1) & (JOB_HC - 1);
      }
      if (!done) s_over = 1;
    }
  }
  if (lane == 0) { redi[warp] = cnt; redu[warp] = dsum; }
  __syncthreads();
  int tcnt = 0; unsigned long long tsum = 0ull;
  for (int i = 0; i < NW; i++) { tcnt += redi[i]; tsum += redu[i]; }      // every thread: same order, same value
  bool keep = tcnt > 0;
  if (keep && ((double)tsum / (double)tcnt) / (double)scale > filter_distance) keep = false;   // generic.py:126-127 (see k_m3d_jobs)
  if (!keep) { if (!WRITE && tid == 0) job_cells[job] = 0; return; }
  __syncthreads();
  // ---- phase 2: compact the table, min bound over the distinct nodes
  double mn[3] = {INFINITY, INFINITY, INFINITY};
  for (int sl = tid; sl < JOB_HC; sl += JOB_THREADS) {
    const int n = hk[sl];
    if (n >= 0) {
      const int pos = atomicAdd(&s_nent, 1);
      if (pos < JOB_SC) sk[pos] = (unsigned long long)sl;
#pragma unroll
      for (int k = 0; k < 3; k++) mn[k] = fmin(mn[k], nodes[(long long)n * 3 + k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    for (int o = 16; o > 0; o >>= 1) mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
    if (lane == 0) redd[k * 16 + warp] = mn[k];
  }
  __syncthreads();
  const int n_ent = s_nent;
  if (s_over || n_ent > JOB_SC) {                                        // does not fit on chip: the batch falls back to the global path
    if (tid == 0) { atomicExch(&counters[1], 1); if (!WRITE) job_cells[job] = 0; }
    return;
  }
  double vmin[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double v = redd[k * 16];
    for (int i = 1; i < NW; i++) v = fmin(v, redd[k * 16 + i]);
    vmin[k] = __dsub_rn(v, __dmul_rn(vs, 0.5));                          // Open3D: min_bound - voxel_size * 0.5
  }
  // ---- phase 3: voxel keys, bitonic sort of (voxel << 14 | slot)
  int P2 = 32;
  while (P2 < n_ent) P2 <<= 1;
  for (int e = tid; e < P2; e += JOB_THREADS) {
    unsigned long long key = ~0ull;
    if (e < n_ent) {
      const int sl = (int)sk[e];
      const long long n = hk[sl];
      key = 0ull;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        long long cc = (long long)floor(cell_coord(nodes[n * 3 + k], vmin[k], vs));
        if (cc < 0 || cc > 16383) { atomicExch(&counters[2], 1); cc = 0; }
        key |= (unsigned long long)cc << (28 - 14 * k);
      }
      key = (key << 14) | (unsigned long long)sl;
    }
    sk[e] = key;
  }
  __syncthreads();
  for (int k2 = 2; k2 <= P2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += JOB_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sk[i], b = sk[ixj];
          const bool up = (i & k2) == 0;
          if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- phase 4: voxel runs -> ranks (block scan over per-thread chunks)
  const int per = (n_ent + JOB_THREADS - 1) / JOB_THREADS;
  const int e0 = min(tid * per, n_ent), e1 = min(e0 + per, n_ent);
  int local = 0;
  for (int e = e0; e < e1; e++) local += (e == 0 || (sk[e] >> 14) != (sk[e - 1] >> 14)) ? 1 : 0;
  int incl = local;
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) redi[warp] = incl;
  __syncthreads();
  int wbase = 0, total = 0;
  for (int i = 0; i < NW; i++) { const int v = redi[i]; if (i < warp) wbase += v; total += v; }
  if (!WRITE) { if (tid == 0) job_cells[job] = total; return; }
  // ---- phase 5: one voxel per run head: nodes visited in ascending id (deterministic), sum += k_n * c_n, mean
  const long long obase = job_off[job];
  int rank = wbase + incl - local;
  for (int e = e0; e < e1; e++) {
    const unsigned long long vkey = sk[e] >> 14;
    if (!(e == 0 || vkey != (sk[e - 1] >> 14))) continue;
    int q1 = e + 1;
    while (q1 < n_ent && (sk[q1] >> 14) == vkey) q1++;
    double sacc[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long tot = 0;
    int last = -1;
    for (int it = e; it < q1; it++) {
      int best = 0x7FFFFFFF; uint32_t bc = 0;
      for (int q = e; q < q1; q++) {
        const int sl = (int)(sk[q] & 0x3FFFull);
        const int nd = hk[sl];
        if (nd > last && nd < best) { best = nd; bc = hc[sl]; }
      }
      last = best;
      const double kd = (double)bc;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        sacc[k] = __dadd_rn(sacc[k], __dmul_rn(kd, nodes[(long long)best * 3 + k]));
        sacc[3 + k] = __dadd_rn(sacc[3 + k], __dmul_rn(kd, nrgb[(long long)best * 3 + k]));
      }
      tot += bc;
    }
    const long long v = obase + rank;
    const double t = (double)tot;
#pragma unroll
    for (int k = 0; k < 3; k++) { oxyz[v * 3 + k] = __ddiv_rn(sacc[k], t); orgb[v * 3 + k] = __ddiv_rn(sacc[3 + k], t); }
    oijk[v * 3] = (int32_t)((vkey >> 28) & 0x3FFF); oijk[v * 3 + 1] = (int32_t)((vkey >> 14) & 0x3FFF); oijk[v * 3 + 2] = (int32_t)(vkey & 0x3FFF);
    rank++;
  }
}

__global__ void k_m3d_off64(const int* __restrict__ cells, int n_jobs, const int* __restrict__ scan, long long* __restrict__ off) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_jobs) off[j] = scan[j];
  if (j == n_jobs) off[j] = n_jobs > 0 ? (long long)scan[n_jobs - 1] + cells[n_jobs - 1] : 0;
}

// ------------------------------------------------------------------------------------------------ host
static inline unsigned m3d_blocks(long long n) { return (unsigned)((n + MTPB - 1) / MTPB); }

static M3dState* m3d_state(hmsg_ctx* ctx) {
  if (!ctx->m3d) ctx->m3d = new M3dState();
  return ctx->m3d;
}

int32_t masks3d_destroy(hmsg_ctx* ctx) {
  M3dState* st = ctx->m3d;
  if (!st) return HMSG_OK;
  for (auto& c : st->chunks) if (c.base) cudaFree(c.base);
  free_dev(st->hkeys); free_dev(st->hcnt); free_dev(st->entry_slot); free_dev(st->counters); free_dev(st->jobs); free_dev(st->skeys);
  free_dev(st->svals); free_dev(st->heads); free_dev(st->hscan); free_dev(st->tmp); free_dev(st->sout);
  delete st;
  ctx->m3d = nullptr;
  return HMSG_OK;
}

// the pixel -> node map of another mask batch invalidates the scratch result
void masks3d_invalidate_scratch(hmsg_ctx* ctx) {
  if (ctx->m3d) ctx->m3d->scratch_valid = false;
}

static int32_t m3d_chunk_alloc(hmsg_ctx* ctx, M3dState* st, size_t bytes, char** out) {
  bytes = (bytes + 255) & ~(size_t)255;
  while (st->cur < st->chunks.size()) {
    M3dChunk& c = st->chunks[st->cur];
    if (c.used + bytes <= c.cap) { *out = c.base + c.used; c.used += bytes; return HMSG_OK; }
    st->cur++;
    if (st->cur < st->chunks.size()) st->chunks[st->cur].used = 0;
  }
  M3dChunk c;
  c.cap = std::max(bytes, (size_t)1 << 30);
  cudaError_t e = cudaMalloc((void**)&c.base, c.cap);
  if (e != cudaSuccess) return ctx->fail(HMSG_ERR_CUDA, std::string("mask store: cudaMalloc: ") + cudaGetErrorString(e));
  c.used = bytes;
  st->chunks.push_back(c);
  st->cur = st->chunks.size() - 1;
  *out = c.base;
  return HMSG_OK;
}

static int g_m3d_path = 1;   // hmsg_set_option("mask3d_path", 1: global hash + radix sort (default) | 0: per-job on-chip kernels, masks with <= 8 k nodes)
int32_t masks3d_set_option(hmsg_ctx*, const char* key, int value) {
  if (!strcmp(key, "mask3d_path")) { g_m3d_path = value; return HMSG_OK; }
  return -1;
}

static int32_t m3d_alloc_out(hmsg_ctx* ctx, M3dState* st, bool store, long long cap_pts, int n_jobs, M3dBatchRec& rec) {
  int32_t rc;
  rec.cap_pts = std::max<long long>(cap_pts, 1);
  const size_t pts_bytes = (((size_t)rec.cap_pts * 24) + 255) & ~(size_t)255, ijk_bytes = (((size_t)rec.cap_pts * 12) + 255) & ~(size_t)255;
  const size_t off_bytes = (size_t)(n_jobs + 1) * 8;
  const size_t out_bytes = 2 * pts_bytes + ijk_bytes + off_bytes;
  char* out = nullptr;
  if (store) { if ((rc = m3d_chunk_alloc(ctx, st, out_bytes, &out))) return rc; }
  else { if ((rc = ctx->reserve(&st->sout, &st->sout_bytes, out_bytes))) return rc; out = st->sout; }
  rec.xyz = (double*)out;
  rec.rgb = (double*)(out + pts_bytes);
  rec.ijk = (int32_t*)(out + 2 * pts_bytes);
  rec.d_off = (long long*)((char*)rec.ijk + ijk_bytes);
  return HMSG_OK;
}

// per-job on-chip path; *fallback = true when a mask did not fit on chip (nothing was written)
static int32_t m3d_run_jobs(hmsg_ctx* ctx, M3dState* st, int64_t frame_begin, int n_frames, double down_size, double filter_distance, bool store,
                            M3dBatchRec& rec, bool* fallback) {
  int32_t rc;
  *fallback = false;
  const int M = ctx->batch_M, MW = ctx->batch_MW, HW = ctx->cam.H * ctx->cam.W;
  const int fb0 = (int)(frame_begin - ctx->batch_begin);
  const int n_jobs = n_frames * M;
  static bool attr = false;
  if (!attr) {
    HMSG_CUDA(cudaFuncSetAttribute(k_m3d_job<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, JOB_SMEM));
    HMSG_CUDA(cudaFuncSetAttribute(k_m3d_job<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, JOB_SMEM));
    attr = true;
  }
  // job scratch: cells int[n+1] | scan int[n+1] | off ll[n+1]
  size_t jb = (size_t)(n_jobs + 2) * 16 + 64;
  if ((rc = ctx->reserve(&st->jobs, &st->jobs_bytes, jb))) return rc;
  long long* d_off = (long long*)st->jobs;
  int* job_cells = (int*)(d_off + n_jobs + 2);
  int* job_scan = job_cells + n_jobs + 2;
  if (!st->counters) HMSG_CUDA(cudaMalloc((void**)&st->counters, 16));
  const int32_t* pidx = ctx->pix_idx + (size_t)fb0 * HW;
  const uint32_t* mbits = ctx->maskbits + (size_t)fb0 * HW * MW;
  const uint16_t* depth = ctx->depth + (size_t)frame_begin * HW;
  const int32_t* mcnt = ctx->mask_cnt + fb0;
  const int32_t* mrect = ctx->mask_rect + (size_t)fb0 * M * 4;
  ctx->wait_frames(frame_begin, n_frames);
  ctx->prof_begin(PROF_MASK3D);
  HMSG_CUDA(cudaMemsetAsync(st->counters, 0, 16, ctx->stream));
  k_m3d_job<false><<<n_jobs, JOB_THREADS, JOB_SMEM, ctx->stream>>>(pidx, mbits, depth, HW, ctx->cam.W, M, MW, mcnt, mrect, ctx->node_xyz, ctx->node_rgb,
                                                                  down_size, ctx->cam.scale, filter_distance, job_cells, nullptr, st->counters, nullptr,
                                                                  nullptr, nullptr);
  HMSG_LAUNCH_CHECK();
  size_t t0 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, t0, job_cells, job_scan, n_jobs, ctx->stream);
  if ((rc = ctx->reserve(&st->tmp, &st->tmp_bytes, t0))) return rc;
  HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->tmp, t0, job_cells, job_scan, n_jobs, ctx->stream));
  k_m3d_off64<<<m3d_blocks(n_jobs + 1), MTPB, 0, ctx->stream>>>(job_cells, n_jobs, job_scan, d_off);
  HMSG_LAUNCH_CHECK();
  ctx->launches += 2;
  int h[4] = {0, 0, 0, 0};
  long long total = 0;
  HMSG_CUDA(cudaMemcpyAsync(h, st->counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaMemcpyAsync(&total, d_off + n_jobs, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h[1]) { *fallback = true; ctx->prof_end(PROF_MASK3D, 0.0); return HMSG_OK; }
  rec.frame_begin = frame_begin; rec.n_frames = n_frames; rec.M = M;
  rec.counts.assign(ctx->batch_counts.begin() + fb0, ctx->batch_counts.begin() + fb0 + n_frames);
  rec.h_off.clear();
  if ((rc = m3d_alloc_out(ctx, st, store, total, n_jobs, rec))) return rc;
  HMSG_CUDA(cudaMemcpyAsync(rec.d_off, d_off, (size_t)(n_jobs + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  if (total > 0) {
    k_m3d_job<true><<<n_jobs, JOB_THREADS, JOB_SMEM, ctx->stream>>>(pidx, mbits, depth, HW, ctx->cam.W, M, MW, mcnt, mrect, ctx->node_xyz, ctx->node_rgb,
                                                                   down_size, ctx->cam.scale, filter_distance, job_cells, rec.d_off, st->counters, rec.xyz,
                                                                   rec.rgb, rec.ijk);
    HMSG_LAUNCH_CHECK();
  }
  ctx->prof_end(PROF_MASK3D, 0.0);
  return HMSG_OK;
}

static int32_t m3d_run_global(hmsg_ctx* ctx, M3dState* st, int64_t frame_begin, int n_frames, double down_size, double filter_distance, bool store,
                              M3dBatchRec& rec);

static int32_t m3d_run(hmsg_ctx* ctx, M3dState* st, int64_t frame_begin, int n_frames, double down_size, double filter_distance, bool store,
                       M3dBatchRec& rec) {
  int32_t rc;
  if ((rc = features_ensure_pix_idx(ctx))) return rc;
  if (n_frames * ctx->batch_M >= (1 << 20)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_mask_nodes_batch: more than 2^20 (frame, mask) pairs in one call");
  if (ctx->n_nodes >= (1LL << 31)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_mask_nodes_batch: more than 2^31 nodes");
  if (g_m3d_path == 0) {
    bool fallback = false;
    if ((rc = m3d_run_jobs(ctx, st, frame_begin, n_frames, down_size, filter_distance, store, rec, &fallback))) return rc;
    if (!fallback) return HMSG_OK;
  }
  return m3d_run_global(ctx, st, frame_begin, n_frames, down_size, filter_distance, store, rec);
}

static int32_t m3d_run_global(hmsg_ctx* ctx, M3dState* st, int64_t frame_begin, int n_frames, double down_size, double filter_distance, bool store,
                              M3dBatchRec& rec) {
  int32_t rc;
  const int M = ctx->batch_M, MW = ctx->batch_MW, HW = ctx->cam.H * ctx->cam.W;
  const int fb0 = (int)(frame_begin - ctx->batch_begin);
  const int n_jobs = n_frames * M;
  // job arrays: dsum ull | mn ll[3] | cnt int | nent int[+1] | eoff int[+1] | cursor int | keep u8
  size_t jb = (size_t)(n_jobs + 2) * (8 + 24 + 4 * 4 + 1) + 64;
  if ((rc = ctx->reserve(&st->jobs, &st->jobs_bytes, jb))) return rc;
  unsigned long long* job_dsum = (unsigned long long*)st->jobs;
  long long* job_mn = (long long*)(job_dsum + n_jobs + 2);
  int* job_cnt = (int*)(job_mn + 3 * (size_t)(n_jobs + 2));
  int* job_nent = job_cnt + (n_jobs + 2);
  unsigned char* job_keep = (unsigned char*)(job_nent + 3 * (size_t)(n_jobs + 2));
  if (!st->counters) HMSG_CUDA(cudaMalloc((void**)&st->counters, 16));
  size_t want = 1 << 16;
  while (want < (size_t)n_frames * HW / 2) want <<= 1;
  if (st->hcap < want) st->hcap = want;
  const int32_t* pidx = ctx->pix_idx + (size_t)fb0 * HW;
  const uint32_t* mbits = ctx->maskbits + (size_t)fb0 * HW * MW;
  const uint16_t* depth = ctx->depth + (size_t)frame_begin * HW;
  const int32_t* mcnt = ctx->mask_cnt + fb0;
  int h[4] = {0, 0, 0, 0};
  ctx->wait_frames(frame_begin, n_frames);
  ctx->prof_begin(PROF_MASK3D);
  for (;;) {
    const size_t hcap = st->hcap;
    if ((rc = ctx->reserve(&st->hkeys, &st->hkeys_bytes, hcap * 8))) return rc;
    if ((rc = ctx->reserve(&st->hcnt, &st->hcnt_bytes, hcap * 4))) return rc;
    if ((rc = ctx->reserve(&st->entry_slot, &st->entry_slot_bytes, hcap / 2 * 4))) return rc;
    k_m3d_init<<<m3d_blocks((long long)std::max<size_t>(hcap, (size_t)n_jobs)), MTPB, 0, ctx->stream>>>(n_jobs, job_cnt, job_nent, job_dsum, job_mn, st->hkeys,
                                                                                                       st->hcnt, hcap, st->counters);
    dim3 grid(n_frames, (HW + MTPB - 1) / MTPB);
    k_m3d_scan<<<grid, MTPB, 0, ctx->stream>>>(pidx, mbits, depth, HW, M, MW, mcnt, ctx->node_xyz, st->hkeys, st->hcnt, (unsigned long long)hcap - 1,
                                               st->entry_slot, (uint32_t)(hcap / 2), st->counters, job_cnt, job_nent, job_dsum, job_mn);
    k_m3d_compact<<<m3d_blocks((long long)(hcap + 3) / 4), MTPB, 0, ctx->stream>>>(st->hkeys, hcap, st->entry_slot, (uint32_t)(hcap / 2), st->counters);
    k_m3d_jobs<<<m3d_blocks(n_jobs), MTPB, 0, ctx->stream>>>(n_jobs, job_cnt, job_dsum, ctx->cam.scale, filter_distance, job_keep);
    HMSG_LAUNCH_CHECK();
    ctx->launches += 3;
    HMSG_CUDA(cudaMemcpyAsync(h, st->counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
    HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!h[1]) break;
    if (hcap >= ((size_t)1 << 31)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_mask_nodes_batch: (mask, node) table overflow");
    st->hcap = hcap * 4;                                      // rare: many distinct nodes per pixel (far, oblique surfaces)
  }
  const int n_entries = h[0];
  rec.frame_begin = frame_begin; rec.n_frames = n_frames; rec.M = M;
  rec.counts.assign(ctx->batch_counts.begin() + fb0, ctx->batch_counts.begin() + fb0 + n_frames);
  rec.h_off.clear();
  if ((rc = m3d_alloc_out(ctx, st, store, n_entries, n_jobs, rec))) return rc;
  const size_t off_bytes = (size_t)(n_jobs + 1) * 8;
  if (n_entries == 0) {
    HMSG_CUDA(cudaMemsetAsync(rec.d_off, 0, off_bytes, ctx->stream));
    ctx->prof_end(PROF_MASK3D, 0.0);
    return HMSG_OK;
  }
  if ((rc = ctx->reserve(&st->skeys, &st->skeys_bytes, (size_t)n_entries * 16))) return rc;
  if ((rc = ctx->reserve(&st->svals, &st->svals_bytes, (size_t)n_entries * 8))) return rc;
  if ((rc = ctx->reserve(&st->heads, &st->heads_bytes, (size_t)n_entries * 4))) return rc;
  if ((rc = ctx->reserve(&st->hscan, &st->hscan_bytes, (size_t)n_entries * 4))) return rc;
  unsigned long long* k0 = st->skeys; unsigned long long* k1 = st->skeys + n_entries;
  int* v0 = st->svals; int* v1 = st->svals + n_entries;
  // key = job << 42 | voxel; one radix sort over the used bits (cub::DeviceSegmentedSort with one segment per job was
  // measured 1.8x slower here: ~2 k segments of ~2.4 k keys)
  size_t t1 = 0, t2 = 0;
  int end_bit = 42;
  while (end_bit < 63 && ((unsigned long long)n_jobs >> (end_bit - 42)) != 0) end_bit++;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, k0, k1, v0, v1, n_entries, 0, end_bit, ctx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, st->heads, st->hscan, n_entries, ctx->stream);
  if ((rc = ctx->reserve(&st->tmp, &st->tmp_bytes, std::max(t1, t2)))) return rc;
  k_m3d_place<<<m3d_blocks(n_entries), MTPB, 0, ctx->stream>>>(n_entries, st->entry_slot, st->hkeys, job_mn, ctx->node_xyz, down_size, k0, v0, st->counters);
  HMSG_CUDA(cub::DeviceRadixSort::SortPairs(st->tmp, t1, k0, k1, v0, v1, n_entries, 0, end_bit, ctx->stream));
  k_m3d_heads<<<m3d_blocks(n_entries), MTPB, 0, ctx->stream>>>(k1, n_entries, job_keep, st->heads);
  HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->tmp, t2, st->heads, st->hscan, n_entries, ctx->stream));
  k_m3d_means<<<m3d_blocks(n_entries), MTPB, 0, ctx->stream>>>(k1, v1, st->heads, st->hscan, n_entries, st->entry_slot, st->hkeys, st->hcnt, ctx->node_xyz,
                                                              ctx->node_rgb, rec.xyz, rec.rgb, rec.ijk);
  k_m3d_offsets<<<m3d_blocks(n_jobs + 1), MTPB, 0, ctx->stream>>>(k1, st->heads, st->hscan, n_entries, n_jobs, rec.d_off);
  HMSG_LAUNCH_CHECK();
  ctx->launches += 5;
  ctx->prof_end(PROF_MASK3D, 0.0);
  return HMSG_OK;
}

static int32_t m3d_host_offsets(hmsg_ctx* ctx, M3dBatchRec& rec, bool check_overflow) {
  if (!rec.h_off.empty()) return HMSG_OK;
  const int n_jobs = rec.n_frames * rec.M;
  rec.h_off.resize(n_jobs + 1);
  int h[4] = {0, 0, 0, 0};
  HMSG_CUDA(cudaMemcpyAsync(rec.h_off.data(), rec.d_off, (size_t)(n_jobs + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (check_overflow) HMSG_CUDA(cudaMemcpyAsync(h, ctx->m3d->counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
  HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h[2]) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_mask_nodes: a mask spans more than 2^14 voxels per axis");
  return HMSG_OK;
}

extern "C" int32_t hmsg_mask_store_reset(hmsg_ctx* ctx) {
  if (!ctx) return HMSG_ERR_ARG;
  M3dState* st = m3d_state(ctx);
  st->batches.clear();
  st->cur = 0;
  if (!st->chunks.empty()) st->chunks[0].used = 0;
  return HMSG_OK;
}

extern "C" int32_t hmsg_mask_nodes_batch(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, double down_size, double filter_distance,
                                         int32_t keep) {
  if (!ctx) return HMSG_ERR_ARG;
  if (ctx->batch_begin < 0 || n_frames <= 0 || frame_begin < ctx->batch_begin || frame_begin + n_frames > ctx->batch_begin + ctx->batch_n)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_mask_nodes_batch: frames are not inside the current mask batch (hmsg_masks_*)");
  if (!(down_size > 0)) return ctx->fail(HMSG_ERR_ARG, "hmsg_mask_nodes_batch: bad down_size");
  M3dState* st = m3d_state(ctx);
  int32_t rc;
  if (keep) {
    if (!st->batches.empty() && st->batches.back().frame_begin + st->batches.back().n_frames > frame_begin)
      return ctx->fail(HMSG_ERR_STATE, "hmsg_mask_nodes_batch: stored frames must arrive in ascending order (hmsg_mask_store_reset starts over)");
    M3dBatchRec rec;
    if ((rc = m3d_run(ctx, st, frame_begin, n_frames, down_size, filter_distance, true, rec))) return rc;
    st->batches.push_back(std::move(rec));
    return HMSG_OK;
  }
  st->scratch_valid = false;
  if ((rc = m3d_run(ctx, st, frame_begin, n_frames, down_size, filter_distance, false, st->scratch))) return rc;
  st->scratch_valid = true; st->scratch_down = down_size; st->scratch_filter = filter_distance;
  return HMSG_OK;
}

static int32_t m3d_read_frame(hmsg_ctx* ctx, M3dBatchRec& rec, int64_t frame, int64_t* offsets, double* xyz, double* rgb, int32_t* ijk,
                              int32_t* n_masks_out) {
  int32_t rc;
  if ((rc = m3d_host_offsets(ctx, rec, true))) return rc;
  const int fb = (int)(frame - rec.frame_begin);
  const int nm = rec.counts[fb];
  const long long* o = rec.h_off.data() + (size_t)fb * rec.M;
  if (n_masks_out) *n_masks_out = nm;
  if (offsets) for (int m = 0; m <= nm; m++) offsets[m] = o[m] - o[0];
  const long long tot = o[nm] - o[0];
  if (tot > 0) {
    if (xyz) HMSG_CUDA(cudaMemcpyAsync(xyz, rec.xyz + o[0] * 3, (size_t)tot * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (rgb) HMSG_CUDA(cudaMemcpyAsync(rgb, rec.rgb + o[0] * 3, (size_t)tot * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (ijk) HMSG_CUDA(cudaMemcpyAsync(ijk, rec.ijk + o[0] * 3, (size_t)tot * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (xyz || rgb || ijk) HMSG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HMSG_OK;
}

static M3dBatchRec* m3d_find(M3dState* st, int64_t frame) {
  for (auto& b : st->batches) if (frame >= b.frame_begin && frame < b.frame_begin + b.n_frames) return &b;
  return nullptr;
}

extern "C" int32_t hmsg_mask_store_read(hmsg_ctx* ctx, int64_t frame, int32_t* n_masks, int64_t* offsets, double* xyz, double* rgb, int32_t* ijk) {
  if (!ctx) return HMSG_ERR_ARG;
  M3dState* st = m3d_state(ctx);
  M3dBatchRec* rec = m3d_find(st, frame);
  if (!rec) return ctx->fail(HMSG_ERR_STATE, "hmsg_mask_store_read: frame is not in the mask store (hmsg_mask_nodes_batch with keep=1)");
  return m3d_read_frame(ctx, *rec, frame, offsets, xyz, rgb, ijk, n_masks);
}

extern "C" int32_t hmsg_mask_store_count(hmsg_ctx* ctx, int64_t* n_frames, int64_t* n_masks, int64_t* n_points) {
  if (!ctx) return HMSG_ERR_ARG;
  M3dState* st = m3d_state(ctx);
  int64_t nf = 0, nm = 0, np = 0;
  for (auto& b : st->batches) {
    int32_t rc = m3d_host_offsets(ctx, b, false);
    if (rc) return rc;
    nf += b.n_frames;
    for (int f = 0; f < b.n_frames; f++) {
      nm += b.counts[f];
      np += b.h_off[(size_t)f * b.M + b.counts[f]] - b.h_off[(size_t)f * b.M];
    }
  }
  if (n_frames) *n_frames = nf;
  if (n_masks) *n_masks = nm;
  if (n_points) *n_points = np;
  return HMSG_OK;
}

// A7 for one frame of the current mask batch, host outputs (API form of create_3d_masks).  The whole current batch
// is processed once on the device and cached, so asking for its frames one after another costs one pass.
extern "C" int32_t hmsg_mask_nodes(hmsg_ctx* ctx, int64_t frame, double down_size, int64_t* offsets, double* xyz, double* rgb, int32_t* ijk) {
  if (!ctx) return HMSG_ERR_ARG;
  if (ctx->batch_begin < 0 || frame < ctx->batch_begin || frame >= ctx->batch_begin + ctx->batch_n)
    return ctx->fail(HMSG_ERR_STATE, "hmsg_mask_nodes: frame is not in the current mask batch (hmsg_masks_*)");
  if (!offsets || !(down_size > 0)) return ctx->fail(HMSG_ERR_ARG, "hmsg_mask_nodes: bad argument");
  M3dState* st = m3d_state(ctx);
  int32_t rc;
  const double inf = INFINITY;
  if (!st->scratch_valid || st->scratch.frame_begin != ctx->batch_begin || st->scratch.n_frames != ctx->batch_n || st->scratch_down != down_size ||
      st->scratch_filter != inf) {
    st->scratch_valid = false;
    if ((rc = m3d_run(ctx, st, ctx->batch_begin, ctx->batch_n, down_size, inf, false, st->scratch))) return rc;
    st->scratch_valid = true; st->scratch_down = down_size; st->scratch_filter = inf;
  }
  // the ABI of this call returns M+1 offsets (padded slots are empty)
  int32_t nm = 0;
  if ((rc = m3d_read_frame(ctx, st->scratch, frame, offsets, xyz, rgb, ijk, &nm))) return rc;
  for (int m = nm + 1; m <= ctx->batch_M; m++) offsets[m] = offsets[nm];
  return HMSG_OK;
}

// seq_merge over the stored frames in order (graph.py:437-442 -> graph_utils.py:1015-1038): every stored frame is
// one `global = merge_3d_masks(global + frame masks)` iteration fed straight from HBM.
extern "C" int32_t hmsg_objects_merge_stored(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames) {
  if (!ctx) return HMSG_ERR_ARG;
  M3dState* st = m3d_state(ctx);
  std::vector<int64_t> off;
  for (int64_t f = frame_begin; f < frame_begin + n_frames; f++) {
    M3dBatchRec* rec = m3d_find(st, f);
    if (!rec) return ctx->fail(HMSG_ERR_STATE, "hmsg_objects_merge_stored: frame " + std::to_string(f) + " is not in the mask store");
    int32_t rc;
    if ((rc = m3d_host_offsets(ctx, *rec, true))) return rc;
    const int fb = (int)(f - rec->frame_begin);
    const int nm = rec->counts[fb];
    const long long* o = rec->h_off.data() + (size_t)fb * rec->M;
    off.resize(nm + 1);
    for (int m = 0; m <= nm; m++) off[m] = o[m] - o[0];
    if ((rc = hmsg_objects_add_masks(ctx, nm, off.data(), rec->xyz + o[0] * 3, rec->rgb + o[0] * 3, 1))) return rc;
  }
  return HMSG_OK;
}
