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
//   k_m3d_place     (job, node) entries are bucketed by job (per-job entry counts come out of the scan, one exclusive
//                   scan gives the bucket offsets) and get their voxel key floor((c - (min - vs/2)) / vs) computed with
//                   the reference's float64 ops
//   segmented sort  one segment per job (a few thousand 64-bit keys each: on-chip block sorts instead of a 53-bit
//                   global radix sort)  => the canonical ascending (job, i, j, k) output order of the oracle (H2)
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
#include <cub/device/device_segmented_sort.cuh>
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

// one thread per pixel; grid = (ceil(HW / MTPB), n_frames)
__global__ void __launch_bounds__(MTPB) k_m3d_scan(const int32_t* __restrict__ pix_idx, const uint32_t* __restrict__ maskbits,
                                                   const uint16_t* __restrict__ depth, int HW, int M, int MW, const int32_t* __restrict__ mask_cnt,
                                                   const double* __restrict__ nodes, unsigned long long* __restrict__ hkeys,
                                                   uint32_t* __restrict__ hcnt, unsigned long long hmask, uint32_t* __restrict__ entry_slot,
                                                   uint32_t entry_cap, int* __restrict__ counters, int* __restrict__ job_cnt, int* __restrict__ job_nent,
                                                   unsigned long long* __restrict__ job_dsum, long long* __restrict__ job_mn) {
  const int fb = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
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
              if (cur == M3D_EMPTY) {
                const uint32_t e = (uint32_t)atomicAdd(&counters[0], 1);
                if (e < entry_cap) entry_slot[e] = (uint32_t)h; else atomicExch(&counters[1], 1);
                atomicAdd(&job_nent[job], 1);
                cur = key;
              }
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

// entry e -> bucket of its job (order inside a bucket is arbitrary: the segmented sort follows), key = job << 42 | i << 28 | j << 14 | k
__global__ void __launch_bounds__(MTPB) k_m3d_place(int n_entries, const uint32_t* __restrict__ entry_slot, const unsigned long long* __restrict__ hkeys,
                                                    const int* __restrict__ job_eoff, int* __restrict__ job_cursor, const long long* __restrict__ job_mn,
                                                    const double* __restrict__ nodes, double vs, unsigned long long* __restrict__ skeys,
                                                    int* __restrict__ svals, int* __restrict__ counters) {
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
  const int pos = job_eoff[job] + atomicAdd(&job_cursor[job], 1);
  skeys[pos] = key;
  svals[pos] = e;
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

static int32_t m3d_run(hmsg_ctx* ctx, M3dState* st, int64_t frame_begin, int n_frames, double down_size, double filter_distance, bool store,
                       M3dBatchRec& rec) {
  int32_t rc;
  if ((rc = features_ensure_pix_idx(ctx))) return rc;
  const int M = ctx->batch_M, MW = ctx->batch_MW, HW = ctx->cam.H * ctx->cam.W;
  const int fb0 = (int)(frame_begin - ctx->batch_begin);
  const int n_jobs = n_frames * M;
  if (n_jobs >= (1 << 20)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_mask_nodes_batch: more than 2^20 (frame, mask) pairs in one call");
  if (ctx->n_nodes >= (1LL << 31)) return ctx->fail(HMSG_ERR_CAPACITY, "hmsg_mask_nodes_batch: more than 2^31 nodes");
  // job arrays: dsum ull | mn ll[3] | cnt int | nent int[+1] | eoff int[+1] | cursor int | keep u8
  size_t jb = (size_t)(n_jobs + 2) * (8 + 24 + 4 * 4 + 1) + 64;
  if ((rc = ctx->reserve(&st->jobs, &st->jobs_bytes, jb))) return rc;
  unsigned long long* job_dsum = (unsigned long long*)st->jobs;
  long long* job_mn = (long long*)(job_dsum + n_jobs + 2);
  int* job_cnt = (int*)(job_mn + 3 * (size_t)(n_jobs + 2));
  int* job_nent = job_cnt + (n_jobs + 2);
  int* job_eoff = job_nent + (n_jobs + 2);
  int* job_cursor = job_eoff + (n_jobs + 2);
  unsigned char* job_keep = (unsigned char*)(job_cursor + (n_jobs + 2));
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
    dim3 grid((HW + MTPB - 1) / MTPB, n_frames);
    k_m3d_scan<<<grid, MTPB, 0, ctx->stream>>>(pidx, mbits, depth, HW, M, MW, mcnt, ctx->node_xyz, st->hkeys, st->hcnt, (unsigned long long)hcap - 1,
                                               st->entry_slot, (uint32_t)(hcap / 2), st->counters, job_cnt, job_nent, job_dsum, job_mn);
    k_m3d_jobs<<<m3d_blocks(n_jobs), MTPB, 0, ctx->stream>>>(n_jobs, job_cnt, job_dsum, ctx->cam.scale, filter_distance, job_keep);
    HMSG_LAUNCH_CHECK();
    ctx->launches += 2;
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
  rec.cap_pts = std::max(n_entries, 1);
  const size_t pts_bytes = (size_t)rec.cap_pts * 24, ijk_bytes = (((size_t)rec.cap_pts * 12) + 255) & ~(size_t)255, off_bytes = (size_t)(n_jobs + 1) * 8;
  const size_t out_bytes = 2 * ((pts_bytes + 255) & ~(size_t)255) + ijk_bytes + off_bytes;
  char* out = nullptr;
  if (store) { if ((rc = m3d_chunk_alloc(ctx, st, out_bytes, &out))) return rc; }
  else { if ((rc = ctx->reserve(&st->sout, &st->sout_bytes, out_bytes))) return rc; out = st->sout; }
  rec.xyz = (double*)out;
  rec.rgb = (double*)(out + ((pts_bytes + 255) & ~(size_t)255));
  rec.ijk = (int32_t*)(out + 2 * ((pts_bytes + 255) & ~(size_t)255));
  rec.d_off = (long long*)((char*)rec.ijk + ijk_bytes);
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
  // bucket the entries by job, then sort every bucket by voxel key on chip
  size_t t0 = 0, t1 = 0, t2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, t0, job_nent, job_eoff, n_jobs + 1, ctx->stream);
  cub::DeviceSegmentedSort::SortPairs(nullptr, t1, k0, k1, v0, v1, n_entries, n_jobs, job_eoff, job_eoff + 1, ctx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, st->heads, st->hscan, n_entries, ctx->stream);
  if ((rc = ctx->reserve(&st->tmp, &st->tmp_bytes, std::max(t0, std::max(t1, t2))))) return rc;
  HMSG_CUDA(cub::DeviceScan::ExclusiveSum(st->tmp, t0, job_nent, job_eoff, n_jobs + 1, ctx->stream));
  HMSG_CUDA(cudaMemsetAsync(job_cursor, 0, (size_t)n_jobs * 4, ctx->stream));
  k_m3d_place<<<m3d_blocks(n_entries), MTPB, 0, ctx->stream>>>(n_entries, st->entry_slot, st->hkeys, job_eoff, job_cursor, job_mn, ctx->node_xyz, down_size,
                                                              k0, v0, st->counters);
  HMSG_CUDA(cub::DeviceSegmentedSort::SortPairs(st->tmp, t1, k0, k1, v0, v1, n_entries, n_jobs, job_eoff, job_eoff + 1, ctx->stream));
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
