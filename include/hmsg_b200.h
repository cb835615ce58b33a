/*
 * hmsg_b200.h - C-ABI of libhmsg_b200.so: the B200-native (sm_100a) HMSG
 * build-and-retrieve hot path of HorizonRobotics/HoloAgent (FSR-VLN).
 *
 * The reference has no FFI for this path: its boundary is a set of Python call
 * signatures.  Each entry point below names the reference call it
 * replaces (paths relative to /root/reference/fsr_vln/).  INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns int32_t status (HMSG_OK == 0); hmsg_last_error() gives text
 *   - the caller owns every buffer it passes; the library owns all device memory in hmsg_ctx
 *   - `on_device` != 0 means the pointer is a CUDA device pointer on the ctx's device
 *     (inputs already resident in HBM); 0 means host memory (copied inside the call)
 *   - a ctx is single-owner (one host thread at a time), one ctx per GPU, all work is
 *     issued on the ctx-owned CUDA stream; calls that return data to host buffers
 *     synchronise, calls with device outputs are asynchronous until hmsg_sync()
 *   - there is NO CPU fallback: without a CUDA device hmsg_ctx_create fails
 */
#ifndef HMSG_B200_H_
#define HMSG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMSG_OK            0
#define HMSG_ERR_ARG       1
#define HMSG_ERR_CUDA      2
#define HMSG_ERR_STATE     3
#define HMSG_ERR_CAPACITY  4
#define HMSG_ERR_NCCL      5

typedef struct hmsg_ctx hmsg_ctx;

/* ---- context ------------------------------------------------------------------------ */
int32_t     hmsg_ctx_create(int32_t device, hmsg_ctx** out);
int32_t     hmsg_ctx_destroy(hmsg_ctx* ctx);
const char* hmsg_last_error(const hmsg_ctx* ctx);   /* ctx may be NULL: last create error */
int32_t     hmsg_sync(hmsg_ctx* ctx);
/* CUDA stream handle (cudaStream_t) all work is issued on - for event timing by the caller */
void*       hmsg_stream(hmsg_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py "gpu_launches") */
int64_t     hmsg_launch_count(const hmsg_ctx* ctx);
int32_t     hmsg_version(void);
/* tuning / A-B switches (every setting computes the same results; tests compare them):
 *   "gemm_2sm"            0|1  cta_group::2 GEMM (default 1)
 *   "attn_variant"        0 two warps per (image, head) tile (default), 1 / 3 earlier kernels, 2 fp32 reference
 *                         kernel, 4 double-buffered tiles, 5 any-T online-softmax kernel (always used for T > 64)
 *   "last_layer_cls_only" 0|1  last transformer block computes only the class-token row past K/V (default 1)
 *   "crops_mma"           0|1  PIL passes of the mask crops as int8 tensor-core MMAs (default 1) or scalar kernels
 *   "knn_bq"              queries per pass, 0 = auto
 *   "ln_fold"             0|1  encoder: LayerNorms folded into the QKV / FC GEMM epilogues over an fp16 residual stream (default 1)
 *                              or fp32 residual stream + LayerNorm kernels (0); both within 1e-3 of the fp32 oracle
 *   "mask3d_path"         1 global (mask, node) hash + radix sort (default), 0 per-mask on-chip kernels (masks <= 8 k nodes,
 *                              falls back to 1 on overflow) */
int32_t     hmsg_set_option(hmsg_ctx* ctx, const char* key, int32_t value);
/* Per-kernel-class device timing with CUDA events on the ctx stream (bench.py roofline).
 * class: 0 gemm (work = flops), 1 attention, 2 elementwise/LN, 3 knn pass (work = bytes of E
 * streamed), 4 pixel->node + winner, 5 feature scatter, 6 geometry passes, 7 crops, 8 per-mask 3-D node
 * sets (A7), 9 NCCL exchanges (work = bytes this rank sent + received).
 * hmsg_prof_read synchronises, returns the summed event time / launch count / algorithmic
 * work since the last read and resets the class. */
int32_t     hmsg_prof_enable(hmsg_ctx* ctx, uint32_t class_mask);
int32_t     hmsg_prof_read(hmsg_ctx* ctx, int32_t cls, double* ms, int64_t* launches, double* work);

/* ---- scene: resident frame store ---------------------------------------------------- */
/* Graph.__init__ dataset + intrinsics (memory/hmsg/graph/graph.py:203-216;
 * dataloader/generic.py:19-33): depth intrinsics K (row-major 3x3 float64), depth scale
 * (horizon.py:38 -> 1000) and cfg.pipeline.voxel_size (graph.py:349). */
int32_t hmsg_scene_begin(hmsg_ctx* ctx, int32_t H, int32_t W, const double K[9],
                         float depth_scale, double voxel_size, int64_t frame_capacity);
/* dataset[i] -> (rgb, depth, pose) (graph.py:343): uint16 depth [n,H,W], uint8 rgb [n,H,W,3],
 * camera-to-world poses float64 [n,16] row-major.  Frames get ids in arrival order. */
int32_t hmsg_scene_add_frames(hmsg_ctx* ctx, const uint16_t* depth, const uint8_t* rgb,
                              const double* poses, int32_t n_frames, int32_t on_device);
/* same, but at explicit frame ids [frame_begin, frame_begin+n) (a rank that owns a subset of the
 * frames under frame-batch sharding only uploads its own); asynchronous on the ctx stream */
int32_t hmsg_scene_put_frames(hmsg_ctx* ctx, int64_t frame_begin, const uint16_t* depth,
                              const uint8_t* rgb, const double* poses, int32_t n_frames,
                              int32_t on_device);
int32_t hmsg_scene_set_num_frames(hmsg_ctx* ctx, int64_t n_frames);
/* replace ONLY the colour image of stored frames; depth, poses and everything built from them (voxel / node table, pixel->node
 * indices) stay valid.  For datasets whose rgb and depth sizes differ the reference uses two different resized images: point
 * colours come from cv2.resize(..., INTER_AREA) inside create_pcd (dataloader/generic.py:98-104), the crops of the feature pass
 * from PIL `rgb_image.resize(depth_image.size)` (bicubic, graph.py:378-379).  Upload the first with hmsg_scene_put_frames, build
 * the geometry, then swap in the second before hmsg_masks_* / hmsg_encode_crops.  rgb uint8 [n,H,W,3]; ctx stream */
int32_t hmsg_scene_put_rgb(hmsg_ctx* ctx, int64_t frame_begin, const uint8_t* rgb, int32_t n_frames, int32_t on_device);
/* per-frame depth intrinsics (dataloader/iphone.py:290-367 overrides create_pcd with `K = frames[image_id - 1]["K"]`): K float64
 * [n,9] row-major HOST for the frames [frame_begin, frame_begin + n); frames without an entry keep the K of hmsg_scene_begin */
int32_t hmsg_scene_set_intrinsics(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, const double* K);
int64_t hmsg_scene_num_frames(const hmsg_ctx* ctx);
/* forget the stored frames (capacity and intrinsics stay): the next add starts at frame id 0 */
int32_t hmsg_scene_reset_frames(hmsg_ctx* ctx);

/* A1  RGBDDataset.create_pcd (dataloader/generic.py:74-138).  Dense outputs in row-major
 * pixel order: xyz [H*W,3] world float64, rgb [H*W,3] float64 (= u8/255.0), valid [H*W]
 * (depth>0).  Rows of invalid pixels are zero.  The caller compacts (Python shim does). */
int32_t hmsg_unproject_frame(hmsg_ctx* ctx, int64_t frame, double* xyz, double* rgb,
                             uint8_t* valid);

/* A2  `full_pcd += create_pcd(...)` over all frames + full_pcd.voxel_down_sample(voxel_size)
 * (graph.py:339-348, Open3D 0.18.0 VoxelDownSample semantics).  Builds the voxel table in
 * canonical ascending (i,j,k) order.  min_bound_out [3] = min over all points (nullable). */
int32_t hmsg_voxel_build(hmsg_ctx* ctx, int64_t* n_voxels, double* min_bound_out);
/* The same build in stages over frame ranges, so that ranks can split the frames (SURVEY 8e option A):
 *   bounds(range) -> [all-reduce min/max of 6 doubles] -> grid_set -> mark(range) -> [all-gather bitmap,
 *   bitmap_or] -> scan -> accumulate(range) -> [all-reduce sum of acc [n,6] f64 and cnt [n] u32] -> finalize.
 * hmsg_voxel_build is exactly this sequence over all frames on one GPU. */
int32_t hmsg_voxel_bounds(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames, double minmax[6]);
int32_t hmsg_voxel_grid_set(hmsg_ctx* ctx, const double minmax[6]);
int32_t hmsg_voxel_mark(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames);
int32_t hmsg_voxel_bitmap(hmsg_ctx* ctx, uint32_t** bitmap_dev, int64_t* nwords);
int32_t hmsg_voxel_bitmap_or(hmsg_ctx* ctx, const uint32_t* gathered, int32_t world);
int32_t hmsg_voxel_scan(hmsg_ctx* ctx, int64_t* n_voxels);
int32_t hmsg_voxel_accumulate(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames);
int32_t hmsg_voxel_acc(hmsg_ctx* ctx, double** acc_dev, uint32_t** cnt_dev, int64_t* n_voxels);
int32_t hmsg_voxel_finalize(hmsg_ctx* ctx);
int32_t hmsg_voxels_read(hmsg_ctx* ctx, double* xyz, double* rgb, int32_t* ijk,
                         uint32_t* count);           /* any pointer may be NULL */

/* A3  pcd_denoise_dbscan(eps=.01,min_points=100) [identity, SURVEY H6] +
 * remove_radius_outlier(nb_points, radius) + select_by_index (graph.py:352-358).
 * After this call the node table (= Graph.full_pcd) is final. */
int32_t hmsg_radius_filter(hmsg_ctx* ctx, int32_t nb_points, double radius, int64_t* n_nodes);
int32_t hmsg_radius_counts_read(hmsg_ctx* ctx, uint32_t* counts /* [n_voxels] */);
int32_t hmsg_nodes_read(hmsg_ctx* ctx, double* xyz, double* rgb, int32_t* ijk,
                        int64_t* voxel_index);        /* any pointer may be NULL */
int64_t hmsg_num_nodes(const hmsg_ctx* ctx);

/* A4  tree_pcd.query(np.asarray(pcd.points), k=1) (graph.py:362-364, :409; generic.py:181):
 * exact nearest node for every valid pixel.  idx [H*W] (-1 for depth==0), dist [H*W] or NULL. */
int32_t hmsg_pixel_to_node(hmsg_ctx* ctx, int64_t frame, int64_t* idx, double* dist);
/* same query for arbitrary float64 points [n,3] (graph.py:458) */
int32_t hmsg_points_to_node(hmsg_ctx* ctx, const double* xyz, int64_t n, int64_t* idx,
                            double* dist);

/* ---- feature ingest (A5, A6) -------------------------------------------------------- */
/* counter/sum_features = zeros (graph.py:366-368).  d = clip_feat_dim, multiple of 128. */
int32_t hmsg_features_begin(hmsg_ctx* ctx, int32_t d);

/* SAM masks of a batch of consecutive frames ("segmentation" bool [M,H,W] per frame,
 * perception/models/sam_clip_feats_extractor.py:117,184).  Either dense uint8 [n,M,H,W]
 * (0/1) or XYWH rectangles int32 [n,M,4] meaning rect AND (depth>0) (synthetic). */
int32_t hmsg_masks_dense(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, int32_t M,
                         const uint8_t* seg, int32_t on_device);
int32_t hmsg_masks_boxes(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, int32_t M,
                         const int32_t* xywh, int32_t on_device);
/* the same masks as ONE label image per frame: int8 [n,H,W], pixel p belongs to mask labels[p] (0 <= label < M <= 127;
 * other values: no mask) AND (depth > 0) - non-overlapping instance / panoptic masks at 1 byte per pixel */
int32_t hmsg_masks_labels(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, int32_t M,
                          const int8_t* labels, int32_t on_device);

/* Ragged SAM output (extractor.py:117-124 returns as many masks as SAM finds): counts [n_frames] int32 HOST,
 * counts[i] <= M real masks in frame frame_begin + i of the batch just set with hmsg_masks_*.  Slots past the count
 * are padding: they take no part in the softmax of A5 (extractor.py:168-172 runs over the frame's own masks),
 * produce no 3-D mask in A7 and are not appended to the N1 merge list.  Without this call every slot is real. */
int32_t hmsg_masks_counts(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, const int32_t* counts);

/* A5+A6 for the batch whose masks were just set: given the encoder outputs of the batch,
 * feats [n, 2M+1, d] float32 unit rows ordered (M masked crops, M plain crops, 1 full
 * frame) = (cropped_masked_feats, cropped_feats, F_g) of extractor.py:147-158, computes
 * F_p (extractor.py:159-175), the per-pixel map restricted to the pixels that win their
 * node (extractor.py:177-190 incl. .half(); graph.py:404-411 with the last-writer-wins
 * rule) and accumulates sum_features / counter.
 * F_p_out [n,M,d] float32 (device if on_device else host; nullable). */
int32_t hmsg_fuse_scatter(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, int32_t M,
                          const float* feats, float maskedd_weight, float* F_p_out,
                          int32_t on_device);
/* graph.py:413-415: counter[counter==0]=1e-5; full_feats = sum/counter -> [n_nodes,d] f32 */
int32_t hmsg_node_feats_finalize(hmsg_ctx* ctx, float* full_feats, int32_t on_device);
int32_t hmsg_node_feats_raw(hmsg_ctx* ctx, float* sum_features, float* counter);
/* API parity only: the dense per-pixel map `outfeat` [H*W,d] fp16 of extract_feats_per_pixel
 * (extractor.py:177-190) for one frame of the batch last passed to hmsg_fuse_scatter. */
int32_t hmsg_pixel_feature_map(hmsg_ctx* ctx, int64_t frame, uint16_t* out_half);

/* A7  RGBDDataset.create_3d_masks (dataloader/generic.py:140-190) for one frame whose masks
 * were set: per mask the node positions hit by its pixels, re-voxelised (down_size) relative
 * to the mask's own min bound with pixel multiplicity as weight.  Ragged output: offsets
 * [M+1]; xyz/rgb [offsets[M],3] float64; ijk int32.  Call once with NULL data pointers to
 * get offsets, then again with buffers.  The whole current mask batch is processed on the device
 * once and cached; only the requested frame's rows are copied to the host. */
int32_t hmsg_mask_nodes(hmsg_ctx* ctx, int64_t frame, double down_size, int64_t* offsets,
                        double* xyz, double* rgb, int32_t* ijk);
/* A7 for a range of frames inside the current mask batch, results kept in HBM (the ingest path:
 * graph.py:391-402 calls create_3d_masks for every frame and appends to frames_pcd).  filter_distance =
 * cfg.pipeline.max_mask_distance (generic.py:126: a mask whose mean depth exceeds it yields an empty cloud).
 * keep = 1 appends the frames to the mask store (ascending frame order), keep = 0 leaves them in a scratch
 * area overwritten by the next call.  Voxel keys / counts are exact; means are sums of (pixel count x node
 * centroid) in ascending node order, within 1e-12 of Open3D's per-pixel sequential sums. */
int32_t hmsg_mask_nodes_batch(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, double down_size,
                              double filter_distance, int32_t keep);
int32_t hmsg_mask_store_reset(hmsg_ctx* ctx);
int32_t hmsg_mask_store_count(hmsg_ctx* ctx, int64_t* n_frames, int64_t* n_masks, int64_t* n_points);
/* one stored frame -> host: n_masks (real masks of the frame), offsets [n_masks+1], xyz/rgb [offsets[n_masks],3]
 * float64, ijk int32; any output pointer may be NULL */
int32_t hmsg_mask_store_read(hmsg_ctx* ctx, int64_t frame, int32_t* n_masks, int64_t* offsets, double* xyz,
                             double* rgb, int32_t* ijk);

/* ---- A9 encoder: open_clip ViT visual tower ------------------------------------------ */
/* Supported: head dim 64 or 80 (width / heads; 80 = ViT-H-14, graph.py:105-111), width / mlp / out_dim multiples of 256, image a multiple of patch,
 * up to 4096 tokens: ViT-B-32 (graph.py:112-119, clip_feat_dim 512; values in the comments below), ViT-B-16 and the
 * config default ViT-L/14 (graph.py:98-104: image 224, patch 14, width 1024, 24 layers, 16 heads, mlp 4096, out 768). */
typedef struct hmsg_vit_desc {
  int32_t image;      /* 224 */
  int32_t patch;      /* 32  */
  int32_t width;      /* 768 */
  int32_t layers;     /* 12  */
  int32_t heads;      /* 12  */
  int32_t mlp;        /* 3072 */
  int32_t out_dim;    /* 512 */
  int32_t quick_gelu; /* 0: erf GELU (open_clip ViT-B-32), 1: x*sigmoid(1.702x) (openai cfg) */
} hmsg_vit_desc;

/* weights: one float32 host blob in the order documented in DESIGN.md ("encoder blob"),
 * i.e. open_clip VisionTransformer.state_dict() flattened by holoagent_b200.encoder.pack().
 * Values are rounded to fp16 on upload (graph.py:117 precision='fp16'). */
int32_t hmsg_encoder_load(hmsg_ctx* ctx, const hmsg_vit_desc* desc, const float* blob,
                          int64_t blob_floats);
/* clip_model.encode_image(x).float(); F.normalize(dim=-1) (utils/clip_utils.py:75-76, :91-92).
 * x [B,3,image,image] float32 (already preprocessed); out [B,out_dim] float32 unit rows.
 * normalize=0 returns the raw projection. */
int32_t hmsg_encode_images(hmsg_ctx* ctx, const float* nchw, int32_t B, float* out,
                           int32_t normalize, int32_t on_device);
/* debug / unit test: C[M,N] (f32) = A[M,K] (f16 bits) * W[N,K]^T (f16 bits), device ptrs */
int32_t hmsg_gemm_f16_debug(hmsg_ctx* ctx, const void* A, const void* W, float* C,
                            int32_t M, int32_t N, int32_t K);

/* ---- A8 (N3) crops + preprocess on device (bit-exact cv2 bilinear + PIL antialiased bicubic) -------------------------------------------- */
/* crop_all_bounding_boxs x2 + preprocess (utils/sam_utils.py:119-181; clip_utils.py:88-89)
 * for the frames whose masks were set with hmsg_masks_*: writes [n, 2M+1, 3, 224, 224] f32
 * into the ctx crop buffer (returned device pointer) in (masked, plain, full) order. */
int32_t hmsg_make_crops(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, int32_t M,
                        const int32_t* xywh, int32_t bbox_margin, int32_t on_device,
                        float** crops_dev_out);
/* hmsg_make_crops + hmsg_encode_images fused for the ingest path: with a patch-32 tower (ViT-B-32) the crops
 * are resampled straight into the encoder's fp16 patch matrix (no fp32 crop tensor, no im2col pass); other
 * towers go through the fp32 crops internally.  feats_out is a
 * DEVICE pointer [n*(2M+1), out_dim] float32 unit rows in (masked, plain, full) order per frame;
 * xywh is host or device per on_device. */
int32_t hmsg_encode_crops(hmsg_ctx* ctx, int64_t frame_begin, int32_t n_frames, int32_t M,
                          const int32_t* xywh, int32_t bbox_margin, int32_t on_device,
                          float* feats_out);
/* copy the first n_crops [3,224,224] float32 crops of the last hmsg_make_crops to host (tests) */
int32_t hmsg_crops_read(hmsg_ctx* ctx, int64_t n_crops, float* host_out);
/* host-only debug export (no ctx, no GPU): the int8 MMA coefficient fragments of PIL's antialiased
 * bicubic 512 -> 224 pass (Resample.c precompute_coeffs / normalize_coeffs_8bpc as called by the
 * open_clip preprocess, utils/clip_utils.py:88-89).  frag_out [28][3][32][2] uint32 = per 8-output
 * tile the (d2, d1, d0) digit planes in m16n8k32 B-fragment order, x0_out [28] = window starts. */
int32_t hmsg_debug_pil_mma_table(uint32_t* frag_out, int32_t* x0_out);

/* ---- A11 retrieval -------------------------------------------------------------------- */
/* object_embs = np.array([obj.embedding ...]) (graph.py:3126): E [N,d] float32, copied into
 * HBM (on_device 0/1) or borrowed without a copy (on_device 2). */
int32_t hmsg_index_set(hmsg_ctx* ctx, const float* E, int64_t N, int32_t d, int32_t on_device);
/* np.dot(q, E.T); np.argsort(...)[::-1][:k] per query row (graph.py:2196-2200, :3127-3133,
 * :2888-2897).  Q [nq,d]; ids [nq,k] int64; scores [nq,k] float32 (descending; ties -> lower
 * index).  row_mask: optional uint8 [N] device/host like Q (rows with 0 are skipped; the
 * room filter of graph.py:3112-3122).  Any k >= 1: k <= 32 runs the fused matvec + top-k pass,
 * larger k (callers that rank a whole room) a dense-score + key-sort pass; slots past the number
 * of eligible rows hold id -1 / score -inf. */
int32_t hmsg_query_topk(hmsg_ctx* ctx, const float* Q, int32_t nq, int32_t k,
                        const uint8_t* row_mask, int64_t* ids, float* scores,
                        int32_t on_device);
/* dense sim = np.dot(Q, E.T) -> scores [nq,N] for small tables: room names / room view embeddings
 * (graph.py:3204, :3250-3257, :3345-3350) and class labels (identify_object, graph.py:1452). */
int32_t hmsg_query_scores(hmsg_ctx* ctx, const float* Q, int32_t nq, float* scores, int32_t on_device);
/* query_hmsg_object core with negative prompts (graph.py:3134-3151): for each request r the
 * Qp rows Q[r] are (query + negatives); objects whose column-argmax is query_id, sorted by
 * -max score, first k.  n_found[r] = number returned (< k possible; 0 => caller falls back to
 * the plain top-k exactly as the reference keeps `top_index`, graph.py:3133). */
int32_t hmsg_query_object(hmsg_ctx* ctx, const float* Q, int32_t n_req, int32_t Qp,
                          int32_t query_id, int32_t k, const uint8_t* row_mask,
                          int64_t* ids, float* scores, int32_t* n_found, int32_t on_device);

/* ---- N1 object instances: 3-D mask merging across frames ------------------------------- */
/* seq_merge(frames_pcd, init_overlap_thresh, voxel_size, iou_thresh) (graph.py:437-442 ->
 * utils/graph_utils.py:1015-1038): hmsg_objects_begin resets the global mask list; every
 * hmsg_objects_add_masks call is one loop iteration (`global = merge_3d_masks(global + frame
 * masks)`; the first call only stores its masks); hmsg_objects_finish applies the final merge and
 * the small-mask removal of graph.py:444-448 (`is_empty() or len(points) < min_points`).
 * merge_3d_masks = AABB-IoU gate (compute_3d_bbox_iou, float64) -> find_overlapping_ratio_faiss
 * (float32 exact-L2: fraction of points with a neighbour within 1.5*down_size, both ways, max) ->
 * connected components of `ratio > overlap_thresh` -> concat in list order ->
 * pcd_denoise_dbscan(eps 0.1, min_points 10) keeping the largest cluster unless it has < 5 points
 * (graph_utils.py:620-679, :827-956).  A frame's masks are ragged: offsets int64 [n_masks+1] (HOST,
 * offsets[0] = 0), xyz / rgb float64 [offsets[n_masks], 3] host or device per on_device (rgb may be
 * NULL).  All point data stays in HBM between calls. */
int32_t hmsg_objects_begin(hmsg_ctx* ctx, double overlap_thresh, double down_size, double iou_thresh);
int32_t hmsg_objects_add_masks(hmsg_ctx* ctx, int32_t n_masks, const int64_t* offsets,
                               const double* xyz, const double* rgb, int32_t on_device);
/* Same iteration fed from the device: the 3-D masks of `frame` (create_3d_masks, generic.py:141-190,
 * with filter_distance = cfg.pipeline.max_mask_distance) are built from the current mask batch
 * (hmsg_masks_*) and merged without leaving HBM.  Voxel sums run over the mask's pixels in row-major
 * order (stable sort), so the point sets are bit-identical to Open3D's sequential accumulation. */
int32_t hmsg_objects_add_frame(hmsg_ctx* ctx, int64_t frame, double down_size, double filter_distance);
/* The same iterations fed from the mask store (hmsg_mask_nodes_batch with keep = 1): frames
 * [frame_begin, frame_begin + n_frames) in order, nothing visits the host but ragged offsets. */
int32_t hmsg_objects_merge_stored(hmsg_ctx* ctx, int64_t frame_begin, int64_t n_frames);
int32_t hmsg_objects_finish(hmsg_ctx* ctx, int32_t min_points, int64_t* n_objects, int64_t* n_points);
/* current list (after finish: the objects, self.mask_pcds): offsets [n+1], xyz / rgb [n_points,3] -> host */
int32_t hmsg_objects_read(hmsg_ctx* ctx, int64_t* offsets, double* xyz, double* rgb);
int32_t hmsg_objects_count(hmsg_ctx* ctx, int64_t* n_masks, int64_t* n_points, int64_t* gated_pairs);

/* ---- N2 per-object feature ------------------------------------------------------------ */
/* graph.py:451-488 over the objects of hmsg_objects_finish: voxel_down_sample(voxel_size) ->
 * nearest node (tree_pcd.query) -> keep matches with dist <= max_dist (0.8) ->
 * np.nan_to_num(full_feats[idx]) -> feats_denoise_dbscan(eps, min_points) (utils/graph_utils.py:
 * 682-728: sklearn DBSCAN(metric="cosine"), largest cluster mean, or the mean of all rows when
 * nothing clusters; objects without rows get zeros).  full_feats [n_nodes,d] float32 is
 * self.full_feats_array (hmsg_node_feats_finalize); out [n_objects,d] float32 = self.mask_feats.
 * Both host or both device per on_device. */
int32_t hmsg_object_feats(hmsg_ctx* ctx, const float* full_feats, int32_t d, double voxel_size,
                          double max_dist, float eps, int32_t min_points, float* out,
                          int32_t on_device);

/* ---- multi-GPU: frame-batch sharding over the GPUs of one box, NCCL over NVLink ---------- */
/* One ctx (one process or thread) per GPU.  Either the library owns the communicator -
 * hmsg_comm_unique_id on rank 0, the 128 bytes reach the other ranks by any means, hmsg_comm_init on every
 * rank - or the caller supplies its own: every collective entry point takes `void* nccl_comm` (an ncclComm_t;
 * NULL = the ctx's communicator).  NCCL is resolved with dlopen at the first call (the copy already loaded in
 * the process, else HMSG_NCCL_LIB, else libnccl.so.2); failures return HMSG_ERR_NCCL. */
int32_t hmsg_comm_unique_id(uint8_t id_out[128]);
int32_t hmsg_comm_init(hmsg_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world);
int32_t hmsg_comm_attach(hmsg_ctx* ctx, void* nccl_comm);
/* rank / world of the ctx's communicator and the bytes this rank sent + received in the last hmsg_allgather_nodes */
int32_t hmsg_comm_info(hmsg_ctx* ctx, int32_t* rank, int32_t* world, double* last_exchange_bytes);
/* hmsg_voxel_build where this rank only touches its own frame ranges (int64 pairs (begin, count), HOST): local
 * bounds / occupancy / accumulation merged by all-reduce(min) of 6 doubles, all-gather + OR of the bitmap and
 * all-reduce(sum) of the f64 accumulators; every rank ends with the identical voxel table. */
int32_t hmsg_voxel_build_sharded(hmsg_ctx* ctx, void* nccl_comm, const int64_t* ranges, int32_t n_ranges,
                                 int64_t* n_voxels, double* min_bound_out);
/* hmsg_radius_filter with the neighbour counts computed for this rank's slice of the voxel table and exchanged */
int32_t hmsg_radius_filter_sharded(hmsg_ctx* ctx, void* nccl_comm, int32_t nb_points, double radius,
                                   int64_t* n_nodes);
/* The node-embedding merge: after every rank has scattered its own frames, sum_features / counter
 * hold dense partials.  Row slices are exchanged all-to-all, the owner sums the `world` partials of its slice in
 * rank order (deterministic) and the finished slices are all-gathered: every rank ends with the full sums.
 * Optionally the per-frame mask embeddings (frames_feats, graph.py:402) are gathered in the same group:
 * Fp_local (device, fp_stride_floats floats of which fp_floats are used) -> Fp_all [world, fp_stride_floats]
 * (device); pass NULLs to skip. */
int32_t hmsg_allgather_nodes(hmsg_ctx* ctx, void* nccl_comm, const float* Fp_local, int64_t fp_floats,
                             float* Fp_all, int64_t fp_stride_floats);
/* Building blocks of the torch.distributed form of the same merge (holoagent_b200/ingest.py drives either):
 * device views of the partials, pack into one all-gather send buffer, rank-order sum of gathered partials. */
int32_t hmsg_node_feats_device(hmsg_ctx* ctx, float** sum_features, float** counter,
                               int64_t* n_nodes, int32_t* d);
int32_t hmsg_node_feats_pack(hmsg_ctx* ctx, float* dst, const float* Fp_rows, int64_t fp_floats);
int32_t hmsg_node_feats_merge(hmsg_ctx* ctx, const float* gathered, int32_t world,
                              int64_t stride_floats);

#ifdef __cplusplus
}
#endif
#endif /* HMSG_B200_H_ */
