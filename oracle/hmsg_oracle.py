"""CPU oracle for the HMSG build-and-retrieve hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``holoagent_b200``) never does; it fails loudly when the CUDA library is missing.

PARITY PINNING.  The reference (HorizonRobotics/HoloAgent @ 9eddd6e5) has no tests, golden vectors
or fixtures for ``fsr_vln``, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF RUN IN THE BUILD CONTAINER: tests/golden/make_reference_golden.py imports the
unmodified sources from /root/reference/fsr_vln and runs ``Graph.create_feature_map`` (whole method,
graph.py:262-491: create_pcd, extract_feats_per_pixel, get_img_feats*, crop_*, create_3d_masks, the
index_put accumulation, seq_merge / merge_3d_masks / overlap ratio / bbox IoU / dbscan denoise,
feats_denoise_dbscan) and ``Graph.query_hmsg_object`` on a seeded scene; the results are committed as
tests/golden/ref_build.npz / ref_query.npz and tests/test_reference_golden.py holds this oracle to them
(bit-exact node table and object point sets, <= 1e-6 on embeddings / node features / scores).
STILL RESTATED, NOT PINNED (absent from the container and un-vendored): Open3D 0.18.0 internals
(VoxelDownSample, RemoveRadiusOutliers, ClusterDBSCAN, transform - fsr_vln/environment.yaml:16) and
faiss IndexFlatL2 - the fixture run used this file's restatement of exactly those routines behind an
``open3d`` / ``faiss`` shim (tests/golden/ref_shims.py), so for them the fixtures pin the reference's
glue around the call, not the library's arithmetic; open_clip's image_transform is restated with the
same PIL / torch calls (checked against PIL and HF CLIP in tests/test_oracle.py).  SAM and the CLIP
weights are inputs.  Every implementation-defined behaviour is pinned by an explicit rule (H1..H8);
H1 was CONFIRMED by the reference run: ``sum_features[idx] += F_2D`` is thread-count dependent in
torch (13 / 17522 rows differ between 1 and 8 intra-op threads, run-to-run differences at 64); the
single-thread meaning (last pixel wins) is the one recorded and implemented.

Reference files are cited as file:line relative to /root/reference/fsr_vln/.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial import cKDTree

# ----------------------------------------------------------------------------
# A1  depth unprojection + rigid transform
# memory/hmsg/dataloader/generic.py:74-138
# ----------------------------------------------------------------------------

def create_pcd(rgb, depth, K, scale, camera_pose, mask_img=False, filter_distance=np.inf):
    """Restates RGBDDataset.create_pcd (generic.py:94-138).

    rgb   : uint8 [H,W,3]  (or bool [H,W] when mask_img, generic.py:115-116)
    depth : uint16 [H,W]
    Returns (points float64 [n,3] world, colors float64 [n,3] or None, valid bool [H,W]).
    Open3D ``pcd.transform(T)`` (generic.py:137) is restated as
    ``p' = (T[:, :3] * p).sum + T[:, 3]`` evaluated left to right in float64 without
    fused multiply-add, followed by the division by w (H4).
    """
    depth = np.asarray(depth)
    H, W = depth.shape[0], depth.shape[1]
    y, x = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")          # :110  int64
    depth_f = depth.astype(np.float32) / scale                                # :111  float32 quotient
    if mask_img:
        depth_f = depth_f * np.asarray(rgb)                                   # :115-116
    mask = depth_f > 0                                                        # :117
    xs, ys, d = x[mask], y[mask], depth_f[mask]
    X = (xs - K[0, 2]) * d / K[0, 0]                                          # :122  float64
    Y = (ys - K[1, 2]) * d / K[1, 1]                                          # :123
    Z = d                                                                     # :124  float32
    if Z.size and Z.mean() > filter_distance:                                 # :126-127
        return np.zeros((0, 3)), None, mask
    pts = np.hstack((X.reshape(-1, 1), Y.reshape(-1, 1), Z.reshape(-1, 1)))   # :129-130 -> float64
    colors = None
    if not mask_img:
        colors = np.asarray(rgb)[mask] / 255.0                                # :134-135
    return transform_points(pts, camera_pose), colors, mask


def transform_points(pts, camera_pose):
    """Open3D Geometry3D::TransformPoints (generic.py:137): Eigen 4x4 * (x,y,z,1), then / w (H4)."""
    T = np.asarray(camera_pose, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    px, py, pz = pts[:, 0], pts[:, 1], pts[:, 2]
    out = np.empty_like(pts)
    w = ((T[3, 0] * px + T[3, 1] * py) + T[3, 2] * pz) + T[3, 3] * 1.0
    for r in range(3):
        out[:, r] = (((T[r, 0] * px + T[r, 1] * py) + T[r, 2] * pz) + T[r, 3] * 1.0) / w
    return out


# ----------------------------------------------------------------------------
# A2  Open3D 0.18.0 PointCloud::VoxelDownSample (graph.py:348; generic.py:188)
# ----------------------------------------------------------------------------

def voxel_keys(points, voxel_size, min_bound=None):
    """key = floor((p - (min(points) - vs/2)) / vs) per axis, int32 (H3)."""
    if min_bound is None:
        min_bound = points.min(axis=0)
    vmin = min_bound - voxel_size * 0.5
    ref = (points - vmin) / voxel_size
    return np.floor(ref).astype(np.int32), vmin


def voxel_down_sample(points, colors, voxel_size):
    """Returns (pts [n,3], cols [n,3] or None, ijk int32 [n,3], inverse int64 [N]).

    Open3D accumulates each voxel's points sequentially in input order in double
    and divides by the count; np.bincount(weights=...) is the same sequential loop.
    Output order is unordered_map iteration order in Open3D => canonical order here
    is ascending (i,j,k) (H2).
    """
    if len(points) == 0:
        return np.zeros((0, 3)), (None if colors is None else np.zeros((0, 3))), np.zeros((0, 3), np.int32), np.zeros(0, np.int64)
    ijk, _ = voxel_keys(points, voxel_size)
    lin = (ijk[:, 0].astype(np.int64) << 42) | (ijk[:, 1].astype(np.int64) << 21) | ijk[:, 2].astype(np.int64)
    uniq, inv = np.unique(lin, return_inverse=True)
    n = len(uniq)
    cnt = np.bincount(inv, minlength=n).astype(np.float64)
    pts = np.stack([np.bincount(inv, weights=points[:, a], minlength=n) for a in range(3)], 1) / cnt[:, None]
    cols = None
    if colors is not None:
        cols = np.stack([np.bincount(inv, weights=colors[:, a], minlength=n) for a in range(3)], 1) / cnt[:, None]
    out_ijk = np.stack([(uniq >> 42), (uniq >> 21) & 0x1FFFFF, uniq & 0x1FFFFF], 1).astype(np.int32)
    return pts, cols, out_ijk, inv


# ----------------------------------------------------------------------------
# A3  outlier filters (graph.py:352-358)
# ----------------------------------------------------------------------------

def pcd_denoise_dbscan_identity(points, eps=0.01, min_points=100):
    """graph.py:352-353 -> utils/graph_utils.py:827-880.  With voxel_size >= 0.02 no
    point has ``min_points`` neighbours within ``eps`` (at most 8 voxel centroids fit in
    a 1 cm ball) so every DBSCAN label is -1, the cluster counter is empty and the
    cloud is returned unchanged (H6).  Asserted, then identity."""
    if len(points) > min_points:
        t = cKDTree(points)
        c = t.query_ball_point(points[: min(len(points), 2000)], eps, return_length=True)
        assert c.max() < min_points, "DBSCAN step is not a no-op for this cloud"
    return points


def radius_outlier_keep(points, nb_points=1000, radius=1.0):
    """Open3D RemoveRadiusOutliers (graph.py:355-358): keep i iff the number of points j
    (self included) with squared distance d2 < radius^2 is > nb_points (H5).  d2 is
    ((dx*dx + dy*dy) + dz*dz) in float64 (nanoflann L2_Simple accumulation order)."""
    n = len(points)
    tree = cKDTree(points)
    keep = np.zeros(n, dtype=bool)
    r2 = radius * radius
    for s in range(0, n, 4096):
        cand = tree.query_ball_point(points[s:s + 4096], radius * (1 + 1e-9) + 1e-12)
        for o, js in enumerate(cand):
            js = np.asarray(js, dtype=np.int64)
            dlt = points[js] - points[s + o]
            d2 = (dlt[:, 0] * dlt[:, 0] + dlt[:, 1] * dlt[:, 1]) + dlt[:, 2] * dlt[:, 2]
            keep[s + o] = np.count_nonzero(d2 < r2) > nb_points
    return np.nonzero(keep)[0]


def radius_counts(points, radius=1.0):
    """Neighbour counts (self included, d2 < r^2) for every point - test helper."""
    n = len(points)
    tree = cKDTree(points)
    out = np.zeros(n, dtype=np.int64)
    r2 = radius * radius
    for s in range(0, n, 4096):
        cand = tree.query_ball_point(points[s:s + 4096], radius * (1 + 1e-9) + 1e-12)
        for o, js in enumerate(cand):
            js = np.asarray(js, dtype=np.int64)
            dlt = points[js] - points[s + o]
            d2 = (dlt[:, 0] * dlt[:, 0] + dlt[:, 1] * dlt[:, 1]) + dlt[:, 2] * dlt[:, 2]
            out[s + o] = np.count_nonzero(d2 < r2)
    return out


# ----------------------------------------------------------------------------
# A4  pixel -> node nearest neighbour (graph.py:362-364, :409) - the reference's own call
# ----------------------------------------------------------------------------

def build_tree(node_xyz):
    return cKDTree(node_xyz)


def pixel_to_node(tree, pts):
    if len(pts) == 0:
        return np.zeros(0), np.zeros(0, np.int64)
    dis, idx = tree.query(pts, k=1, workers=-1)
    return dis, idx.astype(np.int64)


# ----------------------------------------------------------------------------
# A5  per-mask feature fusion (perception/models/sam_clip_feats_extractor.py:159-175)
# ----------------------------------------------------------------------------

def fuse_mask_feats(cropped_masked_feats, cropped_feats, F_g, maskedd_weight):
    """-> F_p float32 [M,d] (torch CPU ops exactly as the reference issues them)."""
    fused = torch.from_numpy(maskedd_weight * cropped_masked_feats + (1 - maskedd_weight) * cropped_feats)  # :159-160
    F_l = torch.nn.functional.normalize(fused, p=2, dim=-1).cpu().numpy()                                # :161-162
    cos = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)
    phi = cos(torch.from_numpy(F_l), torch.from_numpy(F_g))                                              # :167-168
    w_i = torch.nn.functional.softmax(phi, dim=0).reshape(-1, 1)                                         # :169
    F_p = w_i * F_g + (1 - w_i) * F_l.reshape(F_l.shape[0], F_l.shape[1])                                # :172
    F_p = torch.nn.functional.normalize(F_p, p=2, dim=-1)                                                # :175
    return F_p.float().numpy()


def pixel_feature_map(F_p, segs, H, W):
    """Dense per-pixel map of extractor.py:177-190 (zeros, += per mask, normalise,
    .half()), on CPU.  segs: bool [M,H,W].  Returns float16 [H*W,d] (H7)."""
    d = F_p.shape[1]
    out = torch.zeros(H * W, d)
    Fp = torch.from_numpy(F_p)
    flat = torch.from_numpy(np.asarray(segs).reshape(len(segs), -1))
    for i in range(len(segs)):
        ids = torch.argwhere(flat[i] == 1)
        out[ids, :] += Fp[i, :]
    out = torch.nn.functional.normalize(out, p=2, dim=-1)
    return out.half()


def pixel_features_at(F_p, segs_flat_at, ):
    """Sparse form of pixel_feature_map: rows only for the given pixels.
    segs_flat_at: bool [M,P] membership of each requested pixel in each mask.
    Accumulates in mask order in float32 exactly like the dense loop does."""
    M, P = segs_flat_at.shape
    out = torch.zeros(P, F_p.shape[1])
    Fp = torch.from_numpy(F_p)
    m = torch.from_numpy(np.asarray(segs_flat_at))
    for i in range(M):
        ids = torch.argwhere(m[i] == 1)
        out[ids, :] += Fp[i, :]
    out = torch.nn.functional.normalize(out, p=2, dim=-1)
    return out.half()


# ----------------------------------------------------------------------------
# A6  feature -> node aggregation (graph.py:404-415), hazard H1
# ----------------------------------------------------------------------------

def scatter_node_feats(sum_features, counter, idx, F_2D_valid):
    """``sum_features[idx] += F_2D ; counter[idx] += 1`` (graph.py:410-411) with the
    single-thread CPU semantics of torch's non-accumulating index_put_: for duplicate
    indices the LAST row in row-major pixel order wins and counter grows by one per
    frame (H1).  Uses torch itself with one thread."""
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        idx_t = torch.from_numpy(np.asarray(idx, dtype=np.int64))
        sum_features[idx_t] += F_2D_valid
        counter[idx_t] += 1
    finally:
        torch.set_num_threads(nt)


def winners(idx, n_nodes):
    """H1 as an explicit rule: for each node hit this frame, the position (in the
    frame's valid-pixel order) of the last pixel mapping to it.  Returns (nodes, pos)."""
    last = np.full(n_nodes, -1, dtype=np.int64)
    last[idx] = np.arange(len(idx))          # numpy fancy assignment: last write wins
    nodes = np.nonzero(last >= 0)[0]
    return nodes, last[nodes]


def finalize_node_feats(sum_features, counter):
    counter = counter.clone()
    counter[counter == 0] = 1e-5              # graph.py:413
    return (sum_features / counter).cpu().numpy()   # :414-415


# ----------------------------------------------------------------------------
# A7  per-mask 3-D node sets (generic.py:140-190)
# ----------------------------------------------------------------------------

def create_3d_masks(segs, depth, K, scale, camera_pose, node_xyz, node_rgb, tree, down_size):
    """Returns list of (pts [k,3], cols [k,3], ijk int32 [k,3]) per mask, canonical
    ascending-key order (H2); re-voxelised relative to the mask's own min bound with
    pixel multiplicity as weight (H3)."""
    out = []
    for seg in segs:
        pts, _, _ = create_pcd(np.asarray(seg), depth, K, scale, camera_pose, mask_img=True)   # :172-178
        if len(pts) == 0:
            out.append((np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3), np.int32)))
            continue
        _, indices = pixel_to_node(tree, pts)                                                # :181
        p, c, ijk, _ = voxel_down_sample(node_xyz[indices], node_rgb[indices], down_size)    # :182-188
        out.append((p, c, ijk))
    return out


# ----------------------------------------------------------------------------
# A8  crops + open_clip preprocess (utils/sam_utils.py:58-81,119-181; clip_utils.py:72-73,88-89)
# ----------------------------------------------------------------------------
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def increase_bbox_by_margin(bbox, margin):
    x, y, w, h = bbox
    x -= margin; y -= margin; w += margin * 2; h += margin * 2      # sam_utils.py:68-71
    if x < 0:
        w += x; x = 0                                                # :73-75
    if y < 0:
        h += y; y = 0                                                # :78-80
    return (x, y, w, h)


def crop_all_bounding_boxs(image, masks, block_background=False, bbox_margin=0):
    import cv2
    images = []
    for mask in masks:
        if block_background:                                         # sam_utils.py:140-141 -> crop_image :149-162
            x, y, w, h = mask["bbox"]
            masked = image * np.expand_dims(mask["segmentation"], -1)
            x, y, w, h = int(x), int(y), int(w), int(h)
            crop = masked[y:y + h, x:x + w, :]
        else:                                                        # :142-143 -> crop_bbox :165-181
            x, y, w, h = increase_bbox_by_margin(tuple(mask["bbox"]), bbox_margin)
            x, y, w, h = int(x), int(y), int(w), int(h)
            crop = image[y:y + h, x:x + w]
        crop = cv2.resize(crop, (512, 512))                          # :144  (INTER_LINEAR)
        images.append(crop)
    return images


def clip_preprocess(img_u8):
    """open_clip image_transform (eval): Resize(224, bicubic, antialias) on the shorter
    side, CenterCrop(224), RGB, ToTensor, Normalize(CLIP mean/std).  Un-vendored
    (open-clip-torch, environment.yaml:22); restated with PIL + torch."""
    from PIL import Image
    return clip_preprocess_pil(Image.fromarray(np.uint8(img_u8)))


def clip_preprocess_pil(img):
    """The `preprocess` callable the reference passes around (clip_utils.py:72-73,88-89), PIL image in."""
    from PIL import Image
    img = img.convert("RGB")
    w, h = img.size
    short, long = (w, h) if w <= h else (h, w)
    if short != 224:
        new_short, new_long = 224, int(224 * long / short)
        nw, nh = (new_short, new_long) if w <= h else (new_long, new_short)
        img = img.resize((nw, nh), Image.BICUBIC)
    w, h = img.size
    left = int(round((w - 224) / 2.0)); top = int(round((h - 224) / 2.0))
    img = img.crop((left, top, left + 224, top + 224))
    a = np.asarray(img, dtype=np.uint8)
    t = torch.from_numpy(a.copy()).permute(2, 0, 1).float().div(255.0)
    mean = torch.tensor(CLIP_MEAN).view(3, 1, 1); std = torch.tensor(CLIP_STD).view(3, 1, 1)
    return (t - mean) / std


# ----------------------------------------------------------------------------
# A9  ViT-B/32 visual tower (open_clip VisionTransformer, un-vendored) + clip_utils.py:63-94
# ----------------------------------------------------------------------------

@torch.no_grad()
def vit_forward(sd, x, heads=12, quick_gelu=False, return_tokens=False):
    """fp32 forward with the given (fp16-rounded) weights.  x float32 [B,3,224,224].
    Returns un-normalised [B,out_dim]."""
    F = torch.nn.functional
    w = sd["conv1.weight"]
    width = w.shape[0]
    x = F.conv2d(x, w, stride=w.shape[-1])                                  # [B,width,7,7]
    x = x.reshape(x.shape[0], width, -1).permute(0, 2, 1)                   # [B,49,width]
    cls = sd["class_embedding"].view(1, 1, width).expand(x.shape[0], 1, width)
    x = torch.cat([cls, x], 1) + sd["positional_embedding"]
    x = F.layer_norm(x, (width,), sd["ln_pre.weight"], sd["ln_pre.bias"], 1e-5)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    B, T, _ = x.shape
    hd = width // heads
    for i in range(n_layers):
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (width,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = h @ sd[p + "attn.in_proj_weight"].T + sd[p + "attn.in_proj_bias"]
        q, k, v = qkv.split(width, dim=-1)
        q = q.view(B, T, heads, hd).transpose(1, 2)
        k = k.view(B, T, heads, hd).transpose(1, 2)
        v = v.view(B, T, heads, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, T, width)
        x = x + o @ sd[p + "attn.out_proj.weight"].T + sd[p + "attn.out_proj.bias"]
        h = F.layer_norm(x, (width,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = h @ sd[p + "mlp.c_fc.weight"].T + sd[p + "mlp.c_fc.bias"]
        h = h * torch.sigmoid(1.702 * h) if quick_gelu else F.gelu(h)
        x = x + h @ sd[p + "mlp.c_proj.weight"].T + sd[p + "mlp.c_proj.bias"]
    if return_tokens:
        return x
    pooled = F.layer_norm(x[:, 0], (width,), sd["ln_post.weight"], sd["ln_post.bias"], 1e-5)
    return pooled @ sd["proj"]


@torch.no_grad()
def get_img_feats_batch_tensor(sd, x, **kw):
    """clip_utils.py:90-93: encode_image(...).float(); F.normalize(dim=-1); np.float32."""
    f = vit_forward(sd, x, **kw).float()
    f = torch.nn.functional.normalize(f, dim=-1)
    return np.float32(f.cpu())


# ----------------------------------------------------------------------------
# A10  query template mean (clip_utils.py:336-347)
# ----------------------------------------------------------------------------

def template_mean(text_feats, n_templates=2):
    tf = text_feats.reshape((-1, n_templates, text_feats.shape[-1]))
    return np.mean(tf, axis=1)


# ----------------------------------------------------------------------------
# A11  retrieval cores (memory/hmsg/graph/graph.py)
# ----------------------------------------------------------------------------

def _argsort_desc(scores):
    """np.argsort(s)[::-1] with ties pinned to the lower index first (H8)."""
    return np.lexsort((np.arange(len(scores)), -scores))


def query_topk(q, E, k):
    """graph.py:2196-2200 / :3127-3133: sim = np.dot(q, E.T); argsort desc; top-k.
    q [d] or [1,d]; returns (ids [k], scores [k])."""
    sim = np.dot(np.asarray(q).reshape(1, -1), E.T)[0]
    top = _argsort_desc(sim)[:k]
    return top, sim[top]


def query_object_core(query_feats, E, query_id, top_k, has_negative):
    """graph.py:3126-3151.  query_feats [Q',d] (row query_id is the target, the others are
    the negative prompts).  Returns (top_index, scores sim[query_id][top_index])."""
    sim = np.dot(query_feats, E.T)
    top = _argsort_desc(sim[query_id])[:top_k]                       # :3133
    if has_negative:
        cls_ids = np.argmax(sim, axis=0)                             # :3136
        max_scores = np.max(sim, axis=0)                             # :3139
        obj_ids = np.where(cls_ids == query_id)[0]                   # :3141
        if len(obj_ids) > 0:
            obj_scores = max_scores[obj_ids]
            resort = np.lexsort((np.arange(len(obj_ids)), -obj_scores))   # argsort(-scores), ties -> lower index
            top = obj_ids[resort][:top_k]                            # :3149-3150
    return top, sim[query_id][top]


def identify_object(object_feat, text_feats):
    """graph.py:1452-1454 -> argmax class index."""
    return int(np.argmax(np.dot(object_feat.reshape(1, -1), text_feats.T)))


def rooms_by_view_embedding(q, room_embs):
    """graph.py:3247-3257 / :3345-3350: per-room max of q . emb, rooms sorted by it."""
    scores = np.array([np.max(np.dot(q.reshape(1, -1), np.asarray(e).T)) for e in room_embs])
    return _argsort_desc(scores), scores


# ----------------------------------------------------------------------------
# End-to-end build (graph.py:339-415) on in-memory frames: BASELINE config 1
# ----------------------------------------------------------------------------

# ----------------------------------------------------------------------------
# N1 / N2 building blocks (utils/graph_utils.py:620-728, :827-880, :883-956, :1015-1038)
# ----------------------------------------------------------------------------

def cluster_dbscan(points, eps, min_points):
    """Open3D 0.18 PointCloud::ClusterDBSCAN (graph_utils.py:838-841): neighbours = points with
    d2 < eps^2 (self included, nanoflann radius set, same d2 accumulation as H5); core iff
    |nbs| >= min_points; clusters are numbered in order of their lowest-index core point and
    fully expanded before the next one starts, so a border point takes the LOWEST-numbered
    cluster among its core neighbours; everything else is -1.  (The result does not depend on
    Open3D's unordered_set pop order.)"""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    n = len(pts)
    labels = np.full(n, -1, dtype=np.int64)
    if n == 0:
        return labels
    tree = cKDTree(pts)
    cand = tree.query_ball_point(pts, eps * (1 + 1e-9) + 1e-12)
    e2 = eps * eps
    nbs = []
    for i, js in enumerate(cand):
        js = np.asarray(js, dtype=np.int64)
        dlt = pts[js] - pts[i]
        d2 = (dlt[:, 0] * dlt[:, 0] + dlt[:, 1] * dlt[:, 1]) + dlt[:, 2] * dlt[:, 2]
        nbs.append(js[d2 < e2])
    core = np.array([len(v) >= min_points for v in nbs])
    comp = np.full(n, -1, dtype=np.int64)
    c = 0
    for i in range(n):                       # components of the core-core adjacency, by lowest core index
        if not core[i] or comp[i] >= 0:
            continue
        stack = [i]; comp[i] = c
        while stack:
            u = stack.pop()
            for v in nbs[u]:
                if core[v] and comp[v] < 0:
                    comp[v] = c; stack.append(v)
        c += 1
    labels[core] = comp[core]
    for i in range(n):
        if not core[i]:
            cn = [comp[v] for v in nbs[i] if core[v]]
            if cn:
                labels[i] = min(cn)
    return labels


def largest_label(labels):
    """Counter(labels) minus -1, most_common(1): largest count, ties -> first label in array order
    (graph_utils.py:848-856 / :699-709).  Returns None when every label is -1."""
    from collections import Counter
    cnt = Counter(int(v) for v in labels)
    cnt.pop(-1, None)
    if not cnt:
        return None
    return cnt.most_common(1)[0][0]


def pcd_denoise_dbscan(points, colors, eps=0.02, min_points=10):
    """graph_utils.py:827-880: keep the largest DBSCAN cluster unless it has < 5 points."""
    lab = cluster_dbscan(points, eps, min_points)
    best = largest_label(lab)
    if best is None:
        return points, colors
    m = lab == best
    if m.sum() < 5:
        return points, colors
    return points[m], (None if colors is None else colors[m])


def flat_l2_nn(q, x):
    """faiss IndexFlatL2.search(k=1) restated (faiss-gpu 1.7.x, fsr_vln/environment.yaml): float32
    squared L2, ((dx*dx + dy*dy) + dz*dz) per pair, lowest index on ties.  (faiss switches to a
    |x|^2+|y|^2-2xy BLAS form for >= 20 queries, whose float32 rounding differs in the last ulp;
    this restatement uses the exact form throughout - thresholded counts agree except at ulp
    distance from the threshold.)"""
    q = np.ascontiguousarray(q, np.float32); x = np.ascontiguousarray(x, np.float32)
    D = np.empty(len(q), np.float32); I = np.empty(len(q), np.int64)
    for s in range(0, len(q), 1024):
        d = q[s:s + 1024, None, :] - x[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        I[s:s + 1024] = d2.argmin(1)
        D[s:s + 1024] = d2[np.arange(d2.shape[0]), I[s:s + 1024]]
    return D, I


def bbox_iou_3d(lo1, hi1, lo2, hi2):
    """graph_utils.py:883-916 compute_3d_bbox_iou (padding 0)."""
    size = np.maximum(np.minimum(hi1, hi2) - np.maximum(lo1, lo2), 0.0)
    ov = np.prod(size); v1 = np.prod(hi1 - lo1); v2 = np.prod(hi2 - lo2)
    return ov / (v1 + v2 - ov)


def overlapping_ratio(p1, p2, radius):
    """graph_utils.py:620-662 find_overlapping_ratio_faiss."""
    if len(p1) == 0 or len(p2) == 0:
        return 0
    D1, _ = flat_l2_nn(p1, p2); D2, _ = flat_l2_nn(p2, p1)
    return np.max([np.sum(D1 < radius ** 2) / len(p1), np.sum(D2 < radius ** 2) / len(p2)])


def merge_3d_masks(masks, overlap_threshold, radius, iou_thresh):
    """graph_utils.py:919-956.  masks: list of (pts, cols).  Returns merged list (component order =
    scipy connected_components labels, members concatenated in index order, then
    pcd_denoise_dbscan(eps=0.1, min_points=10))."""
    from scipy.sparse.csgraph import connected_components
    n = len(masks)
    if n == 0:
        return masks
    lo = [m[0].min(0) if len(m[0]) else np.zeros(3) for m in masks]
    hi = [m[0].max(0) if len(m[0]) else np.zeros(3) for m in masks]
    ov = np.zeros((n, n))
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n):
            for j in range(i + 1, n):
                if bbox_iou_3d(lo[i], hi[i], lo[j], hi[j]) > iou_thresh:
                    ov[i, j] = overlapping_ratio(masks[i][0], masks[j][0], 1.5 * radius)
    ncomp, lab = connected_components(ov > overlap_threshold)
    out = []
    for c in range(ncomp):
        idx = np.where(lab == c)[0]
        pts = np.concatenate([masks[i][0] for i in idx], 0)
        cols = np.concatenate([masks[i][1] for i in idx], 0)
        out.append(pcd_denoise_dbscan(pts, cols, eps=0.1, min_points=10))
    return out


def seq_merge(frames_masks, th, down_size, proxy_th):
    """graph_utils.py:1015-1038."""
    g = list(frames_masks[0])
    for i in range(1, len(frames_masks)):
        g = merge_3d_masks(g + list(frames_masks[i]), th, down_size, proxy_th)
    return merge_3d_masks(g, th, down_size, proxy_th)


def merge_adjacent_frames(frames_masks, th, down_size, proxy_th):
    """graph_utils.py:958-985: pairs (0,1), (2,3), ...; an odd last frame is carried over unmerged."""
    out = []
    for i in range(0, len(frames_masks), 2):
        if i == len(frames_masks) - 1:
            out.append(list(frames_masks[i]))
            break
        out.append(merge_3d_masks(list(frames_masks[i]) + list(frames_masks[i + 1]), th, down_size, proxy_th))
    return out


def hierarchical_merge(frames_masks, th, th_factor, down_size, proxy_th):
    """graph_utils.py:989-1012: merge adjacent frame lists level by level, the threshold drops by
    th_factor * (n - 2) / max(1, n - 1) after every level that leaves n > 1 lists; one more merge at 0.75."""
    frames_masks = [list(f) for f in frames_masks]
    while len(frames_masks) > 1:
        frames_masks = merge_adjacent_frames(frames_masks, th, down_size, proxy_th)
        if len(frames_masks) > 1:
            th -= th_factor * (len(frames_masks) - 2) / max(1, len(frames_masks) - 1)
    return merge_3d_masks(frames_masks[0], 0.75, down_size, proxy_th)


def feats_denoise_dbscan(feats, eps=0.01, min_points=100):
    """graph_utils.py:682-728 - sklearn DBSCAN(metric="cosine") is the reference's own call."""
    from sklearn.cluster import DBSCAN
    feats = np.array(feats)
    labels = DBSCAN(eps=eps, min_samples=min_points, metric="cosine").fit(feats).labels_
    best = largest_label(labels)
    if best is None:
        return np.mean(feats, axis=0)
    sel = feats[labels == best]
    return np.mean(sel, axis=0) if len(sel) > 1 else sel


def object_feats(mask_pts, node_xyz, tree, full_feats, voxel_size, dim):
    """graph.py:451-488: per merged mask -> voxel_down_sample -> NN into the node table ->
    drop matches farther than 0.8 -> nan_to_num -> feats_denoise_dbscan(0.01, 100)."""
    out = []
    for pts in mask_pts:
        p, _, _, _ = voxel_down_sample(pts, None, voxel_size)
        dist, idx = tree.query(p, k=1)
        feats = np.nan_to_num(full_feats[idx[dist <= 0.8]])
        out.append(np.zeros((1, dim), full_feats.dtype) if feats.shape[0] == 0 else feats_denoise_dbscan(feats, 0.01, 100))
    return out


def build_geometry(depths, rgbs, poses, K, scale, voxel_size, nb_points=1000, radius=1.0):
    """graph.py:339-364.  Returns dict with the voxel table (pre-filter), kept indices and nodes."""
    P, C = [], []
    for f in range(len(depths)):
        p, c, _ = create_pcd(rgbs[f], depths[f], K, scale, poses[f])
        P.append(p); C.append(c)
    P = np.concatenate(P); C = np.concatenate(C)
    vx, vc, ijk, _ = voxel_down_sample(P, C, voxel_size)
    pcd_denoise_dbscan_identity(vx)
    keep = radius_outlier_keep(vx, nb_points, radius)
    return {"voxel_xyz": vx, "voxel_rgb": vc, "voxel_ijk": ijk, "keep": keep,
            "node_xyz": vx[keep], "node_rgb": vc[keep], "node_ijk": ijk[keep],
            "min_bound": P.min(axis=0), "n_points": len(P)}


def ingest_frame(sum_features, counter, tree, n_nodes, depth, rgb, pose, K, scale, F_p, segs, idx=None):
    """graph.py:390, :404-411 for one frame given its mask embeddings F_p and masks.
    Uses the sparse-at-winners form of the dense feature map (equal by construction,
    verified against the dense form in tests).  ``idx`` overrides the KD-tree result
    (stage-wise parity: exact-distance ties are implementation-defined in cKDTree)."""
    pts, _, valid = create_pcd(rgb, depth, K, scale, pose)
    if idx is None:
        _, idx = pixel_to_node(tree, pts)
    nodes, pos = winners(idx, n_nodes)
    vpix = np.nonzero(valid.reshape(-1))[0]
    wpix = vpix[pos]
    member = np.asarray(segs).reshape(len(segs), -1)[:, wpix]
    feats = pixel_features_at(F_p, member).float()
    sum_features[torch.from_numpy(nodes)] += feats
    counter[torch.from_numpy(nodes)] += 1
    return idx, nodes, wpix
