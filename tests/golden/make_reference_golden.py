"""Runs the UNMODIFIED reference code (fsr_vln, /root/reference) on a small synthetic scene and commits
what it produced as fixtures (tests/golden/ref_*.npz).  These pin the oracle - and through it the
CUDA path - against the reference itself rather than against a restatement.

What executes here is the reference's own source, imported from /root/reference/fsr_vln:
  * Graph.create_feature_map                      memory/hmsg/graph/graph.py:262-491  (whole method)
  * RGBDDataset.create_pcd / create_3d_masks      memory/hmsg/dataloader/generic.py:74-190
  * extract_feats_per_pixel                       perception/models/sam_clip_feats_extractor.py:93-190
  * get_img_feats / get_img_feats_batch           memory/hmsg/utils/clip_utils.py:63-92
  * crop_all_bounding_boxs & friends              memory/hmsg/utils/sam_utils.py:58-181
  * seq_merge / merge_3d_masks / find_overlapping_ratio_faiss / compute_3d_bbox_iou /
    merge_point_clouds_list / pcd_denoise_dbscan / feats_denoise_dbscan   memory/hmsg/utils/graph_utils.py
  * Graph.query_hmsg_object                       memory/hmsg/graph/graph.py (retrieval half)
together with the real scipy cKDTree / connected_components, sklearn DBSCAN, torch, cv2 and PIL.
Stand-ins (tests/golden/ref_shims.py): open3d.PointCloud and faiss.IndexFlatL2 (absent here) are
functional shims over the oracle's restatement of those libraries; SAM (`mask_generator`), the CLIP
towers (`clip_model`, a small seeded ViT evaluated in fp32 torch) and `preprocess` are constructor
arguments of the reference and are passed in.

    python tests/golden/make_reference_golden.py        (container only; deterministic)
"""
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
from holoagent_b200 import synth  # noqa: E402
from oracle import hmsg_oracle as O  # noqa: E402

import memory.hmsg.graph.graph as ref_graph  # noqa: E402   (reference)
import memory.hmsg.dataloader.generic as ref_generic  # noqa: E402   (reference)

H, W, F, M = 96, 128, 6, 5
VS = 0.05
VIT = dict(image=224, patch=32, width=256, layers=2, heads=4, mlp=512, out_dim=256)      # smallest shape the CUDA encoder takes
ROW_STEP, PIX_STEP = 8, 97                                                                # fixture keeps every 8th node row / 97th pixel
FRAME_IDS = np.arange(F) * 3


class SynthDataset(ref_generic.RGBDDataset):
    """Concrete subclass of the reference's abstract RGBDDataset over synthetic frames: only the
    abstract loaders are supplied; create_pcd / create_3d_masks are the reference's own."""

    def __init__(self, depth, rgb, poses, K, scale):
        self._d, self._c, self._T = depth, rgb, poses
        super().__init__({"root_dir": "", "transforms": None, "depth_cut": 10.0})
        self.depth_intrinsics = K
        self.rgb_intrinsics = K
        self.scale = scale

    def _get_data_list(self):
        return list(range(len(self._d)))

    def __getitem__(self, i):
        return Image.fromarray(self._c[i]), Image.fromarray(self._d[i]), self._T[i], None, self.depth_intrinsics


class FixedMasks:
    """mask_generator stand-in: SAM's output format (sam_utils.py / extractor.py keys) for seeded masks."""

    def __init__(self, per_frame):
        self.per_frame, self.i, self.calls = per_frame, 0, []

    def generate(self, image):
        m = self.per_frame[self.i % len(self.per_frame)]
        self.i += 1
        return m


class SmallClip:
    def __init__(self, sd, heads):
        self.sd, self.heads = sd, heads

    def encode_image(self, x):
        return O.vit_forward(self.sd, x.float(), heads=self.heads)


def make_masks(fid, depth):
    """5 masks per frame: seeded rectangles AND depth>0, one of them L-shaped (bbox != filled rect)."""
    rs = np.random.RandomState(500 + int(fid))
    out = []
    for m in range(M):
        w, h = rs.randint(20, 70), rs.randint(20, 60)
        x, y = rs.randint(0, W - w + 1), rs.randint(0, H - h + 1)
        seg = np.zeros((H, W), bool)
        seg[y:y + h, x:x + w] = True
        if m == 1:
            seg[y:y + h // 2, x:x + w // 2] = False
        seg &= depth > 0
        ys, xs = np.nonzero(seg)
        bbox = [int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)]
        out.append({"segmentation": seg, "bbox": bbox, "area": int(seg.sum()), "predicted_iou": 0.9, "stability_score": 0.95})
    return out


def main():
    # `sum_features[idx] += F_2D` (graph.py:410) has duplicate indices; torch's CPU index_put_ splits the
    # index list across intra-op threads, so WHICH pixel's row survives depends on the thread count and,
    # from ~64 threads, on timing (measured here: 13 / 17522 node rows differ between 1 and 8 threads, 87
    # rows differ between two runs at 64 threads).  One thread gives the sequential meaning (last pixel
    # in row-major order wins) - that is what the fixture records.
    torch.set_num_threads(1)
    depth, rgb, T, K = synth.make_frames_np(FRAME_IDS, H, W)
    sh = synth.VitB32Shape(**VIT)
    sd = synth.make_vit_weights(sh, seed=11)
    masks = [make_masks(fid, depth[i]) for i, fid in enumerate(FRAME_IDS)]

    g = ref_graph.Graph.__new__(ref_graph.Graph)
    NS = types.SimpleNamespace
    g.cfg = NS(pipeline=NS(skip_frames=1, voxel_size=VS, clip_bbox_margin=10, clip_masked_weight=0.4418, max_mask_distance=6.0,
                           merge_type="sequential", init_overlap_thresh=0.75, overlap_thresh_factor=0.025, iou_thresh=0.05),
               main=NS(save_path="/tmp/ref_golden_out"))
    g.dataset = SynthDataset(depth, rgb, T, K, 1000.0)
    import open3d as o3d
    g.full_pcd = o3d.geometry.PointCloud()
    g.mask_generator = FixedMasks(masks)
    g.clip_model = SmallClip(sd, VIT["heads"])
    g.preprocess = O.clip_preprocess_pil          # open_clip image_transform(224, is_train=False) restated (absent here)
    g.clip_feat_dim = VIT["out_dim"]

    # record what the reference's extractor returns for every frame
    rec = []
    real_extract = ref_graph.extract_feats_per_pixel

    def spy(*a, **k):
        r = real_extract(*a, **k)
        rec.append((r[0].clone(), r[1].clone(), np.array(r[3])))
        return r
    ref_graph.extract_feats_per_pixel = spy
    # the reference hard-codes remove_radius_outlier(nb_points=1000, radius=1.0): keep it.
    g.create_feature_map()
    ref_graph.extract_feats_per_pixel = real_extract

    node_xyz = np.asarray(g.full_pcd.points)
    node_rgb = np.asarray(g.full_pcd.colors)
    print("nodes", node_xyz.shape, "objects", len(g.mask_pcds), [len(p.points) for p in g.mask_pcds])
    F2D = np.stack([r[0].numpy().reshape(H * W, -1) for r in rec])                # fp16 [F,H*W,d]
    Fp = np.stack([r[1].numpy() for r in rec])                                    # [F,M,d]
    Fg = np.stack([r[2].reshape(-1) for r in rec])
    obj_off = np.cumsum([0] + [len(p.points) for p in g.mask_pcds])
    obj_pts = np.concatenate([np.asarray(p.points) for p in g.mask_pcds], 0) if len(g.mask_pcds) else np.zeros((0, 3))
    mask_feats = np.stack([np.asarray(f).reshape(-1) for f in g.mask_feats]) if len(g.mask_feats) else np.zeros((0, VIT["out_dim"]))

    # retrieval half: the reference's Graph.query_hmsg_object (graph.py:3056-3161) over 300 seeded objects
    # in 3 rooms; the CLIP text tower is out of scope -> seeded unit "text features" keyed by string.
    rs = np.random.RandomState(77)
    d = VIT["out_dim"]
    words = ["chair", "table", "lamp", "sofa", "background", "wall"]
    tf = rs.randn(len(words), d).astype(np.float32); tf /= np.linalg.norm(tf, axis=1, keepdims=True)
    ref_graph.get_text_feats_multiple_templates = lambda q, m, dim: np.stack([tf[words.index(w)] for w in q])
    emb = rs.randn(300, d).astype(np.float32); emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    emb = (0.6 * emb + 0.4 * tf[rs.randint(0, len(words), 300)]).astype(np.float32)
    room_of = rs.randint(0, 3, 300)
    g.objects = [NS(embedding=emb[i], object_id="obj_%d" % i, room_id="room_%d" % room_of[i]) for i in range(300)]
    g.rooms = [NS(room_id="room_%d" % r, objects=[o for o in g.objects if o.room_id == "room_%d" % r]) for r in range(3)]
    g.floors = []
    q_cases = [("chair", [0, 1, 2], 5, []), ("table", [1], 3, []), ("lamp", [0, 2], 7, ["background", "wall"]),
               ("sofa", [0, 1, 2], 4, ["sofa", "chair", "background"]), ("wall", [2], 300, ["background"])]
    q_out = {}
    for ci, (q, rooms, k, neg) in enumerate(q_cases):
        ids, rids, sc = g.query_hmsg_object(q, room_ids=rooms, top_k=k, negative_prompt=list(neg))
        q_out["q%d_ids" % ci] = np.array(ids, np.int64); q_out["q%d_rooms" % ci] = np.array(rids, np.int64)
        q_out["q%d_scores" % ci] = np.array(sc, np.float32)
    import json
    q_out["q_cases"] = np.array(json.dumps(q_cases)); q_out["q_words"] = np.array(json.dumps(words))
    q_out["q_text_feats"] = tf; q_out["q_obj_emb"] = emb; q_out["q_obj_room"] = room_of
    np.savez_compressed(os.path.join(HERE, "ref_query.npz"), **q_out)
    print("saved ref_query.npz", os.path.getsize(os.path.join(HERE, "ref_query.npz")) // 1024, "KiB")

    segs = np.stack([np.stack([m["segmentation"] for m in fm]) for fm in masks])
    bboxes = np.array([[m["bbox"] for m in fm] for fm in masks], np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_build.npz"), frame_ids=FRAME_IDS, H=H, W=W, voxel_size=VS, vit=np.array(list(VIT.values())),
                        vit_seed=11, bbox_margin=10, maskedd_weight=0.4418, segs=np.packbits(segs, axis=-1), bboxes=bboxes,
                        node_xyz=node_xyz, node_rgb=node_rgb, F2D_sample=F2D[:, ::PIX_STEP], pix_step=PIX_STEP, row_step=ROW_STEP, F_p=Fp, F_g=Fg,
                        full_feats_rows=g.full_feats_array[::ROW_STEP], full_feats_rowsum=g.full_feats_array.astype(np.float64).sum(1), obj_off=obj_off, obj_pts=obj_pts, mask_feats=mask_feats)
    print("saved ref_build.npz", os.path.getsize(os.path.join(HERE, "ref_build.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
