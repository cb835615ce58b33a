"""Generates the committed known-answer fixtures under tests/golden/ from the CPU oracle
(which itself calls the real scipy cKDTree / torch / numpy routines the reference calls).
The reference has no tests or golden vectors for this path and cannot be
imported offline, so these fixtures pin the ORACLE; run `python tests/golden/make_golden.py`
to regenerate (deterministic)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from holoagent_b200 import synth  # noqa: E402
from oracle import hmsg_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tiny_scene():
    H, W, F, M, d = 30, 40, 4, 6, 128
    vs, nb, radius, w = 0.25, 3, 0.6, 0.4418
    depth, rgb, T, K = synth.make_frames_np(np.arange(F) * 7, H, W)
    g = O.build_geometry(depth, rgb, T, K, 1000.0, vs, nb_points=nb, radius=radius)
    tree = O.build_tree(g["node_xyz"])
    n = len(g["node_xyz"])
    rs = np.random.RandomState(42)
    feats = rs.randn(F, 2 * M + 1, d).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=-1, keepdims=True)
    boxes = np.stack([synth.make_mask_boxes(100 + f, H, W, M) for f in range(F)])
    boxes[:, :, 2] = np.minimum(boxes[:, :, 2], 20); boxes[:, :, 3] = np.minimum(boxes[:, :, 3], 16)
    boxes[:, :, 0] = np.minimum(boxes[:, :, 0], W - boxes[:, :, 2]); boxes[:, :, 1] = np.minimum(boxes[:, :, 1], H - boxes[:, :, 3])
    sum_f = torch.zeros(n, d); cnt = torch.zeros(n, 1)
    idxs, Fps = [], []
    for f in range(F):
        valid = depth[f] > 0
        segs = np.zeros((M, H, W), bool)
        for m, (x, y, ww, hh) in enumerate(boxes[f]):
            segs[m, y:y + hh, x:x + ww] = valid[y:y + hh, x:x + ww]
        Fp = O.fuse_mask_feats(feats[f, :M], feats[f, M:2 * M], feats[f, 2 * M:], w)
        idx, _, _ = O.ingest_frame(sum_f, cnt, tree, n, depth[f], rgb[f], T[f], K, 1000.0, Fp, segs)
        full_idx = np.full(H * W, -1, np.int64); full_idx[valid.reshape(-1)] = idx
        idxs.append(full_idx); Fps.append(Fp)
    full = O.finalize_node_feats(sum_f, cnt)
    E = full.astype(np.float32)
    q = feats[:3, 0] * 0.8
    top = np.stack([O.query_topk(q[i], E, 4)[0] for i in range(3)])
    tops = np.stack([O.query_topk(q[i], E, 4)[1] for i in range(3)])
    np.savez_compressed(os.path.join(HERE, "tiny_scene.npz"), depth=depth, rgb=rgb, poses=T, K=K, voxel_size=vs, nb_points=nb, radius=radius,
                        maskedd_weight=w, feats=feats, boxes=boxes, min_bound=g["min_bound"], voxel_xyz=g["voxel_xyz"], voxel_rgb=g["voxel_rgb"],
                        voxel_ijk=g["voxel_ijk"], keep=g["keep"], node_xyz=g["node_xyz"], pix_idx=np.stack(idxs), F_p=np.stack(Fps),
                        sum_features=sum_f.numpy(), counter=cnt.numpy().reshape(-1), full_feats=full, query=q, top_ids=top, top_scores=tops)
    print("tiny_scene: voxels", len(g["voxel_xyz"]), "nodes", n)


def small_vit():
    sh = synth.VitB32Shape(image=64, patch=32, width=256, layers=2, heads=4, mlp=512, out_dim=256)
    sd = synth.make_vit_weights(sh, seed=3)
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(9))
    out = O.get_img_feats_batch_tensor(sd, x, heads=4)
    np.savez_compressed(os.path.join(HERE, "small_vit.npz"), x=x.numpy(), out=out)
    print("small_vit:", out.shape)


if __name__ == "__main__":
    tiny_scene()
    small_vit()
