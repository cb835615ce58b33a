"""Oracle vs fixtures produced by the reference's OWN code (tests/golden/make_reference_golden.py ran
Graph.create_feature_map and Graph.query_hmsg_object from /root/reference/fsr_vln here and committed
what they returned).  CPU-only; the CUDA path is checked against the same fixtures in
test_gpu_reference_golden.py."""
import json
import os

import numpy as np
import pytest
import torch

from holoagent_b200 import synth
from oracle import hmsg_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_build():
    z = np.load(os.path.join(GOLD, "ref_build.npz"))
    H, W = int(z["H"]), int(z["W"])
    depth, rgb, T, K = synth.make_frames_np(z["frame_ids"], H, W)
    segs = np.unpackbits(z["segs"], axis=-1)[..., :W].astype(bool)
    vit = [int(v) for v in z["vit"]]
    sh = synth.VitB32Shape(image=vit[0], patch=vit[1], width=vit[2], layers=vit[3], heads=vit[4], mlp=vit[5], out_dim=vit[6])
    return z, depth, rgb, T, K, segs, sh


def masks_of(segs_f, bboxes_f):
    return [{"segmentation": segs_f[m], "bbox": [int(v) for v in bboxes_f[m]]} for m in range(len(segs_f))]


@pytest.fixture(scope="module")
def built():
    """The oracle's build of the fixture scene, stage by stage."""
    z, depth, rgb, T, K, segs, sh = load_build()
    sd = synth.make_vit_weights(sh, seed=int(z["vit_seed"]))
    g = O.build_geometry(depth, rgb, T, K, 1000.0, float(z["voxel_size"]))
    tree = O.build_tree(g["node_xyz"])
    n = len(g["node_xyz"])
    d = sh.out_dim
    sum_f = torch.zeros(n, d); cnt = torch.zeros(n, 1)
    Fps, Fgs, samples, frames_masks = [], [], [], []
    step = int(z["pix_step"])
    for f in range(len(depth)):
        masks = masks_of(segs[f], z["bboxes"][f])
        plain = O.crop_all_bounding_boxs(rgb[f], masks, block_background=False, bbox_margin=int(z["bbox_margin"]))
        masked = O.crop_all_bounding_boxs(rgb[f], masks, block_background=True, bbox_margin=int(z["bbox_margin"]))
        enc = lambda imgs: O.get_img_feats_batch_tensor(sd, torch.stack([O.clip_preprocess(i) for i in imgs]), heads=sh.heads)
        F_g = enc([rgb[f]])
        Fp = O.fuse_mask_feats(enc(masked), enc(plain), F_g, float(z["maskedd_weight"]))
        Fps.append(Fp); Fgs.append(F_g.reshape(-1))
        samples.append(O.pixel_feature_map(Fp, segs[f], depth.shape[1], depth.shape[2]).numpy()[::step])
        O.ingest_frame(sum_f, cnt, tree, n, depth[f], rgb[f], T[f], K, 1000.0, Fp, segs[f])
        frames_masks.append([(p, c) for p, c, _ in O.create_3d_masks(segs[f], depth[f], K, 1000.0, T[f], g["node_xyz"], g["node_rgb"], tree, float(z["voxel_size"]))])
    full = O.finalize_node_feats(sum_f, cnt)
    return dict(z=z, g=g, tree=tree, Fp=np.stack(Fps), Fg=np.stack(Fgs), samples=np.stack(samples), full=full, frames_masks=frames_masks, d=d)


def test_geometry_matches_reference_run(built):
    z, g = built["z"], built["g"]
    assert g["node_xyz"].shape == z["node_xyz"].shape
    assert np.array_equal(g["node_xyz"], z["node_xyz"])          # same canonical order, bit-identical doubles
    assert np.array_equal(g["node_rgb"], z["node_rgb"])


def test_mask_embeddings_match_reference_run(built):
    z = built["z"]
    assert np.abs(built["Fg"] - z["F_g"]).max() < 1e-6
    assert np.abs(built["Fp"] - z["F_p"]).max() < 1e-6
    a, b = built["samples"].astype(np.float32), z["F2D_sample"].astype(np.float32)
    assert a.shape == b.shape and np.abs(a - b).max() <= 2 ** -10       # fp16 maps: at most one fp16 ulp near 1.0


def test_node_features_match_reference_run(built):
    z, full = built["z"], built["full"]
    assert np.abs(full[::int(z["row_step"])] - z["full_feats_rows"]).max() < 1e-6
    assert np.abs(full.astype(np.float64).sum(1) - z["full_feats_rowsum"]).max() < 1e-4


def test_object_layer_matches_reference_run(built):
    """N1 + N2: seq_merge -> small-mask removal -> per-object DBSCAN feature (graph.py:424-488)."""
    z = built["z"]
    objs = O.seq_merge(built["frames_masks"], 0.75, float(z["voxel_size"]), 0.05)
    objs = [o for o in objs if len(o[0]) >= 10]
    off = z["obj_off"]
    assert len(objs) == len(off) - 1
    for i, (p, _) in enumerate(objs):
        assert np.array_equal(p, z["obj_pts"][off[i]:off[i + 1]]), i
    feats = O.object_feats([o[0] for o in objs], built["g"]["node_xyz"], built["tree"], built["full"], float(z["voxel_size"]), built["d"])
    got = np.stack([np.asarray(f).reshape(-1) for f in feats])
    assert np.abs(got - z["mask_feats"]).max() < 1e-6


def test_query_object_matches_reference_run():
    z = np.load(os.path.join(GOLD, "ref_query.npz"))
    cases = json.loads(str(z["q_cases"])); words = json.loads(str(z["q_words"]))
    tf, emb, room = z["q_text_feats"], z["q_obj_emb"], z["q_obj_room"]
    for ci, (q, rooms, k, neg) in enumerate(cases):
        sel = np.concatenate([np.nonzero(room == r)[0] for r in rooms])          # graph.py:3112-3123 candidate order
        if q in neg:
            qid, names = neg.index(q), neg
        else:
            qid, names = 0, [q] + neg
        qf = np.stack([tf[words.index(w)] for w in names])
        ids, scores = O.query_object_core(qf, emb[sel], qid, k, len(neg) > 0)
        assert np.array_equal(sel[ids], z["q%d_ids" % ci]), ci
        assert np.array_equal(room[sel[ids]], z["q%d_rooms" % ci]), ci
        assert np.allclose(scores, z["q%d_scores" % ci], rtol=0, atol=1e-6), ci


def test_cluster_dbscan_restatement():
    """Open3D ClusterDBSCAN restatement vs a literal transcription of its sequential loop, and vs
    sklearn's euclidean DBSCAN on the core points."""
    from sklearn.cluster import DBSCAN
    rs = np.random.RandomState(5)
    pts = np.concatenate([rs.randn(150, 3) * 0.1, rs.randn(150, 3) * 0.1 + [0.45, 0, 0], rs.rand(60, 3) * 2 - 1])
    eps, mp = 0.12, 8
    lab = O.cluster_dbscan(pts, eps, mp)
    d2 = ((pts[:, None] - pts[None]) ** 2).sum(-1)
    nbs = [np.nonzero(d2[i] < eps * eps)[0] for i in range(len(pts))]
    ref = np.full(len(pts), -2); c = 0
    for i in range(len(pts)):
        if ref[i] != -2:
            continue
        if len(nbs[i]) < mp:
            ref[i] = -1; continue
        nxt, seen = set(nbs[i].tolist()), {i}
        ref[i] = c
        while nxt:
            nb = nxt.pop(); seen.add(nb)
            if ref[nb] == -1:
                ref[nb] = c
            if ref[nb] != -2:
                continue
            ref[nb] = c
            if len(nbs[nb]) >= mp:
                nxt.update(int(q) for q in nbs[nb] if q not in seen)
        c += 1
    assert np.array_equal(lab, ref)
    sk = DBSCAN(eps=eps, min_samples=mp).fit(pts)
    core = np.zeros(len(pts), bool); core[sk.core_sample_indices_] = True
    pairs = {(a, b) for a, b in zip(lab[core], sk.labels_[core])}
    assert len(pairs) == len({a for a, _ in pairs}) == len({b for _, b in pairs})    # same partition of the core points
