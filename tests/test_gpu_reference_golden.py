"""GPU: the CUDA path (through the drop-in Graph and the C-ABI) against fixtures produced by the
reference's OWN code (tests/golden/make_reference_golden.py: Graph.create_feature_map and
Graph.query_hmsg_object from /root/reference/fsr_vln executed in the build container)."""
import json
import os
import types

import numpy as np
import pytest
import torch

from holoagent_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _DS:
    def __init__(self, z):
        self.H, self.W = int(z["H"]), int(z["W"])
        self.depth, self.rgb, self.T, self.K = synth.make_frames_np(z["frame_ids"], self.H, self.W)
        self.depth_intrinsics = self.K
        self.scale = 1000.0

    def __len__(self):
        return len(self.depth)

    def __getitem__(self, i):
        return self.rgb[i], self.depth[i], self.T[i], self.K, self.K


class _SAM:
    def __init__(self, z, W):
        segs = np.unpackbits(z["segs"], axis=-1)[..., :W].astype(bool)
        self.q = [[{"segmentation": segs[f][m], "bbox": [int(v) for v in z["bboxes"][f][m]], "predicted_iou": 0.9} for m in range(segs.shape[1])]
                  for f in range(segs.shape[0])]
        self.i = 0

    def generate(self, image):
        m = self.q[self.i % len(self.q)]
        self.i += 1
        return m


def test_build_matches_reference_run(engine):
    from holoagent_b200.memory.hmsg.graph.graph import Graph
    from holoagent_b200.memory.hmsg.utils.clip_utils import B200ClipModel
    z = np.load(os.path.join(GOLD, "ref_build.npz"))
    vit = [int(v) for v in z["vit"]]
    sh = synth.VitB32Shape(image=vit[0], patch=vit[1], width=vit[2], layers=vit[3], heads=vit[4], mlp=vit[5], out_dim=vit[6])
    sd = synth.make_vit_weights(sh, seed=int(z["vit_seed"]))
    ds = _DS(z)
    clip = B200ClipModel(engine, sd, image=sh.image, patch=sh.patch, width=sh.width, layers=sh.layers, heads=sh.heads, mlp=sh.mlp, out_dim=sh.out_dim)
    cfg = {"pipeline": {"voxel_size": float(z["voxel_size"]), "skip_frames": 1, "clip_bbox_margin": int(z["bbox_margin"]),
                        "clip_masked_weight": float(z["maskedd_weight"]), "max_mask_distance": 6.0}}
    g = Graph(cfg, dataset=ds, clip_model=clip, mask_generator=_SAM(z, ds.W), clip_feat_dim=sh.out_dim)
    full = g.create_feature_map()
    # node table: same nodes, same canonical order, centroids to the last ulps (sequential vs tree f64 sums)
    nxyz = np.asarray(g.full_pcd.points)
    assert nxyz.shape == z["node_xyz"].shape
    assert np.allclose(nxyz, z["node_xyz"], rtol=1e-12, atol=1e-12)
    assert np.allclose(np.asarray(g.full_pcd.colors), z["node_rgb"], rtol=1e-12, atol=1e-12)
    # per-mask embeddings F_p of every frame (fp16-operand encoder vs the reference's fp32 torch run)
    Fp = np.stack([f.numpy() for f in g.frames_feats])
    assert np.abs(Fp - z["F_p"]).max() < 1e-3
    # node features: every stored row within 1e-3 except rows whose winning pixel sits on an exact
    # nearest-node distance tie (implementation-defined in cKDTree; bounded and counted here)
    rows = full[::int(z["row_step"])]
    err = np.abs(rows - z["full_feats_rows"]).max(1)
    assert (err > 1e-3).sum() <= max(2, len(err) // 500), (int((err > 1e-3).sum()), float(err.max()))
    assert np.median(np.abs(full.astype(np.float64).sum(1) - z["full_feats_rowsum"])) < 1e-3
    # N1: the object instances of the reference run (same lists, points to float64 noise: the reference summed
    # its node centroids sequentially, the GPU in tree order)
    off = np.cumsum([0] + [len(np.asarray(p.points)) for p in g.mask_pcds])
    assert np.array_equal(off, z["obj_off"])
    assert np.allclose(np.concatenate([np.asarray(p.points) for p in g.mask_pcds]), z["obj_pts"], rtol=1e-11, atol=1e-11)
    # N2: per-object features; the fp16-operand encoder moves node features by ~1e-3, which can move single rows
    # across the DBSCAN threshold, so this is a tolerance on the cluster means, not on memberships
    mf = np.stack(g.mask_feats)
    err = np.abs(mf - z["mask_feats"]).max(1)
    assert np.median(err) < 2e-3 and (err < 2e-2).mean() >= 0.9, err
    # ... and the chained N2 stage held to the 1e-3 contract: the oracle (sklearn cosine DBSCAN, graph.py:451-488 /
    # graph_utils.py:682-728) run on the GPU's OWN node features and object point sets must give the GPU's object features
    from oracle import hmsg_oracle as O
    ref = O.object_feats([np.asarray(p.points) for p in g.mask_pcds], nxyz, O.build_tree(nxyz), np.asarray(full), float(z["voxel_size"]), sh.out_dim)
    err2 = np.abs(mf - np.stack([np.asarray(r, np.float32).reshape(-1) for r in ref])).max(1)
    # (an object point that sits exactly between two nodes may pick the other one in cKDTree: one swapped row of a cluster)
    assert np.median(err2) < 1e-4 and (err2 < 1e-3).mean() >= 0.95 and err2.max() < 3e-3, err2


def test_query_object_matches_reference_run(engine):
    from holoagent_b200.memory.hmsg.graph.graph import Graph
    z = np.load(os.path.join(GOLD, "ref_query.npz"))
    cases = json.loads(str(z["q_cases"])); words = json.loads(str(z["q_words"]))
    tf, emb, room = z["q_text_feats"], z["q_obj_emb"], z["q_obj_room"]
    NS = types.SimpleNamespace
    g = Graph({"pipeline": {}}, engine=engine, clip_feat_dim=emb.shape[1])
    g.objects = [NS(embedding=emb[i], object_id="obj_%d" % i, room_id="room_%d" % room[i]) for i in range(len(emb))]
    g.rooms = [NS(room_id="room_%d" % r, objects=[o for o in g.objects if o.room_id == "room_%d" % r]) for r in range(3)]
    for ci, (q, rooms, k, neg) in enumerate(cases):
        names = neg if q in neg else [q] + neg
        qf = np.stack([tf[words.index(w)] for w in names])
        ids, rids, sc = g.query_hmsg_object(q, room_ids=rooms, top_k=k, negative_prompt=list(neg), query_feats=qf)
        assert ids == [int(v) for v in z["q%d_ids" % ci]], ci
        assert rids == [int(v) for v in z["q%d_rooms" % ci]], ci
        assert np.allclose(sc, z["q%d_scores" % ci], rtol=0, atol=1e-5), ci


def test_room_and_object_retrieval_match_reference_run(engine):
    """query_hmsg_room / query_room / query_object / identify_object on libhmsg_b200.so == the unmodified reference's
    outputs (tests/golden/ref_retrieval.npz)."""
    from tests.retrieval_golden_cases import check_retrieval_against_reference_run
    assert check_retrieval_against_reference_run(engine) >= 20


def test_room_names_match_reference_run(engine):
    """Appendix A sites room.py:160-168 / :303-306 through generate_room_names on the B200"""
    from tests.retrieval_golden_cases import check_room_names_against_reference_run
    check_room_names_against_reference_run(engine)
