"""GPU: the reference-facing drop-in modules end to end (BASELINE configs[0]-style plumbing
case, scaled down): Graph.create_feature_map + query_* against the CPU oracle pipeline."""
import numpy as np
import pytest
import torch

from holoagent_b200 import synth
from oracle import hmsg_oracle as O

pytestmark = pytest.mark.gpu

H, W, NF, M, D = 120, 160, 5, 4, 512


class _DS:
    """RGBDDataset-shaped object: dataset[i] -> (rgb, depth, pose, rgb_K, depth_K)."""

    def __init__(self):
        self.depth, self.rgb, self.T, self.K = synth.make_frames_np(np.arange(NF) * 6, H, W)
        self.depth_intrinsics = self.K
        self.scale = 1000.0

    def __len__(self):
        return NF

    def __getitem__(self, i):
        return self.rgb[i], self.depth[i], self.T[i], self.K, self.K


class _SAM:
    def __init__(self, ds):
        self.q = [synth.make_masks(i, ds.depth[i], M) for i in range(NF)]
        self.i = 0

    def generate(self, image):
        m = self.q[self.i % NF]
        self.i += 1
        return m


class _Obj:
    def __init__(self, oid, rid, emb):
        self.object_id, self.room_id, self.embedding = oid, rid, emb


class _Room:
    def __init__(self, rid, objs, embs):
        self.room_id, self.objects, self.embeddings, self.name = rid, objs, embs, rid


def test_graph_create_feature_map_and_queries(engine):
    from holoagent_b200.memory.hmsg.graph.graph import Graph
    from holoagent_b200.memory.hmsg.utils.clip_utils import B200ClipModel
    ds = _DS()
    sd = synth.make_vit_weights()
    clip = B200ClipModel(engine, sd)
    cfg = {"pipeline": {"voxel_size": 0.05, "skip_frames": 1, "clip_bbox_margin": 50, "clip_masked_weight": 0.4418, "max_mask_distance": 10000}}
    g = Graph(cfg, dataset=ds, clip_model=clip, mask_generator=_SAM(ds), clip_feat_dim=D)
    # the synthetic scene is small: use the reference's own radius filter parameters anyway
    full = g.create_feature_map()
    n_nodes = full.shape[0]
    # ---- oracle pipeline on the same inputs (node table from the GPU: stage-wise for ties)
    geo = O.build_geometry(ds.depth, ds.rgb, ds.T, ds.K, 1000.0, 0.05, 1000, 1.0)
    assert n_nodes == len(geo["keep"])
    nxyz = np.asarray(g.full_pcd.points)
    assert np.allclose(nxyz, geo["node_xyz"], rtol=1e-12, atol=1e-12)
    if n_nodes:
        tree = O.build_tree(nxyz)
        sum_f = torch.zeros(n_nodes, D); cnt = torch.zeros(n_nodes, 1)
        sam = _SAM(ds)
        for f in range(NF):
            masks = sam.generate(ds.rgb[f])
            crops = O.crop_all_bounding_boxs(ds.rgb[f], masks, True, 50) + O.crop_all_bounding_boxs(ds.rgb[f], masks, False, 50) + [ds.rgb[f]]
            x = torch.stack([O.clip_preprocess(c) for c in crops])
            fe = O.get_img_feats_batch_tensor(sd, x)
            Fp = O.fuse_mask_feats(fe[:M], fe[M:2 * M], fe[2 * M:], 0.4418)
            assert np.allclose(g.frames_feats[f].numpy(), Fp, atol=1e-3)
            # A9b (graph.py:1119-1136): the view embedding room building needs == get_img_feats(full frame) == the F_g row the build kept
            assert np.allclose(g.frame_global_feats[f], fe[2 * M], atol=1e-3)
            gidx, _ = engine.pixel_to_node(f, want_dist=False)
            O.ingest_frame(sum_f, cnt, tree, n_nodes, ds.depth[f], ds.rgb[f], ds.T[f], ds.K, 1000.0, Fp, np.stack([m["segmentation"] for m in masks]),
                           idx=gidx[(ds.depth[f] > 0).reshape(-1)])
        ref = O.finalize_node_feats(sum_f, cnt)
        hit = cnt.numpy().reshape(-1) > 0
        assert np.allclose(full[hit], ref[hit], atol=2e-3)
    assert len(g.frames_pcd) == NF and len(g.frames_pcd[0]) == M
    # ---- retrieval through the reference-shaped methods with injected objects / rooms
    rs = np.random.RandomState(3)
    embs = rs.randn(60, D).astype(np.float32); embs /= np.linalg.norm(embs, axis=1, keepdims=True)
    g.objects = [_Obj(100 + i, f"room_{i % 3}", embs[i].astype(np.float64)) for i in range(60)]
    g.rooms = [_Room(f"room_{r}", [o for o in g.objects if o.room_id == f"room_{r}"], [embs[r * 5 + k] for k in range(5)]) for r in range(3)]
    q = (embs[7] * 0.8)[None]
    neg = rs.randn(2, D).astype(np.float32) * 0.03
    qf = np.concatenate([q, neg])
    ids, rooms, scores = g.query_hmsg_object("chair", top_k=3, negative_prompt=["background", "wall"], query_feats=qf)
    top, osc = O.query_object_core(qf, embs, 0, 3, True)
    assert ids == list(top) and np.allclose(scores, osc, atol=1e-5)
    assert rooms == [int(t) % 3 for t in top]
    ids2, rooms2 = g.query_object("chair", room_ids=[1], top_k=2, query_feats=q)
    sub = [i for i in range(60) if i % 3 == 1]
    t2, _ = O.query_topk(q[0], embs[sub], 2)
    assert ids2 == [sub[i] for i in t2] and rooms2 == [1, 1]
    assert g.query_graph("chair", query_feats=q).object_id == 107
    assert g.identify_object(embs[9], embs[:20], [f"c{i}" for i in range(20)]) == "c9"
    order, sc = O.rooms_by_view_embedding(q[0], [np.stack(r.embeddings) for r in g.rooms])
    assert g.query_hmsg_room("kitchen", query_method="view_embedding", query_feats=q) == [int(o) for o in order][:5]
    assert g.query_room("kitchen", query_feats=q) == [int(o) for o in order][:3]


def test_extract_feats_per_pixel_dropin(engine):
    from holoagent_b200.memory.hmsg.utils.clip_utils import B200ClipModel, get_img_feats
    from holoagent_b200.perception.models.sam_clip_feats_extractor import extract_feats_per_pixel
    ds = _DS()
    sd = synth.make_vit_weights()
    clip = B200ClipModel(engine, sd)
    sam = _SAM(ds)
    outfeat, Fp, masks, Fg = extract_feats_per_pixel(ds.rgb[0], sam, clip, None, clip_feat_dim=D, bbox_margin=50, maskedd_weight=0.4418)
    assert outfeat.shape == (H, W, D) and outfeat.dtype == torch.float16 and Fp.shape == (M, D) and Fg.shape == (1, D)
    crops = O.crop_all_bounding_boxs(ds.rgb[0], masks, True, 50) + O.crop_all_bounding_boxs(ds.rgb[0], masks, False, 50) + [ds.rgb[0]]
    fe = O.get_img_feats_batch_tensor(sd, torch.stack([O.clip_preprocess(c) for c in crops]))
    oFp = O.fuse_mask_feats(fe[:M], fe[M:2 * M], fe[2 * M:], 0.4418)
    assert np.allclose(Fp.numpy(), oFp, atol=1e-3) and np.allclose(Fg, fe[2 * M:], atol=1e-3)
    dense = O.pixel_feature_map(Fp.numpy(), np.stack([m["segmentation"] for m in masks]), H, W).reshape(H, W, D)
    assert np.allclose(outfeat.float().numpy(), dense.float().numpy(), atol=1e-3)
    f2 = get_img_feats(ds.rgb[0], O.clip_preprocess, clip)
    assert np.allclose(f2, fe[2 * M:], atol=1e-3)


def test_query_graph_loaded_from_reference_json(engine, tmp_path):
    """N4: a graph written in the reference's JSON schema is served by the B200 retrieval path."""
    from holoagent_b200.memory.hmsg.graph.graph import Graph
    from tests.test_store import _write_graph
    embs = _write_graph(str(tmp_path), d=512)
    g = Graph({"pipeline": {}}, engine=engine, clip_feat_dim=512).load_hmsg_graph(str(tmp_path))
    assert len(g.objects) == 6 and len(g.rooms) == 2          # object 5 has no embedding
    # self.objects follows the reference loader: sorted file names = id strings (graph.py:1928-1930)
    keep = sorted((i for i in range(7) if i != 5), key=lambda i: f"0_{i % 2}_{i}")
    assert [o.object_id for o in g.objects] == [f"0_{i % 2}_{i}" for i in keep]
    E = embs[keep].astype(np.float32)
    q = (embs[2] * 0.5).astype(np.float32)[None]
    ids, rooms, scores = g.query_hmsg_object("x", top_k=3, query_feats=q)
    top, osc = O.query_topk(q[0], E, 3)
    assert ids == [int(t) for t in top] and np.allclose(scores, osc, rtol=1e-5, atol=1e-4)
    assert rooms == [keep[int(t)] % 2 for t in top]


def test_view_and_room_retrieval_variants(engine, tmp_path):
    """Appendix A sites served from a reference-schema graph: global view top-24 (graph.py:2864-2897), re-match inside
    a view (:2977-2984), per-room max over view embeddings (:3250-3272, :3345-3359), floor by name (:2248-2251)."""
    import json, os
    from holoagent_b200.memory.hmsg.graph.graph import Graph
    from tests.test_store import _write_graph
    embs = _write_graph(str(tmp_path), d=512)
    g = Graph({"pipeline": {}}, engine=engine, clip_feat_dim=512).load_hmsg_graph(str(tmp_path))
    rs = np.random.RandomState(3)
    q = rs.randn(1, 512).astype(np.float32) * 0.05
    # global view retrieval: numpy restatement of the reference lines
    ids, em = [], []
    for r in g.rooms:
        ids.extend(r.sample_images); em.extend(r.clip_embeddings)
    sims = np.dot(q[0], np.stack(em).astype(np.float32).T)
    top_k = min(24, sims.shape[0])
    top_idx = np.argsort(sims)[-top_k:][::-1]
    best, top_ids, sc = g.query_views("x", query_feats=q)
    assert best == ids[int(np.argmax(sims))] and top_ids == [ids[i] for i in top_idx]
    assert np.allclose(sc, sims[top_idx], rtol=1e-5, atol=1e-5)
    # re-match inside a view
    in_view = ["0_0_2", "0_1_3", "0_0_6"]
    by_id = {o.object_id: o for o in g.objects}
    e = np.stack([by_id[i].embedding for i in in_view]).astype(np.float32)
    oid, s = g.rematch_in_view("x", in_view, query_feats=q)
    ref = np.dot(q[0], e.T)
    assert oid == in_view[int(np.argmax(ref))] and abs(s - ref.max()) < 1e-5
    # rooms by view embedding
    room_max = [np.dot(q[0], np.stack(r.embeddings).astype(np.float32).T).max() for r in g.rooms]
    order = sorted(range(len(g.rooms)), key=lambda r: room_max[r], reverse=True)
    assert g.query_hmsg_room("kitchen", query_feats=q) == order[:5]
    assert g.query_room("kitchen", query_feats=q) == order[:3]
    # floor by name
    fl = rs.randn(3, 512).astype(np.float32)
    assert g.query_floor("x", fl, query_feats=q) == int(np.argsort(np.dot(q, fl.T)[0])[::-1][0])


def test_rgb_larger_than_depth_uses_both_resizes(engine):
    """Datasets whose colour image is larger than the depth image: the point colours come from cv2 INTER_AREA
    (generic.py:98-104), the feature pass from PIL bicubic (graph.py:378-379).  Limit (2) of round 1 is gone."""
    import cv2
    from PIL import Image
    from holoagent_b200.memory.hmsg.graph.graph import Graph
    from holoagent_b200.memory.hmsg.utils.clip_utils import B200ClipModel
    base = _DS()
    rs = np.random.RandomState(5)
    big = [np.clip(cv2.resize(base.rgb[i], (2 * W, 2 * H), interpolation=cv2.INTER_CUBIC).astype(np.int32) + rs.randint(-20, 21, (2 * H, 2 * W, 3)), 0, 255).astype(np.uint8)
           for i in range(NF)]

    class _DS2(_DS):
        def __getitem__(self, i):
            return Image.fromarray(big[i]), Image.fromarray(self.depth[i]), self.T[i], self.K, self.K

    ds = _DS2()
    area = np.stack([cv2.resize(big[i], (W, H), interpolation=cv2.INTER_AREA) for i in range(NF)])
    cubic = np.stack([np.asarray(Image.fromarray(big[i]).resize((W, H))) for i in range(NF)])
    assert (area != cubic).mean() > 0.2                      # the two images really differ
    sd = synth.make_vit_weights()
    cfg = {"pipeline": {"voxel_size": 0.05, "skip_frames": 1, "clip_bbox_margin": 50, "clip_masked_weight": 0.4418, "max_mask_distance": 10000}}
    g = Graph(cfg, dataset=ds, clip_model=B200ClipModel(engine, sd), mask_generator=_SAM(base), clip_feat_dim=D)
    g.merge_objects = False
    g.create_feature_map()
    geo = O.build_geometry(base.depth, area, base.T, base.K, 1000.0, 0.05, 1000, 1.0)
    assert np.allclose(np.asarray(g.full_pcd.points), geo["node_xyz"], rtol=1e-12, atol=1e-12)
    assert np.allclose(np.asarray(g.full_pcd.colors), geo["node_rgb"], rtol=1e-12, atol=1e-12)
    sam = _SAM(base)
    for f in range(2):
        masks = sam.generate(cubic[f])
        crops = O.crop_all_bounding_boxs(cubic[f], masks, True, 50) + O.crop_all_bounding_boxs(cubic[f], masks, False, 50) + [cubic[f]]
        fe = O.get_img_feats_batch_tensor(sd, torch.stack([O.clip_preprocess(c) for c in crops]))
        assert np.allclose(g.frames_feats[f].numpy(), O.fuse_mask_feats(fe[:M], fe[M:2 * M], fe[2 * M:], 0.4418), atol=1e-3)
