"""CPU: host-side glue of the Graph mirror's retrieval methods (room selection, negative prompts, id mapping,
label tie window, view retrieval) against (1) the outputs the UNMODIFIED reference `Graph.query_hmsg_object` returned
in this container (tests/golden/ref_query.npz) and (2) numpy restatements of the reference lines.  The engine is a
numpy test double (tests/fake_engine.py); the same assertions run against libhmsg_b200.so in test_gpu_*.py."""
import json
import os
import types

import numpy as np

from holoagent_b200.memory.hmsg.graph.graph import Graph
from tests.fake_engine import OracleRetrievalEngine
from tests.test_store import _write_graph

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NS = types.SimpleNamespace


def test_query_hmsg_object_matches_reference_run():
    z = np.load(os.path.join(GOLD, "ref_query.npz"))
    cases = json.loads(str(z["q_cases"])); words = json.loads(str(z["q_words"]))
    tf, emb, room = z["q_text_feats"], z["q_obj_emb"], z["q_obj_room"]
    g = Graph({"pipeline": {}}, engine=OracleRetrievalEngine(), clip_feat_dim=emb.shape[1])
    g.objects = [NS(embedding=emb[i], object_id="obj_%d" % i, room_id="room_%d" % room[i]) for i in range(len(emb))]
    g.rooms = [NS(room_id="room_%d" % r, objects=[o for o in g.objects if o.room_id == "room_%d" % r]) for r in range(3)]
    assert len(cases) >= 4
    for ci, (q, rooms, k, neg) in enumerate(cases):
        names = neg if q in neg else [q] + neg
        qf = np.stack([tf[words.index(w)] for w in names])
        ids, rids, sc = g.query_hmsg_object(q, room_ids=rooms, top_k=k, negative_prompt=list(neg), query_feats=qf)
        assert ids == [int(v) for v in z["q%d_ids" % ci]], ci
        assert rids == [int(v) for v in z["q%d_rooms" % ci]], ci
        assert np.allclose(sc, z["q%d_scores" % ci], rtol=0, atol=1e-5), ci
        ids2, rids2 = g.query_object(q, room_ids=rooms, top_k=k, negative_prompt=list(neg), query_feats=qf)
        assert (ids2, rids2) == (ids, rids)


def test_index_is_cached_per_selection():
    rs = np.random.RandomState(0)
    emb = rs.randn(12, 128)
    eng = OracleRetrievalEngine()
    g = Graph({"pipeline": {}}, engine=eng, clip_feat_dim=128)
    g.objects = [NS(embedding=emb[i], object_id="o%d" % i, room_id="r%d" % (i % 2)) for i in range(12)]
    g.rooms = [NS(room_id="r%d" % r, objects=[o for o in g.objects if o.room_id == "r%d" % r]) for r in range(2)]
    q = rs.randn(1, 128).astype(np.float32)
    g.query_hmsg_object("x", top_k=3, query_feats=q); g.query_hmsg_object("y", top_k=2, query_feats=q)
    assert eng.index_sets == 1                       # same object selection: the device matrix is reused
    g.query_hmsg_object("x", room_ids=[1], top_k=3, query_feats=q)
    assert eng.index_sets == 2
    ids, rids, _ = g.query_hmsg_object("x", room_ids=[1], top_k=50, query_feats=q)     # top_k larger than the selection
    assert len(ids) == 6 and set(rids) == {1} and all(g.objects[i].room_id == "r1" for i in ids)


def test_room_view_floor_variants(tmp_path):
    embs = _write_graph(str(tmp_path), d=128)
    g = Graph({"pipeline": {}}, engine=OracleRetrievalEngine(), clip_feat_dim=128).load_hmsg_graph(str(tmp_path))
    assert len(g.objects) == 6 and len(g.rooms) == 2 and len(g.floors) == 1
    rs = np.random.RandomState(3)
    q = rs.randn(1, 128).astype(np.float32) * 0.05
    # plain top-5 (graph.py:2196-2200) and class labelling (graph.py:1452-1454)
    E = np.stack([o.embedding for o in g.objects]).astype(np.float32)
    assert g.query_graph("x", query_feats=q) is g.objects[int(np.argsort(np.dot(q, E.T)[0])[::-1][0])]
    labels = rs.randn(7, 128).astype(np.float32)
    assert g.identify_object(E[2], labels, list("abcdefg")) == "abcdefg"[int(np.argmax(np.dot(E[2:3], labels.T)))]
    # global view retrieval (graph.py:2864-2897)
    ids, em = [], []
    for r in g.rooms:
        ids.extend(r.sample_images); em.extend(r.clip_embeddings)
    sims = np.dot(q[0], np.stack(em).astype(np.float32).T)
    top_idx = np.argsort(sims)[-min(24, len(sims)):][::-1]
    best, top_ids, sc = g.query_views("x", query_feats=q)
    assert best == ids[int(np.argmax(sims))] and top_ids == [ids[i] for i in top_idx] and np.allclose(sc, sims[top_idx], atol=1e-6)
    # re-match inside a view (graph.py:2977-2984)
    in_view = ["0_0_2", "0_1_3", "0_0_6"]
    by_id = {o.object_id: o for o in g.objects}
    ref = np.dot(q[0], np.stack([by_id[i].embedding for i in in_view]).astype(np.float32).T)
    oid, s = g.rematch_in_view("x", in_view, query_feats=q)
    assert oid == in_view[int(np.argmax(ref))] and abs(s - ref.max()) < 1e-6
    assert g.rematch_in_view("x", [], query_feats=q) == (None, None)
    # rooms by view embedding: per-room max, top 5 / top 3 (graph.py:3250-3272, :3345-3359); "unknown" text -> top 10
    room_max = [np.dot(q[0], np.stack(r.embeddings).astype(np.float32).T).max() for r in g.rooms]
    order = sorted(range(len(g.rooms)), key=lambda r: room_max[r], reverse=True)
    assert g.query_hmsg_room("kitchen", query_feats=q) == order[:5]
    assert g.query_room("kitchen", query_feats=q) == order[:3]
    assert g.query_hmsg_room("unknown room", query_feats=q) == order[:10]
    # rooms by label with the 1e-3 tie window (graph.py:3204-3230)
    names = np.stack([q[0] * 2.0, q[0] * 2.0 + 1e-5]).astype(np.float32)          # two names within the window
    assert sorted(g.query_hmsg_room("kitchen", query_method="label", query_feats=q, room_name_feats=names)) == [0, 1]
    names[1] = -names[1]
    assert g.query_hmsg_room("kitchen", query_method="label", query_feats=q, room_name_feats=names) == [0]
    # floor by name (graph.py:2248-2251)
    fl = rs.randn(3, 128).astype(np.float32)
    assert g.query_floor("x", fl, query_feats=q) == int(np.argsort(np.dot(q, fl.T)[0])[::-1][0])
    assert g.query_floor("x", fl, query_feats=q, zero_level_order_ids=[7, 8, 9]) in (7, 8, 9)


def test_room_and_object_retrieval_match_reference_run():
    from tests.retrieval_golden_cases import check_retrieval_against_reference_run
    assert check_retrieval_against_reference_run(OracleRetrievalEngine()) >= 20


def test_room_names_match_reference_run():
    from tests.retrieval_golden_cases import check_room_names_against_reference_run
    check_room_names_against_reference_run(OracleRetrievalEngine())


def test_two_graphs_sharing_one_engine_do_not_see_each_others_index():
    rs = np.random.RandomState(5)
    eng = OracleRetrievalEngine()
    graphs = []
    for n in (9, 9):                                    # same sizes: only the engine-side tag tells them apart
        emb = rs.randn(n, 128)
        g = Graph({"pipeline": {}}, engine=eng, clip_feat_dim=128)
        g.objects = [NS(embedding=emb[i], object_id="o%d" % i, room_id="r0") for i in range(n)]
        g.rooms = [NS(room_id="r0", objects=g.objects)]
        graphs.append((g, emb))
    q = rs.randn(1, 128).astype(np.float32)
    for _ in range(2):
        for g, emb in graphs:
            ids, _, sc = g.query_hmsg_object("x", top_k=1, query_feats=q)
            assert ids == [int(np.argmax(emb.astype(np.float32) @ q[0]))]
    # in-place edits need an explicit invalidation
    g, emb = graphs[0]
    g.query_hmsg_object("x", top_k=1, query_feats=q)
    g.objects[3].embedding = q[0] * 10.0
    g.invalidate_index()
    assert g.query_hmsg_object("x", top_k=1, query_feats=q)[0] == [3]


def test_dropin_mode_reads_the_visual_tower_of_the_reference_clip_model():
    """Drop-in mode: the reference constructor leaves an open_clip model in self.clip_model (graph.py:98-119); the first
    build hands its visual tower to the engine with the shape read off the state dict."""
    import torch
    from holoagent_b200 import synth
    shape = synth.VitB32Shape(image=64, patch=16, width=128, layers=2, heads=2, mlp=256, out_dim=64)
    sd = synth.make_vit_weights(shape)

    class _Attn:
        num_heads = 2

    class _Block:
        attn = _Attn()

    class _Visual:
        image_size = (64, 64)
        transformer = types.SimpleNamespace(resblocks=[_Block(), _Block()])

        def state_dict(self):
            return {k: v.half() for k, v in sd.items()}      # precision='fp16' checkpoints

    got = {}

    class _Eng:
        vit = None

        def encoder_load(self, state, **kw):
            got["kw"] = kw
            got["sd"] = state

    g = Graph({"pipeline": {}}, engine=_Eng(), clip_feat_dim=64)
    g.clip_model = types.SimpleNamespace(visual=_Visual())
    g._ensure_encoder()
    assert got["kw"] == dict(image=64, patch=16, width=128, layers=2, heads=2, mlp=256, out_dim=64, quick_gelu=False)
    assert all(t.dtype == torch.float32 for t in got["sd"].values()) and torch.equal(got["sd"]["proj"], sd["proj"])
    # OpenAI / *-quickgelu towers and ViT-H/14 (16 heads of 80) are recognised from the modules, not guessed from the width
    class QuickGELU:
        pass

    _Block.mlp = types.SimpleNamespace(gelu=QuickGELU())
    _Attn.num_heads = 4
    g._ensure_encoder()
    assert got["kw"]["quick_gelu"] is True and got["kw"]["heads"] == 4
    g.clip_model = types.SimpleNamespace()                    # no visual tower and nothing loaded: loud failure
    try:
        g._ensure_encoder()
        assert False
    except RuntimeError as e:
        assert "no encoder loaded" in str(e)
