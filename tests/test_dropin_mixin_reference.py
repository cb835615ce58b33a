"""CPU, build container only (skips where /root/reference is absent, e.g. on the GPU box): the mixin form of the drop-in,
`class Graph(B200HotPath, <reference Graph>)`, over the UNMODIFIED reference class imported through the fixture
harness.  The reference's own glue - query_hierarchy_protected (graph.py:3593, what goal_pose_publisher.py:220 calls) and
query_hierarchy_protected_icra - must reach the B200 retrieval cores through normal method resolution and return exactly
what the pure reference returns on the same graph.  The engine is the numpy test double (tests/fake_engine.py): this
checks the host glue / MRO, the kernels are checked on the GPU against the same fixtures."""
import contextlib
import io
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/fsr_vln"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
NS = types.SimpleNamespace
D = 128


def _ref_graph():
    """import the reference behind the fixture harness (patches sys.modules / torch: only ever called in the child process)"""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import ref_shims
    ref_shims.install()
    import memory.hmsg.graph.graph as rg
    return rg


def _populate(g, rs, table):
    n_rooms, n_obj = 4, 60
    emb = rs.randn(n_obj, D).astype(np.float32); emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    room_of = rs.randint(0, n_rooms, n_obj)
    names = ["kitchen", "office", "hall", "lab"]
    g.objects = [NS(embedding=emb[i], object_id="0_%d_%d" % (room_of[i], i), room_id="0_%d" % room_of[i], name="o%d" % i) for i in range(n_obj)]
    g.rooms = [NS(room_id="0_%d" % r, name=names[r], embeddings=[table["v%d_%d" % (r, k)] for k in range(3)],
                  objects=[o for o in g.objects if o.room_id == "0_%d" % r]) for r in range(n_rooms)]
    g.floors = [NS(floor_id="0", rooms=g.rooms[:2], floor_zero_level=0.0), NS(floor_id="1", rooms=g.rooms[2:], floor_zero_level=3.0)]
    g.clip_model, g.clip_feat_dim, g.cfg = None, D, None


def test_reference_glue_reaches_b200_cores(tmp_path):
    """runs `_check` in a child process: the import harness replaces open3d / faiss / torch.Tensor.cuda process-wide"""
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MIXIN_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


class _Patch:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)

    def chdir(self, p):
        os.chdir(str(p))


def _check(ref_graph, tmp_path, monkeypatch):
    from holoagent_b200.memory.hmsg.graph.graph import B200HotPath, dropin_graph_class
    from tests.fake_engine import OracleRetrievalEngine
    rs = np.random.RandomState(7)
    words = ["kitchen", "office", "hall", "lab", "chair", "mug", "unknown", "floor 0", "floor 1", "monitor", "wall", "speaker"] + \
        ["background", "divider", "ledge", "pillar", "tape", "stairs", "door", "doors", "stair", "window", "glass", "railing", "glass doors",
         "whiteboard", "sliding door", "carpet", "ceiling", "curtain"] + ["v%d_%d" % (r, k) for r in range(4) for k in range(3)]
    tf = rs.randn(len(words), D).astype(np.float32); tf /= np.linalg.norm(tf, axis=1, keepdims=True)
    table = dict(zip(words, tf))
    monkeypatch.setattr(ref_graph, "get_text_feats_multiple_templates", lambda q, m, dim: np.stack([table[w] for w in q]))
    parses = {"chair in the kitchen": (None, "kitchen", "chair"), "a mug in the lab on floor 2": ("2", "lab", "mug"), "find the mug": (None, "unknown", "mug")}
    monkeypatch.setattr(ref_graph, "parse_hier_query_use_prompt_insentence_parse", lambda cfg, q: parses[q])
    monkeypatch.setattr(ref_graph, "parse_hier_query_use_prompt_insentence_parse_icra", lambda cfg, q: parses[q])
    monkeypatch.chdir(tmp_path)                       # the reference appends to room_obj_query_log.txt in the cwd

    Drop = dropin_graph_class(ref_graph.Graph)
    assert issubclass(Drop, B200HotPath) and issubclass(Drop, ref_graph.Graph)
    # hot-path methods resolve to the mixin, the glue to the reference
    assert Drop.query_hmsg_object is B200HotPath.query_hmsg_object and Drop.query_hmsg_room is B200HotPath.query_hmsg_room
    assert Drop.create_feature_map is B200HotPath.create_feature_map
    assert Drop.query_hierarchy_protected is ref_graph.Graph.query_hierarchy_protected
    assert Drop.load_hmsg_graph is ref_graph.Graph.load_hmsg_graph

    pure = ref_graph.Graph.__new__(ref_graph.Graph)
    drop = Drop.__new__(Drop)                         # the reference constructor loads SAM / CLIP checkpoints: state is set by hand
    for g in (pure, drop):
        _populate(g, np.random.RandomState(3), table)
    eng = OracleRetrievalEngine()
    drop._b200_init(eng, None)
    drop.text_feats_fn = lambda texts: ref_graph.get_text_feats_multiple_templates(texts, None, D)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        for ins, k in [("chair in the kitchen", 3), ("a mug in the lab on floor 2", 2), ("find the mug", 4)]:
            for fn in ("query_hierarchy_protected", "query_hierarchy_protected_icra"):
                a = getattr(pure, fn)(ins, top_k=k, use_gpt=False)
                n0 = eng.index_sets
                b = getattr(drop, fn)(ins, top_k=k, use_gpt=False)
                assert eng.index_sets > n0 or k == 0          # the B200 engine really served the request
                assert (a[0] is None) == (b[0] is None) and (a[0] is None or a[0].floor_id == b[0].floor_id)
                assert [r.room_id for r in a[1]] == [r.room_id for r in b[1]], (ins, fn)
                assert [o.object_id for o in a[2]] == [o.object_id for o in b[2]], (ins, fn)
                assert np.allclose(a[3].get("object_scores", []), b[3].get("object_scores", []), atol=1e-6)


if __name__ == "__main__":
    _check(_ref_graph(), sys.argv[1], _Patch())
    print("MIXIN_OK")
