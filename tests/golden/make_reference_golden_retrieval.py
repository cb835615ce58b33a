"""Retrieval-only fixtures from the UNMODIFIED reference (fsr_vln, /root/reference), imported behind the same
harness as make_reference_golden.py.  What executes is the reference's own source:
  * Graph.query_hmsg_room   memory/hmsg/graph/graph.py:3164-3272  ("view_embedding", "label", "unknown" text)
  * Graph.query_room        graph.py:3277-3359
  * Graph.query_object      graph.py:3363-3481
  * Graph.identify_object   graph.py:1441-1454
  * Graph.query_hierarchy_protected / _icra   graph.py:3593-3716 / :3484-3591 (fast path, use_gpt=False; the LLM parse of the
    instruction - llm_utils, a network call - is replaced by a lookup table, exactly like the text tower)
  * Graph.query_floor       graph.py:2216-2252
  * a multi-floor graph queried with floor_id=-1 (rooms "0_1" and "1_1" collapse in the view-embedding branch, :3259-3272)
The CLIP text tower is out of scope: `get_text_feats_multiple_templates` is replaced by a seeded lookup table keyed
by string (the same substitution as in make_reference_golden.py).  Output: tests/golden/ref_retrieval.npz

    python tests/golden/make_reference_golden_retrieval.py        (container only; deterministic)
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
import memory.hmsg.graph.graph as ref_graph  # noqa: E402   (reference)

NS = types.SimpleNamespace
D = 256


def main():
    rs = np.random.RandomState(2024)
    words = ["kitchen", "bedroom", "office", "corridor", "unknown area", "chair", "mug", "plant", "background", "floor 0", "floor 1",
             "divider", "ledge", "pillar", "tape", "stairs", "door", "doors", "stair", "window", "glass", "railing", "glass doors", "whiteboard",
             "sliding door", "carpet", "ceiling", "curtain", "monitor", "wall", "speaker"]
    tf = rs.randn(len(words), D).astype(np.float32); tf /= np.linalg.norm(tf, axis=1, keepdims=True)
    room_names = ["kitchen", "bed room", "office", "hall", "kitchen"]          # two rooms share a name -> identical label scores (tie window)
    name_tf = {"kitchen": tf[0], "bed room": tf[1] * 0.9 + tf[0] * 0.1, "office": tf[2], "hall": tf[3]}
    table = {w: tf[i] for i, w in enumerate(words)}
    table.update(name_tf)
    ref_graph.get_text_feats_multiple_templates = lambda q, m, dim: np.stack([table[w] for w in q])

    g = ref_graph.Graph.__new__(ref_graph.Graph)
    g.clip_model, g.clip_feat_dim = None, D
    n_rooms, n_obj = 5, 120
    views = [rs.randint(3, 25) for _ in range(n_rooms)]
    room_embs = []
    for r in range(n_rooms):
        e = rs.randn(views[r], D).astype(np.float32); e /= np.linalg.norm(e, axis=1, keepdims=True)
        e = (0.7 * e + 0.3 * tf[r % 4]).astype(np.float32)
        room_embs.append(e)
    emb = rs.randn(n_obj, D).astype(np.float32); emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    emb = (0.6 * emb + 0.4 * tf[5 + rs.randint(0, 4, n_obj)]).astype(np.float32)
    room_of = rs.randint(0, n_rooms, n_obj)
    g.objects = [NS(embedding=emb[i], object_id="0_%d_%d" % (room_of[i], i), room_id="0_%d" % room_of[i], name="obj%d" % i) for i in range(n_obj)]
    g.rooms = [NS(room_id="0_%d" % r, name=room_names[r], embeddings=list(room_embs[r]),
                  objects=[o for o in g.objects if o.room_id == "0_%d" % r]) for r in range(n_rooms)]
    g.floors = [NS(floor_id="0", rooms=g.rooms[:3]), NS(floor_id="1", rooms=g.rooms[3:])]

    out = {"words": np.array(json.dumps(words)), "text_feats": tf, "room_names": np.array(json.dumps(room_names)),
           "room_name_feats": np.stack([name_tf[n] for n in room_names]), "room_view_counts": np.array(views),
           "room_embs": np.concatenate(room_embs), "obj_emb": emb, "obj_room": room_of}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        room_cases = [("kitchen", -1, "view_embedding"), ("office", -1, "view_embedding"), ("unknown area", -1, "view_embedding"),
                      ("bedroom", 0, "view_embedding"), ("corridor", 1, "view_embedding"),
                      ("kitchen", -1, "label"), ("bedroom", -1, "label"), ("office", 0, "label"), ("unknown area", -1, "label")]
        for ci, (q, fl, method) in enumerate(room_cases):
            out["hmsg_room_%d" % ci] = np.array(g.query_hmsg_room(q, floor_id=fl, query_method=method), np.int64)
            out["room_%d" % ci] = np.array(g.query_room(q, floor_id=fl, query_method=method), np.int64)
        obj_cases = [("chair", [], [0, 1], 4, []), ("mug", [], [2], 3, ["background"]), ("plant", [], [0, 3, 4], 6, ["plant", "chair", "background"]),
                     ("chair", 0, [1, 2], 5, []), ("mug", 1, [0], 2, ["background", "plant"])]
        for ci, (q, _, rooms, k, neg) in enumerate(obj_cases):
            fl = obj_cases[ci][1] if obj_cases[ci][1] != [] else -1
            ids, rids = g.query_object(q, floor_id=fl, room_ids=rooms, top_k=k, negative_prompt=list(neg))
            out["obj_%d_ids" % ci] = np.array(ids, np.int64); out["obj_%d_rooms" % ci] = np.array(rids, np.int64)
        # ---- the entry point the robot calls (goal_pose_publisher.py:220): instruction -> (floor, rooms, objects, res_dict)
        parses = {"go to the chair in the kitchen": (None, "kitchen", "chair"),
                  "find a mug in the office on floor 1": ("1", "office", "mug"),
                  "bring me the plant": (None, "unknown area", "plant"),
                  "the mug in the bedroom upstairs": ("floor 1", "bedroom", "mug")}
        ref_graph.parse_hier_query_use_prompt_insentence_parse = lambda cfg, q: parses[q]
        ref_graph.parse_hier_query_use_prompt_insentence_parse_icra = lambda cfg, q: parses[q]
        g.cfg = None
        for f, lvl in zip(g.floors, (0.0, 3.1)):
            f.floor_zero_level = lvl
        cwd = os.getcwd(); os.chdir("/tmp")            # the reference appends to room_obj_query_log.txt in the cwd
        hier = []
        for hi, (ins, tk) in enumerate([("go to the chair in the kitchen", 3), ("find a mug in the office on floor 1", 2), ("bring me the plant", 4),
                                        ("the mug in the bedroom upstairs", 1)]):
            for fn in ("query_hierarchy_protected", "query_hierarchy_protected_icra"):
                fl, rooms, objs, res = getattr(g, fn)(ins, top_k=tk, use_gpt=False)
                hier.append({"fn": fn, "instruction": ins, "top_k": tk, "parse": parses[ins], "floor": None if fl is None else fl.floor_id,
                             "rooms": [r.room_id for r in rooms], "objects": [o.object_id for o in objs], "negative_labels": res["negative_labels"]})
        os.chdir(cwd)
        out["hier_cases"] = np.array(json.dumps(hier))
        # ---- multi-floor graph, floor_id = -1: duplicate trailing room numbers
        g2 = ref_graph.Graph.__new__(ref_graph.Graph)
        g2.clip_model, g2.clip_feat_dim = None, D
        mf_ids = ["0_0", "0_1", "1_0", "1_1", "1_2"]
        mf_embs = []
        for r in range(len(mf_ids)):
            e = rs.randn(4 + r, D).astype(np.float32); e /= np.linalg.norm(e, axis=1, keepdims=True)
            mf_embs.append((0.7 * e + 0.3 * tf[(r + 1) % 4]).astype(np.float32))
        g2.rooms = [NS(room_id=mf_ids[r], name="room", embeddings=list(mf_embs[r]), objects=[]) for r in range(len(mf_ids))]
        g2.floors = [NS(floor_id="0", rooms=g2.rooms[:2]), NS(floor_id="1", rooms=g2.rooms[2:])]
        g2.objects = []
        out["mf_room_ids"] = np.array(json.dumps(mf_ids)); out["mf_view_counts"] = np.array([len(e) for e in mf_embs]); out["mf_embs"] = np.concatenate(mf_embs)
        for qi, q in enumerate(["kitchen", "office", "unknown area"]):
            out["mf_hmsg_room_%d" % qi] = np.array(g2.query_hmsg_room(q, floor_id=-1, query_method="view_embedding"), np.int64)
            out["mf_room_%d" % qi] = np.array(g2.query_room(q, floor_id=-1, query_method="view_embedding"), np.int64)
        classes = ["chair", "mug", "plant", "background"]
        label_feats = np.stack([table[c] for c in classes])
        out["identify"] = np.array(json.dumps([g.identify_object(emb[i], label_feats, classes) for i in range(0, n_obj, 7)]))
    out["room_cases"] = np.array(json.dumps(room_cases))
    out["obj_cases"] = np.array(json.dumps([(q, (fl if fl != [] else -1), rooms, k, neg) for (q, fl, rooms, k, neg) in obj_cases]))
    out["classes"] = np.array(json.dumps(classes))
    np.savez_compressed(os.path.join(HERE, "ref_retrieval.npz"), **out)
    print("saved ref_retrieval.npz", os.path.getsize(os.path.join(HERE, "ref_retrieval.npz")) // 1024, "KiB")
    for k in sorted(out):
        if k.startswith(("hmsg_room_", "room_", "obj_")) and not k.endswith("cases") and out[k].ndim == 1 and out[k].dtype == np.int64:
            print(k, out[k].tolist())


if __name__ == "__main__":
    main()
