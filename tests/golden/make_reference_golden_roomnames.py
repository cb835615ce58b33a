"""Room-naming fixtures from the UNMODIFIED reference (fsr_vln, /root/reference) behind the import harness of
make_reference_golden.py.  What executes is the reference's own source:
  * Graph.generate_room_names            memory/hmsg/graph/graph.py:2146-2186
  * Room.infer_room_type_from_view_embedding   memory/hmsg/graph/room.py:131-168   (per-view arg-max + majority vote)
  * Room.infer_room_type_from_objects("obj_embedding")   room.py:241-306  -> feats_denoise_dbscan (graph_utils.py:682-728,
    real sklearn)
  * Graph.set_room_names                 graph.py:2129-2144
The CLIP text tower is replaced by a seeded lookup table keyed by string, as in the other generators.
Output: tests/golden/ref_roomnames.npz

    python tests/golden/make_reference_golden_roomnames.py        (container only; deterministic)
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
import memory.hmsg.graph.room as ref_room    # noqa: E402   (reference)

NS = types.SimpleNamespace
D = 256


def main():
    rs = np.random.RandomState(77)
    types_ = ["kitchen", "bedroom", "office", "corridor", "bathroom"]
    tf = rs.randn(len(types_), D).astype(np.float32); tf /= np.linalg.norm(tf, axis=1, keepdims=True)
    table = {w: tf[i] for i, w in enumerate(types_)}
    lookup = lambda q, m, dim: np.stack([table[w] for w in q])
    ref_graph.get_text_feats_multiple_templates = lookup
    ref_room.get_text_feats_multiple_templates = lookup

    n_rooms = 7
    view_counts = [5, 24, 1, 0, 9, 12, 6]           # room 3 has no view embeddings: its name must stay
    obj_counts = [6, 14, 3, 2, 0, 30, 8]            # room 4 has no objects
    rooms, room_embs, obj_embs = [], [], []
    for r in range(n_rooms):
        room = ref_room.Room("0_%d" % r, "0", name="room %d" % r)
        e = rs.randn(view_counts[r], D).astype(np.float32)
        if len(e):
            e /= np.linalg.norm(e, axis=1, keepdims=True)
            # views lean towards two types with nearly equal weight so that the vote, not one arg-max, decides
            lean = np.where(rs.rand(view_counts[r]) < 0.55, r % 5, (r + 2) % 5)
            e = (0.75 * e + 0.25 * tf[lean]).astype(np.float32)
        room.embeddings = list(e)
        oe = rs.randn(obj_counts[r], D).astype(np.float64)
        if len(oe):
            oe /= np.linalg.norm(oe, axis=1, keepdims=True)
            oe = 0.5 * oe + 0.5 * tf[(r + 1) % 5]
            if r == 5:                                # a tight cluster (cosine distance < 0.02) inside the room: DBSCAN keeps it
                oe[:12] = tf[4] + 0.01 * rs.randn(12, D)
        room.objects = [NS(embedding=oe[i], name="obj") for i in range(obj_counts[r])]
        room.vertices = rs.rand(4 + r, 3) * 5
        rooms.append(room); room_embs.append(e); obj_embs.append(oe)

    g = ref_graph.Graph.__new__(ref_graph.Graph)
    g.clip_model, g.clip_feat_dim = object(), D
    g.rooms = rooms
    out = {"types": np.array(json.dumps(types_)), "type_feats": tf, "view_counts": np.array(view_counts), "obj_counts": np.array(obj_counts),
           "room_embs": np.concatenate([e.reshape(-1, D) for e in room_embs]), "obj_embs": np.concatenate([e.reshape(-1, D) for e in obj_embs]),
           "vertices": np.array(json.dumps([r.vertices.tolist() for r in rooms]))}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        g.generate_room_names("view_embedding", types_)
        out["names_view"] = np.array(json.dumps([r.name for r in rooms]))
        for r in rooms:
            r.name = "room"
        # the reference raises inside sklearn for a room without objects (np.array([]) has no features): name those rooms apart
        names_obj = []
        for r in rooms:
            if len(r.objects) == 0:
                names_obj.append(None)
                continue
            r.infer_room_type_from_objects(infer_method="obj_embedding", default_room_types=types_, clip_model=g.clip_model, clip_feat_dim=D)
            names_obj.append(r.name)
        out["names_obj"] = np.array(json.dumps(names_obj))
        new_names = ["n%d" % i for i in range(n_rooms)]
        g.set_room_names(new_names)
        out["centers"] = np.stack([r.room_center_pos for r in rooms])
    np.savez_compressed(os.path.join(HERE, "ref_roomnames.npz"), **out)
    print("saved ref_roomnames.npz", os.path.getsize(os.path.join(HERE, "ref_roomnames.npz")) // 1024, "KiB")
    print("view:", json.loads(str(out["names_view"]))); print("obj :", names_obj)


if __name__ == "__main__":
    main()
