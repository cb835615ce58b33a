"""Shared by the CPU (numpy test double) and GPU (libhmsg_b200.so) runs: the Graph mirror's room / object retrieval
must return exactly what the UNMODIFIED reference returned in this container (tests/golden/ref_retrieval.npz,
made by tests/golden/make_reference_golden_retrieval.py)."""
import json
import os
import types

import numpy as np

from holoagent_b200.memory.hmsg.graph.graph import Graph

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NS = types.SimpleNamespace


def check_retrieval_against_reference_run(engine):
    z = np.load(os.path.join(GOLD, "ref_retrieval.npz"))
    words = json.loads(str(z["words"])); tf = z["text_feats"]
    room_names = json.loads(str(z["room_names"])); name_feats = z["room_name_feats"]
    emb, room_of, views = z["obj_emb"], z["obj_room"], z["room_view_counts"]
    off = np.concatenate([[0], np.cumsum(views)])
    g = Graph({"pipeline": {}}, engine=engine, clip_feat_dim=emb.shape[1])
    g.objects = [NS(embedding=emb[i], object_id="0_%d_%d" % (room_of[i], i), room_id="0_%d" % room_of[i]) for i in range(len(emb))]
    g.rooms = [NS(room_id="0_%d" % r, name=room_names[r], embeddings=list(z["room_embs"][off[r]:off[r + 1]]),
                  objects=[o for o in g.objects if o.room_id == "0_%d" % r]) for r in range(len(views))]
    g.floors = [NS(floor_id="0", rooms=g.rooms[:3]), NS(floor_id="1", rooms=g.rooms[3:])]
    feat = lambda w: tf[words.index(w)][None]
    n_checked = 0
    for ci, (q, fl, method) in enumerate(json.loads(str(z["room_cases"]))):
        rl = g.rooms if fl == -1 else g.floors[fl].rooms
        nf = np.stack([name_feats[g.rooms.index(r)] for r in rl])
        assert g.query_hmsg_room(q, floor_id=fl, query_method=method, query_feats=feat(q), room_name_feats=nf) == z["hmsg_room_%d" % ci].tolist(), (ci, q, fl, method)
        assert g.query_room(q, floor_id=fl, query_method=method, query_feats=feat(q), room_name_feats=nf) == z["room_%d" % ci].tolist(), (ci, q, fl, method)
        n_checked += 2
    for ci, (q, fl, rooms, k, neg) in enumerate(json.loads(str(z["obj_cases"]))):
        names = neg if q in neg else [q] + neg
        qf = np.stack([tf[words.index(w)] for w in names])
        ids, rids = g.query_object(q, floor_id=fl, room_ids=rooms, top_k=k, negative_prompt=list(neg), query_feats=qf)
        assert ids == z["obj_%d_ids" % ci].tolist() and rids == z["obj_%d_rooms" % ci].tolist(), (ci, q, fl, rooms)
        n_checked += 1
    # ---- the robot's entry point (goal_pose_publisher.py:220 -> graph.py:3593 / :3484): instruction -> floor, rooms, objects
    table = {w: tf[i] for i, w in enumerate(words)}
    table.update({n: name_feats[i] for i, n in enumerate(room_names)})
    g.text_feats_fn = lambda texts: np.stack([table[t] for t in texts])
    for f, lvl in zip(g.floors, (0.0, 3.1)):
        f.floor_zero_level = lvl
    for h in json.loads(str(z["hier_cases"])):
        g.hier_query_parser = lambda ins, _p=tuple(h["parse"]): _p
        fl, rooms, objs, res = getattr(g, h["fn"])(h["instruction"], top_k=h["top_k"], use_gpt=False)
        assert (None if fl is None else fl.floor_id) == h["floor"], h
        assert [r.room_id for r in rooms] == h["rooms"] and [o.object_id for o in objs] == h["objects"], h
        assert res["negative_labels"] == h["negative_labels"]
        n_checked += 1
    # ---- multi-floor graph with floor_id = -1: rooms "0_1" / "1_1" collapse to one key (graph.py:3259-3272)
    mf_ids = json.loads(str(z["mf_room_ids"])); mo = np.concatenate([[0], np.cumsum(z["mf_view_counts"])])
    g2 = Graph({"pipeline": {}}, engine=engine, clip_feat_dim=emb.shape[1])
    g2.rooms = [NS(room_id=mf_ids[r], name="room", embeddings=list(z["mf_embs"][mo[r]:mo[r + 1]]), objects=[]) for r in range(len(mf_ids))]
    g2.floors = [NS(floor_id="0", rooms=g2.rooms[:2]), NS(floor_id="1", rooms=g2.rooms[2:])]
    for qi, q in enumerate(["kitchen", "office", "unknown area"]):
        assert g2.query_hmsg_room(q, floor_id=-1, query_method="view_embedding", query_feats=feat(q)) == z["mf_hmsg_room_%d" % qi].tolist(), q
        assert g2.query_room(q, floor_id=-1, query_method="view_embedding", query_feats=feat(q)) == z["mf_room_%d" % qi].tolist(), q
        n_checked += 2
    classes = json.loads(str(z["classes"]))
    label_feats = np.stack([tf[words.index(c)] for c in classes])
    assert [g.identify_object(emb[i], label_feats, classes) for i in range(0, len(emb), 7)] == json.loads(str(z["identify"]))
    return n_checked


def check_room_names_against_reference_run(engine):
    """generate_room_names("view_embedding" / "obj_embedding") and set_room_names must name the rooms exactly as the UNMODIFIED
    reference did (tests/golden/ref_roomnames.npz, made by tests/golden/make_reference_golden_roomnames.py)."""
    z = np.load(os.path.join(GOLD, "ref_roomnames.npz"))
    types_ = json.loads(str(z["types"])); tf = z["type_feats"]
    vc, oc = z["view_counts"], z["obj_counts"]
    voff = np.concatenate([[0], np.cumsum(vc)]); ooff = np.concatenate([[0], np.cumsum(oc)])
    verts = json.loads(str(z["vertices"]))
    g = Graph({"pipeline": {}}, engine=engine, clip_feat_dim=tf.shape[1])
    g.rooms = [NS(room_id="0_%d" % r, name="room %d" % r, embeddings=list(z["room_embs"][voff[r]:voff[r + 1]]), vertices=np.asarray(verts[r]),
                  objects=[NS(embedding=e) for e in z["obj_embs"][ooff[r]:ooff[r + 1]]]) for r in range(len(vc))]
    g.generate_room_names("view_embedding", types_, room_type_feats=tf)
    assert [r.name for r in g.rooms] == json.loads(str(z["names_view"]))      # incl. the room without views keeping its name
    for r in g.rooms:
        r.name = "room"
    g.generate_room_names("obj_embedding", types_, room_type_feats=tf)
    want = json.loads(str(z["names_obj"]))                                    # None: no objects (the reference raises inside sklearn there)
    assert [r.name for r in g.rooms] == [w if w is not None else "room" for w in want]
    g.set_room_names(["n%d" % i for i in range(len(g.rooms))])
    assert [r.name for r in g.rooms] == ["n%d" % i for i in range(len(g.rooms))]
    assert np.allclose(np.stack([r.room_center_pos for r in g.rooms]), z["centers"], rtol=0, atol=1e-12)
    import pytest
    with pytest.raises(NotImplementedError):
        g.generate_room_names("label", types_)
    with pytest.raises(NotImplementedError):
        g.build_hier_multimodal_scene_graph()
