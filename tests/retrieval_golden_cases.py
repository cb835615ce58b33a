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
    classes = json.loads(str(z["classes"]))
    label_feats = np.stack([tf[words.index(c)] for c in classes])
    assert [g.identify_object(emb[i], label_feats, classes) for i in range(0, len(emb), 7)] == json.loads(str(z["identify"]))
    return n_checked
