"""CPU: N4 loader - reads graph artefacts written in the reference's JSON schema."""
import json
import os

import numpy as np

from holoagent_b200.memory.hmsg.graph.store import load_feats_pt, load_graph_nodes


def _write_graph(root, d=128):
    rs = np.random.RandomState(0)
    for sub in ("objects", "rooms", "floors", "views"):
        os.makedirs(os.path.join(root, sub))
    for v in ("0_0_3", "0_1_10", "0_1_2"):
        json.dump({"view_id": v, "room_id": v[:3], "img_id": 7, "object_ids": [1, 2], "img_path": "a.png", "text_discription": ["x"]},
                  open(os.path.join(root, "views", v + ".json"), "w"))
    embs = rs.randn(7, d)
    for i in range(7):
        md = {"object_id": f"0_{i % 2}_{i}", "vertices": rs.rand(8, 3).tolist(), "room_id": f"0_{i % 2}", "name": f"obj{i}",
              "embedding": embs[i].tolist() if i != 5 else "", "view_ids": [1, 2], "best_view_id": 1}
        json.dump(md, open(os.path.join(root, "objects", md["object_id"] + ".json"), "w"))
    for r in range(2):
        md = {"room_id": f"0_{r}", "name": f"room{r}", "floor_id": "0", "objects": [f"0_{r}_{i}" for i in range(7) if i % 2 == r], "views": [],
              "vertices": rs.rand(8, 3).tolist(), "room_height": 2.5, "room_zero_level": 0.0, "embeddings": rs.randn(3, d).tolist(),
              "represent_images": [], "sample_images": [100 * r + i for i in range(5)], "clip_embeddings": rs.randn(5, d).tolist()}
        json.dump(md, open(os.path.join(root, "rooms", md["room_id"] + ".json"), "w"))
    json.dump({"floor_id": "0", "name": "floor_0", "rooms": ["0_0", "0_1"], "vertices": [], "floor_height": 3.0, "floor_zero_level": 0.0},
              open(os.path.join(root, "floors", "0.json"), "w"))
    return embs


def test_load_graph_nodes(tmp_path):
    embs = _write_graph(str(tmp_path))
    floors, rooms, objects, views = load_graph_nodes(str(tmp_path))
    # the reference loader's order: sorted file names = full id strings (graph.py:1912-1930); the object without an embedding
    # is dropped from self.objects AND from its room's list (room-restricted queries index both)
    assert [o.object_id for o in objects] == sorted(f"0_{i % 2}_{i}" for i in range(7) if i != 5)
    by = {o.object_id: o for o in objects}
    assert by["0_1_3"].embedding.dtype == np.float64 and np.array_equal(by["0_1_3"].embedding, embs[3])
    assert [len(r.objects) for r in rooms] == [4, 2] and len(rooms[0].embeddings) == 3 and len(rooms[1].clip_embeddings) == 5
    assert all(o in objects for r in rooms for o in r.objects)
    assert floors[0].rooms == rooms and floors[0].floor_zero_level == 0.0
    assert [v.view_id for v in views] == ["0_0_3", "0_1_10", "0_1_2"] and views[1].room_id == "0_1" and views[0].object_ids == [1, 2]


def test_string_order_matches_reference_listing(tmp_path):
    """ "0_10" sorts before "0_2" exactly like sorted(os.listdir()) in graph.py:1912"""
    root = str(tmp_path)
    for sub in ("objects", "rooms", "floors"):
        os.makedirs(os.path.join(root, sub))
    json.dump({"floor_id": "0", "name": "f", "rooms": [], "vertices": [], "floor_height": 3.0, "floor_zero_level": 0.0}, open(os.path.join(root, "floors", "0.json"), "w"))
    for r in ("0_2", "0_10", "0_1"):
        json.dump({"room_id": r, "name": r, "floor_id": "0", "objects": [], "views": [], "vertices": [], "room_height": 1, "room_zero_level": 0,
                   "embeddings": [], "represent_images": [], "sample_images": [], "clip_embeddings": []}, open(os.path.join(root, "rooms", r + ".json"), "w"))
    _, rooms, _, _ = load_graph_nodes(root)
    assert [r.room_id for r in rooms] == ["0_1", "0_10", "0_2"]


def test_feats_pt_roundtrip(tmp_path):
    import torch
    a = np.random.RandomState(1).randn(10, 16).astype(np.float32)
    torch.save(torch.from_numpy(a), os.path.join(tmp_path, "full_feats.pt"))
    assert np.array_equal(load_feats_pt(os.path.join(tmp_path, "full_feats.pt")), a)
