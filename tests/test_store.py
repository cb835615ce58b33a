"""CPU: N4 loader - reads graph artefacts written in the reference's JSON schema."""
import json
import os

import numpy as np

from holoagent_b200.memory.hmsg.graph.store import load_feats_pt, load_graph_nodes


def _write_graph(root, d=128):
    rs = np.random.RandomState(0)
    for sub in ("objects", "rooms", "floors"):
        os.makedirs(os.path.join(root, sub))
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
    floors, rooms, objects = load_graph_nodes(str(tmp_path))
    assert [o.object_id for o in objects] == [f"0_{i % 2}_{i}" for i in range(7)]
    assert objects[5].embedding is None and objects[3].embedding.dtype == np.float64
    assert np.array_equal(objects[3].embedding, embs[3])
    assert [len(r.objects) for r in rooms] == [4, 3] and len(rooms[0].embeddings) == 3 and len(rooms[1].clip_embeddings) == 5
    assert floors[0].rooms == rooms


def test_feats_pt_roundtrip(tmp_path):
    import torch
    a = np.random.RandomState(1).randn(10, 16).astype(np.float32)
    torch.save(torch.from_numpy(a), os.path.join(tmp_path, "full_feats.pt"))
    assert np.array_equal(load_feats_pt(os.path.join(tmp_path, "full_feats.pt")), a)
