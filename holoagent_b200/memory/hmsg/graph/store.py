"""N4: read the reference's on-disk graph artefacts into retrieval-ready node lists.

Schema followed (JSON metadata; the .ply clouds are not needed for retrieval):
  objects/<object_id>.json  {"object_id","vertices","room_id","name","embedding","view_ids","best_view_id"}
                            (fsr_vln/memory/hmsg/graph/object.py:37-57, load_new :76-91)
  rooms/<room_id>.json      {"room_id","name","floor_id","objects","views","vertices","room_height",
                             "room_zero_level","embeddings","represent_images","sample_images","clip_embeddings"}
                            (room.py:309-333, load_new :354-374)
  floors/<floor_id>.json    {"floor_id","name","rooms","vertices","floor_height","floor_zero_level"} (floor.py:33-66)
  full_feats.pt / mask_feats.pt  torch.save of the node / mask feature arrays (graph.py:3797-3830)
"""
from __future__ import annotations

import glob
import json
import os

import numpy as np


class ObjectNode:
    def __init__(self, md):
        self.object_id = md["object_id"]
        self.room_id = md.get("room_id")
        self.name = md.get("name")
        emb = md.get("embedding", "")
        self.embedding = np.asarray(emb, dtype=np.float64) if emb != "" else None      # object.py:88 (float64 from JSON)
        self.vertices = np.asarray(md.get("vertices", []))
        self.view_ids = md.get("view_ids", [])
        self.best_view_id = md.get("best_view_id")


class RoomNode:
    def __init__(self, md):
        self.room_id = md["room_id"]
        self.name = md.get("name")
        self.floor_id = md.get("floor_id")
        self.object_ids = md.get("objects", [])
        self.objects = []
        self.views = md.get("views", [])
        self.vertices = np.asarray(md.get("vertices", []))
        self.embeddings = [np.asarray(e) for e in md.get("embeddings", [])]           # <= 24 representative view feats
        self.clip_embeddings = [np.asarray(e) for e in md.get("clip_embeddings", [])]  # all view feats
        self.sample_images = md.get("sample_images", [])
        self.represent_images = md.get("represent_images", [])


class FloorNode:
    def __init__(self, md):
        self.floor_id = md["floor_id"]
        self.name = md.get("name")
        self.room_ids = md.get("rooms", [])
        self.rooms = []


def _read_dir(path, cls, sort_key):
    out = []
    for fn in glob.glob(os.path.join(path, "*.json")):
        with open(fn) as f:
            out.append(cls(json.load(f)))
    out.sort(key=sort_key)
    return out


def _tail_int(x):
    try:
        return int(str(x).split("_")[-1])
    except ValueError:
        return str(x)


def load_graph_nodes(graph_path):
    """graph_<timestamp>/{floors,rooms,objects}/ -> (floors, rooms, objects) with the cross links the
    retrieval methods use (room.objects, floor.rooms)."""
    objects = _read_dir(os.path.join(graph_path, "objects"), ObjectNode, lambda o: _tail_int(o.object_id))
    rooms = _read_dir(os.path.join(graph_path, "rooms"), RoomNode, lambda r: _tail_int(r.room_id))
    floors = _read_dir(os.path.join(graph_path, "floors"), FloorNode, lambda f: _tail_int(f.floor_id)) if os.path.isdir(
        os.path.join(graph_path, "floors")) else []
    by_id = {o.object_id: o for o in objects}
    for r in rooms:
        r.objects = [by_id[i] for i in r.object_ids if i in by_id]
    rb = {r.room_id: r for r in rooms}
    for fl in floors:
        fl.rooms = [rb[i] for i in fl.room_ids if i in rb]
    return floors, rooms, objects


def load_feats_pt(path):
    """full_feats.pt / mask_feats.pt (graph.py:3820-3828) -> float32 ndarray"""
    import torch
    a = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(a, "numpy"):
        a = a.numpy()
    if isinstance(a, (list, tuple)):
        a = np.concatenate([np.asarray(x).reshape(-1, np.asarray(x).shape[-1]) for x in a])
    return np.ascontiguousarray(a, dtype=np.float32)
