"""N4: read the reference's on-disk graph artefacts into retrieval-ready node lists.

Schema followed (JSON metadata; the .ply clouds are not needed for retrieval):
  objects/<object_id>.json  {"object_id","vertices","room_id","name","embedding","view_ids","best_view_id"}
                            (fsr_vln/memory/hmsg/graph/object.py:37-57, load_new :76-91)
  rooms/<room_id>.json      {"room_id","name","floor_id","objects","views","vertices","room_height",
                             "room_zero_level","embeddings","represent_images","sample_images","clip_embeddings"}
                            (room.py:309-333, load_new :354-374)
  floors/<floor_id>.json    {"floor_id","name","rooms","vertices","floor_height","floor_zero_level"} (floor.py:33-66)
  views/<view_id>.json      {"view_id","room_id","img_id","object_ids","img_path","text_discription"} (view.py:62-72, :95-108)
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
        self.floor_height = md.get("floor_height")
        self.floor_zero_level = md.get("floor_zero_level", 0.0)


class ViewNode:
    """views/<view_id>.json (view.py:62-72, load :95-108): the image a room keeps plus the objects visible in it"""

    def __init__(self, md, view_id=None):
        self.view_id = md.get("view_id", view_id)
        self.room_id = md.get("room_id")
        self.img_id = md.get("img_id")
        self.img_path = md.get("img_path")
        self.object_ids = md.get("object_ids", [])
        self.text_discription = md.get("text_discription", [])
        self.embedding = None


def _read_dir(path, cls):
    """Node files in the reference loader's order: `sorted(os.listdir())` of the file names (graph.py:1897-1930), i.e. by
    the full id STRING ("0_10" before "0_2").  The reference enumerates the .ply clouds and reads the .json next to
    each; graphs exported without clouds are enumerated by their .json files."""
    if not os.path.isdir(path):
        return []
    names = sorted(os.listdir(path))
    stems = [n[:-4] for n in names if n.endswith(".ply")]
    if not stems:
        stems = [n[:-5] for n in names if n.endswith(".json")]
    out = []
    for st in stems:
        fn = os.path.join(path, st + ".json")
        if not os.path.exists(fn):
            continue
        with open(fn) as f:
            out.append((st, cls(json.load(f))))
    return out


def load_graph_nodes(graph_path):
    """graph_<timestamp>/{floors,rooms,objects,views}/ -> (floors, rooms, objects, views) with the cross links the
    retrieval methods use, built the way graph.py:1892-1987 builds them: an object's room is the "<floor>_<room>" prefix of
    its file name (:1936, :1950), `room.objects` / `floor.rooms` are appended in load order."""
    floors = [n for _, n in _read_dir(os.path.join(graph_path, "floors"), FloorNode)]
    rooms = []
    for st, r in _read_dir(os.path.join(graph_path, "rooms"), RoomNode):
        r.room_id = st
        rooms.append(r)
    rb = {r.room_id: r for r in rooms}
    for fl in floors:
        fl.rooms = []
    for r in rooms:                                             # :1924-1926: floors[int(room.floor_id)].rooms.append(room)
        try:
            fl = floors[int(r.floor_id)]
        except (ValueError, IndexError, TypeError):
            fl = next((f for f in floors if str(f.floor_id) == str(r.floor_id)), None)
        if fl is not None:
            fl.rooms.append(r)
    objects = []
    for st, o in _read_dir(os.path.join(graph_path, "objects"), ObjectNode):
        o.object_id = st
        o.room_id = "_".join(st.split("_")[:2])
        if o.embedding is None:                                 # an object without an embedding cannot be ranked; it is dropped
            continue                                            # from BOTH lists so that room-restricted queries stay aligned
        objects.append(o)
        if o.room_id in rb:
            rb[o.room_id].objects.append(o)                     # :1959 parent_room.add_object(objectt)
    views = []
    for st, v in _read_dir(os.path.join(graph_path, "views"), ViewNode):
        v.view_id = st
        v.room_id = "_".join(st.split("_")[:2])
        views.append(v)
    return floors, rooms, objects, views


def load_feats_pt(path):
    """full_feats.pt / mask_feats.pt (graph.py:3820-3828) -> float32 ndarray"""
    import torch
    a = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(a, "numpy"):
        a = a.numpy()
    if isinstance(a, (list, tuple)):
        a = np.concatenate([np.asarray(x).reshape(-1, np.asarray(x).shape[-1]) for x in a])
    return np.ascontiguousarray(a, dtype=np.float32)
