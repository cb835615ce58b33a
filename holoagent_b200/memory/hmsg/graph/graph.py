"""Drop-in for the hot half of fsr_vln/memory/hmsg/graph/graph.py: ``Graph.create_feature_map``
(:262-491: node table, node features, 3-D mask merging into objects, per-object features) and the
retrieval cores ``query_hmsg_object`` (:3056-3162), ``query_object``
(:3363-3481), ``query_graph`` (:2189-2214), ``identify_object`` (:1441-1454),
``query_hmsg_room`` (:3164-3272), ``query_room`` (:3277-3359), the slow-path global view retrieval
(:2864-2897, ``query_views``) and the re-match inside the chosen view (:2977-2984,
``rematch_in_view``), all on libhmsg_b200.so.  The per-frame global embedding ``F_g`` the build computes
anyway is kept in ``self.frame_global_feats`` so that room building (graph.py:1119-1136, which
re-encodes every kept frame once per floor - SURVEY A9b) can reuse it instead of a third encoder pass.

Outside the hot path (and therefore injected by the caller instead of re-implemented):
SAM (``mask_generator.generate``), the CLIP text tower (``clip_model.text_encoder`` or
pre-computed ``query_feats``), floor/room segmentation, LLM/VLM reasoning, file I/O.  Method
names, argument names/defaults and return shapes follow the reference so that
application scripts and nav_agent's goal_pose_publisher call them unchanged.
"""
from __future__ import annotations

from typing import List

import numpy as np

from holoagent_b200.memory.hmsg.utils.clip_utils import get_text_feats_multiple_templates
from holoagent_b200.runtime import PointCloud, get_engine, to_o3d


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _ns(d):
    if isinstance(d, dict):
        return _Cfg({k: _ns(v) for k, v in d.items()})
    return d


class Graph:
    def __init__(self, cfg, dataset=None, clip_model=None, preprocess=None, mask_generator=None, engine=None, clip_feat_dim=512):
        """cfg: the reference's hydra config (dict / OmegaConf-like) - keys read here are the ones
        the reference reads on the hot path: pipeline.voxel_size,
        pipeline.skip_frames, pipeline.clip_bbox_margin, pipeline.clip_masked_weight,
        pipeline.max_mask_distance.  dataset: an RGBDDataset-like object yielding
        (rgb_image, depth_image, pose, rgb_intrinsics, depth_intrinsics) with .scale and
        .depth_intrinsics (generic.py:19-33)."""
        self.cfg = _ns(cfg) if isinstance(cfg, dict) else cfg
        self.dataset = dataset
        self.clip_model = clip_model
        self.preprocess = preprocess
        self.mask_generator = mask_generator
        self.clip_feat_dim = clip_feat_dim
        self.engine = engine or (clip_model.engine if clip_model is not None else get_engine(0))
        self.full_pcd = PointCloud()
        self.full_feats_array = None
        self.mask_feats = []
        self.mask_pcds = []
        self.frames_feats = []
        self.frames_pcd = []
        self.frame_global_feats = {}     # frame id -> F_g [d] (== get_img_feats(full frame), graph.py:1125-1129)
        self.objects = []
        self.rooms = []
        self.floors = []
        self._index_epoch = 0
        self.frame_batch = 16
        self.keep_frames_pcd = True      # per-frame 3-D masks as host point clouds (the reference's local `frames_pcd`)

    # ------------------------------------------------------------------ build (graph.py:262-415)
    def create_feature_map(self, save_path=None):
        if self.dataset is None:
            print("No dataset loaded")          # graph.py:267-269
            return
        import torch
        eng, p = self.engine, self.cfg.pipeline
        skip = int(p.skip_frames)
        ids = list(range(0, len(self.dataset), skip))
        # ---- pass 1 (graph.py:339-345): frames -> resident HBM store
        first = self.dataset[ids[0]]
        depth0 = np.array(first[1])
        H, W = depth0.shape
        eng.scene_begin(H, W, np.asarray(self.dataset.depth_intrinsics, dtype=np.float64), float(self.dataset.scale), float(p.voxel_size), len(ids))
        rgbs = []
        for i in ids:
            rgb_image, depth_image, pose, _, _ = self.dataset[i]
            rgb = np.array(rgb_image).astype(np.uint8)
            depth = np.array(depth_image).astype(np.uint16)
            if rgb.shape[:2] != depth.shape[:2]:
                import cv2
                rgb = cv2.resize(rgb, (depth.shape[1], depth.shape[0]), interpolation=cv2.INTER_AREA)   # generic.py:98-104
            eng.add_frames(depth[None], rgb[None], np.asarray(pose, dtype=np.float64).reshape(1, 16))
            rgbs.append(rgb)
        # ---- graph.py:348-358: voxel_down_sample, dbscan (identity), remove_radius_outlier
        eng.voxel_build()
        eng.radius_filter(1000, 1.0)
        xyz, rgbc, _, _ = eng.nodes_read()
        pc = PointCloud(xyz, rgbc)
        self.full_pcd = to_o3d(pc)
        frame_of = {i: k for k, i in enumerate(ids)}
        try:
            self.full_pcd._hmsg_engine = eng
            self.full_pcd._hmsg_frame_of = frame_of.__getitem__
        except AttributeError:          # open3d objects do not take attributes: keep the side table on self
            pass
        self._frame_of = frame_of
        # ---- pass 2 (graph.py:373-411)
        d = self.clip_feat_dim
        eng.features_begin(d)
        g = lambda k, dflt: getattr(p, k, dflt) if not isinstance(p, dict) else p.get(k, dflt)
        merge_type = g("merge_type", "sequential")
        if merge_type != "sequential":
            raise NotImplementedError("pipeline.merge_type=%r: only the reference's default 'sequential' merge is on the device" % (merge_type,))
        max_mask_distance = float(g("max_mask_distance", float("inf")))
        eng.objects_begin(float(g("init_overlap_thresh", 0.75)), float(p.voxel_size), float(g("iou_thresh", 0.05)))   # seq_merge args, graph.py:437-442
        self.frames_pcd, self.frames_feats, self.frame_global_feats = [], [], {}
        dev = f"cuda:{eng.device}"
        for b0 in range(0, len(ids), self.frame_batch):
            chunk = ids[b0:b0 + self.frame_batch]
            all_masks = [self.mask_generator.generate(rgbs[b0 + k]) for k in range(len(chunk))]   # SAM: outside the hot path
            M = max(len(m) for m in all_masks)
            if M == 0:
                continue
            n = len(chunk)
            seg = np.zeros((n, M, H, W), np.uint8)
            boxes = np.zeros((n, M, 4), np.int32)
            for k, ms in enumerate(all_masks):
                for j, m in enumerate(ms):
                    seg[k, j] = np.asarray(m["segmentation"]).astype(np.uint8)
                    boxes[k, j] = [int(v) for v in m["bbox"]]
                for j in range(len(ms), M):       # pad with an empty mask on a 1-pixel box (never wins a pixel)
                    boxes[k, j] = (0, 0, 1, 1)
            eng.masks_dense(b0, seg)
            feats = torch.empty((n * (2 * M + 1), d), dtype=torch.float32, device=dev)
            eng.encode_crops(b0, n, M, boxes, int(p.clip_bbox_margin), feats)     # crops + preprocess + encoder, fused
            Fp = eng.fuse_scatter(b0, n, M, feats.view(n, 2 * M + 1, d), float(p.clip_masked_weight),
                                  Fp_out=torch.empty((n, M, d), dtype=torch.float32, device=dev))
            eng.torch_wait()
            Fp = Fp.cpu()
            Fg = feats.view(n, 2 * M + 1, d)[:, 2 * M].cpu().numpy()
            for k, ms in enumerate(all_masks):
                self.frame_global_feats[chunk[k]] = Fg[k]
                self.frames_feats.append(Fp[k, :len(ms)])
                if self.keep_frames_pcd:
                    off, mx, mc, _ = eng.mask_nodes(b0 + k, float(p.voxel_size), M)
                    self.frames_pcd.append([to_o3d(PointCloud(mx[off[j]:off[j + 1]], mc[off[j]:off[j + 1]])) for j in range(len(ms))])
                # create_3d_masks + one seq_merge iteration (graph.py:391-402, :437-442), on the device
                eng.objects_add_frame(b0 + k, float(p.voxel_size), max_mask_distance)
        # ---- graph.py:413-415
        self.full_feats_array = eng.node_feats_finalize()
        # ---- graph.py:424-448: final merge + removal of masks with < 10 points -> self.mask_pcds
        eng.objects_finish(10)
        off, ox, oc = eng.objects_read()
        self.mask_pcds = [to_o3d(PointCloud(ox[off[j]:off[j + 1]], oc[off[j]:off[j + 1]])) for j in range(len(off) - 1)]
        # ---- graph.py:451-488: one feature per object (cosine-DBSCAN largest-cluster mean) -> self.mask_feats
        feats = eng.object_feats(self.full_feats_array, float(p.voxel_size), 0.8, 0.01, 100) if len(self.mask_pcds) else np.zeros((0, d), np.float32)
        self.mask_feats = [feats[j] for j in range(len(self.mask_pcds))]
        return self.full_feats_array

    # ------------------------------------------------------------------ N4: graphs built by the unmodified reference
    def load_hmsg_graph(self, path):
        """graph.py:1892-1987 (metadata only): floors / rooms / objects JSON -> node lists; the object
        embeddings are packed into one device matrix on the first query."""
        from holoagent_b200.memory.hmsg.graph.store import load_graph_nodes
        self.floors, self.rooms, self.objects = load_graph_nodes(path)
        self.objects = [o for o in self.objects if o.embedding is not None]
        self.invalidate_index()
        return self

    load_graph = load_hmsg_graph

    # ------------------------------------------------------------------ artefacts on disk (graph.py:3769-3990)
    def save_full_pcd(self, path):
        """graph.py:3769-3780: <path>/full_pcd.ply"""
        import os
        from holoagent_b200.runtime import write_point_cloud
        os.makedirs(path, exist_ok=True)
        write_point_cloud(os.path.join(path, "full_pcd.ply"), self.full_pcd)
        print("full pcd saved to disk in {}".format(path))
        return None

    def load_full_pcd(self, path):
        """graph.py:3782-3795"""
        import os
        from holoagent_b200.runtime import read_point_cloud
        if not os.path.exists(path):
            print("full pcd not found in {}".format(path))
            return None
        self.full_pcd = read_point_cloud(os.path.join(path, "full_pcd.ply"))
        print("full pcd loaded from disk with shape {}".format(np.asarray(self.full_pcd.points).shape))
        return self.full_pcd

    def save_full_pcd_feats(self, path):
        """graph.py:3797-3830: drops objects with empty clouds, then torch.save of the per-object features
        (mask_feats.pt) and of the node feature array (full_feats.pt)."""
        import os
        import torch
        os.makedirs(path, exist_ok=True)
        keep = [(pc, ft) for pc, ft in zip(self.mask_pcds, self.mask_feats) if len(pc.points) > 0]
        self.mask_pcds = [pc for pc, _ in keep]
        self.mask_feats = [ft for _, ft in keep]
        if len(self.mask_feats) != 0:
            self.mask_feats = np.array(self.mask_feats)
            torch.save(torch.from_numpy(self.mask_feats), os.path.join(path, "mask_feats.pt"))
        if self.full_feats_array is not None and len(self.full_feats_array) != 0:
            torch.save(torch.from_numpy(np.asarray(self.full_feats_array)), os.path.join(path, "full_feats.pt"))
        print("full pcd feats saved to disk in {}".format(path))
        return None

    def load_full_pcd_feats(self, path, full_feats=False, normalize=True):
        """graph.py:3832-3871: full_feats.pt -> self.full_feats_array (full_feats=True) or mask_feats.pt ->
        self.mask_feats, L2-normalised rows unless normalize=False."""
        import os
        import torch
        if not os.path.exists(path):
            print("full pcd feats not found in {}".format(path))
            return None
        name = "full_feats.pt" if full_feats else "mask_feats.pt"
        a = torch.load(os.path.join(path, name), map_location="cpu", weights_only=False)
        a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).float()
        if normalize:
            a = torch.nn.functional.normalize(a, p=2, dim=-1)
        a = a.cpu().numpy()
        print("full pcd feats loaded from disk with shape {}".format(a.shape))
        if full_feats:
            self.full_feats_array = a
        else:
            self.mask_feats = a
        return a

    def save_masked_pcds(self, path, state="both"):
        """graph.py:3880-3942: remove objects with < 10 points, write objects/pcd_<i>.ply and / or the union cloud
        masked_pcd.ply (every object painted a random colour, as the reference does - in place)."""
        import os
        from holoagent_b200.runtime import PointCloud, paint_uniform_color, write_point_cloud
        self.mask_feats = list(self.mask_feats)
        for i, pcd in reversed(list(enumerate(self.mask_pcds))):
            if len(pcd.points) < 10:
                self.mask_pcds.pop(i); self.mask_feats.pop(i)
        os.makedirs(path, exist_ok=True)
        objects_path = os.path.join(path, "objects")
        if state in ("both", "objects"):
            os.makedirs(objects_path, exist_ok=True)      # the reference's "objects" branch forgets to define this path
            for i, pcd in enumerate(self.mask_pcds):
                write_point_cloud(os.path.join(objects_path, "pcd_{}.ply".format(i)), pcd)
        if state in ("both", "full"):
            masked = PointCloud()
            for pcd in self.mask_pcds:
                paint_uniform_color(pcd, np.random.rand(3))
                masked += PointCloud(np.asarray(pcd.points), np.asarray(pcd.colors))
            write_point_cloud(os.path.join(path, "masked_pcd.ply"), masked)
        print("masked pcds saved to disk in {}".format(path))

    def load_masked_pcds_new(self, path):
        """graph.py:3944-3990: objects/pcd_<i>.ply -> self.mask_pcds; features of missing files are dropped."""
        import os
        from holoagent_b200.runtime import read_point_cloud
        if len(self.mask_feats) == 0:
            print("load full pcd feats first")
            return None
        objects_path = os.path.join(path, "objects")
        if not os.path.exists(objects_path):
            print("masked pcds for objects not found in {}".format(path))
            return None
        self.mask_pcds, not_found = [], []
        for i in range(len(os.listdir(objects_path))):
            fn = os.path.join(objects_path, "pcd_{}.ply".format(i))
            if os.path.exists(fn):
                self.mask_pcds.append(read_point_cloud(fn))
            else:
                print("masked pcd {} not found in {}".format(i, path))
                not_found.append(i)
        not_found = [i for i in not_found if i < len(self.mask_feats)]
        self.mask_feats = np.delete(np.asarray(self.mask_feats), not_found, axis=0)
        print("number of masked pcds loaded from disk {}".format(len(self.mask_pcds)))
        return self.mask_pcds

    # ------------------------------------------------------------------ retrieval plumbing
    def _text(self, queries: List[str], query_feats=None):
        if query_feats is not None:
            return np.asarray(query_feats, dtype=np.float32).reshape(len(queries), -1)
        return np.float32(get_text_feats_multiple_templates(queries, self.clip_model, self.clip_feat_dim))

    def _set_index(self, key, rows):
        """object_embs = np.array([obj.embedding ...]) (graph.py:3126) -> one HBM matrix, cached.  The matrix lives in the
        engine, so the tag of what it currently holds is kept ON the engine: another Graph (or a direct index_set) using
        the same engine invalidates it.  Call invalidate_index() after editing embeddings in place."""
        tag = (id(self), self._index_epoch, key)
        if getattr(self.engine, "_graph_index_tag", None) != tag:
            E = np.ascontiguousarray(np.asarray(rows, dtype=np.float32))
            d = E.shape[1]
            if d % 128:
                raise ValueError("embedding dimension must be a multiple of 128")
            self.engine.index_set(E)
            try:
                self.engine._graph_index_tag = tag
            except AttributeError:
                pass
        return self.engine

    def invalidate_index(self):
        """Forget the cached device matrix (objects / rooms were edited in place)."""
        self._index_epoch += 1

    # graph.py:1441-1454
    def identify_object(self, object_feat, text_feats, classes):
        eng = self._set_index(("labels", id(text_feats)), text_feats)
        sim = eng.query_scores(np.asarray(object_feat, dtype=np.float32).reshape(1, -1))
        return classes[int(np.argmax(sim))]

    # graph.py:2189-2214 (visualisation dropped)
    def query_graph(self, query, query_feats=None):
        q = self._text([query], query_feats)
        eng = self._set_index(("objects", len(self.objects)), [o.embedding for o in self.objects])
        ids, _ = eng.query_topk(q, min(5, len(self.objects)))
        return self.objects[int(ids[0][0])]

    # graph.py:3056-3162
    def query_hmsg_object(self, query: str, floor_id: int = -1, room_ids: List[int] = [], query_method: str = "clip", top_k: int = 1,
                          negative_prompt: List[str] = [], query_feats=None):
        if query in negative_prompt:
            query_id = negative_prompt.index(query)
        else:
            query_id = None
        if query_id is None:
            query = [query, *negative_prompt]
            query_id = 0
        else:
            query = negative_prompt
        q = self._text(query, query_feats)
        room_ids_list = []
        for obj in self.objects:
            for i, room in enumerate(self.rooms):
                if obj.room_id == room.room_id:
                    room_ids_list.append(i)
                    break
        objects_list = None
        if len(room_ids) != 0:
            objects_list, room_ids_list = [], []
            for i in room_ids:
                src = self.floors[floor_id].rooms[i].objects if floor_id != -1 else self.rooms[i].objects
                objects_list.extend(src)
                room_ids_list.extend([i] * len(src))
        if objects_list is None:
            # the reference leaves `objects_list` undefined here; searching all objects is
            # what its docstring promises ("Defaults to [], which means search from all rooms")
            objects_list = list(self.objects)
        if query_method != "clip":
            return NotImplementedError
        key = ("objsel", floor_id, tuple(room_ids), len(objects_list))
        eng = self._set_index(key, [o.embedding for o in objects_list])
        top_k_eff = min(top_k, len(objects_list))
        top_index = None
        if len(negative_prompt) > 0:
            ids, sc, nf = eng.query_object(q[None], query_id, top_k_eff)
            if nf[0] > 0:
                top_index, scores = ids[0][:nf[0]], sc[0][:nf[0]]
        if top_index is None:
            ids, sc = eng.query_topk(q[query_id:query_id + 1], top_k_eff)
            top_index, scores = ids[0], sc[0]
        target_object_id = [objects_list[i].object_id for i in top_index]
        target_object_score = [float(s) for s in scores]
        target_room_id = [room_ids_list[i] for i in top_index]
        target_id = [[i for i, x in enumerate(self.objects) if x.object_id == ti][0] for ti in target_object_id]
        return target_id, target_room_id, target_object_score

    # graph.py:3363-3481 (same core, returns (ids, room_ids))
    def query_object(self, query: str, floor_id: int = -1, room_ids: List[int] = [], query_method: str = "clip", top_k: int = 1,
                     negative_prompt: List[str] = [], query_feats=None):
        tid, trid, _ = self.query_hmsg_object(query, floor_id, room_ids, query_method, top_k, negative_prompt, query_feats)
        return tid, trid

    # graph.py:3164-3272
    def query_hmsg_room(self, query: str, floor_id: int = -1, query_method: str = "view_embedding", query_feats=None, room_name_feats=None):
        is_room_text_valid = query is not None and query != "" and "unknown" not in query.lower()
        q = self._text([query], query_feats)
        rooms_list = self.rooms if floor_id == -1 else self.floors[floor_id].rooms
        if query_method == "label" and is_room_text_valid:
            embs = room_name_feats if room_name_feats is not None else get_text_feats_multiple_templates(
                [r.name for r in rooms_list], self.clip_model, self.clip_feat_dim)
            eng = self._set_index(("roomnames", floor_id, len(rooms_list)), embs)
            sim = eng.query_scores(q)[0]
            top_index = np.lexsort((np.arange(len(sim)), -sim))
            tar = sim[top_index[0]]
            same = [int(top_index[0])] + [int(i) for i in top_index[1:] if abs(sim[i] - tar) < 1e-3]      # :3216-3221
            target_room_ids = [rooms_list[i].room_id for i in same]
            return [i for i, x in enumerate(rooms_list) if x.room_id in target_room_ids]
        rows, seg = [], []
        for ri, room in enumerate(rooms_list):
            e = np.stack(room.embeddings)
            rows.append(e); seg.extend([ri] * len(e))
        eng = self._set_index(("roomviews", floor_id, len(seg)), np.concatenate(rows))
        sim = eng.query_scores(q)[0]
        seg = np.asarray(seg)
        room_max = np.array([sim[seg == ri].max() for ri in range(len(rooms_list))])                      # :3250-3253
        order = sorted(range(len(rooms_list)), key=lambda r: room_max[r], reverse=True)
        out = [int(str(rooms_list[r].room_id).split("_")[-1]) for r in order]                             # :3262-3267
        return out[:min(len(out), 5 if is_room_text_valid else 10)]

    # graph.py:2864-2897 (slow path: the goal view over ALL rooms' view embeddings)
    def query_views(self, query, rooms_list=None, top_k: int = 24, query_feats=None):
        """-> (best_image_id, top_image_ids, top_scores): `sims = dot(query_feats[0], stack(clip_embeddings).T)`,
        `argmax`, `argsort(sims)[-top_k:][::-1]` with `top_k = min(24, len(sims))`; image ids come from
        `room.sample_images` (asserted to align with `room.clip_embeddings`, :2870-2871)."""
        rooms_list = self.rooms if rooms_list is None else rooms_list
        q = self._text([query], query_feats)
        ids, embs = [], []
        for room in rooms_list:
            assert len(room.sample_images) == len(room.clip_embeddings), \
                f"Number of images ({len(room.sample_images)}) != embeddings ({len(room.clip_embeddings)})"
            ids.extend(room.sample_images)
            embs.extend(room.clip_embeddings)
        if not embs:
            return None, [], []
        eng = self._set_index(("views", len(embs), id(rooms_list)), np.stack(embs))
        k = min(top_k, len(embs))
        top, sc = eng.query_topk(q[:1], k)
        return ids[int(top[0][0])], [ids[int(i)] for i in top[0]], [float(v) for v in sc[0]]

    # graph.py:2977-2984 (slow path: best object among those visible in the chosen view)
    def rematch_in_view(self, query, object_ids_in_view, query_feats=None):
        """-> (object_id, score) = argmax / max of dot(query_feats[0], embeddings of the view's objects)."""
        q = self._text([query], query_feats)
        by_id = {o.object_id: o for o in self.objects}
        objs = [by_id[i] for i in object_ids_in_view]
        if not objs:
            return None, None
        eng = self._set_index(("inview", tuple(object_ids_in_view)), [o.embedding for o in objs])
        top, sc = eng.query_topk(q[:1], 1)
        return objs[int(top[0][0])].object_id, float(sc[0][0])

    # graph.py:2238-2251 (clip branch; floors are few: the text embeddings of "floor i" are injected)
    def query_floor(self, query, floor_name_feats, query_feats=None, zero_level_order_ids=None):
        q = self._text([query], query_feats)
        eng = self._set_index(("floors", id(floor_name_feats)), floor_name_feats)
        top, _ = eng.query_topk(q[:1], 1)
        i = int(top[0][0])
        return zero_level_order_ids[i] if zero_level_order_ids is not None else i

    # graph.py:3277-3359 (view-embedding branch: per-room max, top 3)
    def query_room(self, query: str, floor_id: int = -1, query_method: str = "view_embedding", query_feats=None, room_name_feats=None):
        is_room_text_valid = query is not None and query != "" and "unknown" not in query.lower()
        if query_method == "label" and is_room_text_valid:      # same label branch as query_hmsg_room (:3300-3334)
            return self.query_hmsg_room(query, floor_id, "label", query_feats, room_name_feats)
        # view-embedding branch, also taken for an "unknown" room text: three highest-ranking rooms (:3345-3359)
        return self.query_hmsg_room("unknown", floor_id, "view_embedding", self._text([query], query_feats))[:3]
