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


def hierarchical_merge(eng, frames, th, th_factor, down_size, proxy_th):
    """graph_utils.py:958-1012 on the device merge kernels: `frames` = per frame (offsets, xyz, rgb) ragged host lists.
    Adjacent lists are merged level by level with a decreasing threshold (each pair = hmsg_objects_begin + two
    hmsg_objects_add_masks: the first list is taken as is, adding the second merges `A + B`).  The last list is left in
    the engine under threshold 0.75: the caller's hmsg_objects_finish applies the closing merge_3d_masks(.., 0.75) of
    :1006-1011 and the `< 10 points` removal of graph.py:444-448."""
    lists = [tuple(f) for f in frames]
    while len(lists) > 1:
        nxt = []
        for i in range(0, len(lists), 2):
            if i == len(lists) - 1:
                nxt.append(lists[i])
                break
            eng.objects_begin(th, down_size, proxy_th)
            eng.objects_add_masks(*lists[i])
            eng.objects_add_masks(*lists[i + 1])
            nxt.append(eng.objects_read())
        lists = nxt
        if len(lists) > 1:
            th -= th_factor * (len(lists) - 2) / max(1, len(lists) - 1)
    eng.objects_begin(0.75, down_size, proxy_th)
    if lists:
        eng.objects_add_masks(*lists[0])


def _feats_denoise_dbscan(feats, eps=0.02, min_points=2):
    """graph_utils.py:682-728 with the defaults room.py:299 uses: sklearn cosine DBSCAN, mean of the largest cluster, mean of
    all rows when every row is noise (host side, a handful of rows per room)"""
    from collections import Counter
    from sklearn.cluster import DBSCAN
    feats = np.asarray(feats)
    labels = DBSCAN(eps=eps, min_samples=min_points, metric="cosine").fit(feats).labels_
    counter = Counter(labels)
    counter.pop(-1, None)
    if not counter:
        return np.mean(feats, axis=0)
    label, _ = counter.most_common(1)[0]
    kept = feats[labels == label]
    return np.mean(kept, axis=0) if len(kept) > 1 else kept


class _LazyFramesPcd:
    """The reference's local `frames_pcd` (graph.py:371, :401): per frame the list of 3-D mask clouds.  They live in
    the engine's HBM mask store; a frame is copied to the host only when somebody indexes it."""

    def __init__(self, eng, frames):
        self._eng, self._frames = eng, list(frames)

    def __len__(self):
        return len(self._frames)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        off, xyz, rgb, _ = self._eng.mask_store_read(self._frames[i])
        return [to_o3d(PointCloud(xyz[off[j]:off[j + 1]], rgb[off[j]:off[j + 1]])) for j in range(len(off) - 1)]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class B200HotPath:
    """The hot half of the reference's Graph on libhmsg_b200.so, written as a MIXIN: every method has the reference's
    name and signature, so `class Graph(B200HotPath, <reference Graph>)` (see `dropin_graph_class`) lets the reference's
    own glue - query_hierarchy_protected (graph.py:3593, what nav_agent's goal_pose_publisher.py:220 calls), room / floor
    building, LLM reasoning - reach the B200 cores through normal method resolution.  `Graph` below is the same mixin
    stand-alone (no reference checkout needed)."""

    def _b200_init(self, engine=None, clip_feat_dim=None):
        if clip_feat_dim is not None:
            self.clip_feat_dim = clip_feat_dim
        clip_model = getattr(self, "clip_model", None)
        self.engine = engine or (getattr(clip_model, "engine", None) or get_engine(0))
        self.frame_global_feats = {}     # frame id -> F_g [d] (== get_img_feats(full frame), graph.py:1125-1129)
        self._index_epoch = 0
        self.frame_batch = 32
        self.merge_objects = True        # graph.py:424-488 (N1 + N2) after the ingest
        self.exact_mask_sums = False     # True: per-frame create_3d_masks with Open3D's ordered float64 sums (bit-identical, slower)
        self.views = getattr(self, "views", [])

    def _ensure_encoder(self):
        """Drop-in mode: the reference constructor built `self.clip_model` with open_clip (graph.py:98-119).  Its visual
        tower is copied into the engine once; the text tower stays where it is."""
        eng = self.engine
        if getattr(eng, "vit", None) is not None:
            return
        cm = getattr(self, "clip_model", None)
        v = getattr(cm, "visual", None)
        if v is None or not hasattr(v, "state_dict"):
            raise RuntimeError("no encoder loaded: pass a B200ClipModel (or an open_clip model with a .visual tower) as clip_model")
        from holoagent_b200.memory.hmsg.utils.clip_utils import visual_tower_shape
        sd, shape = visual_tower_shape(v)
        eng.encoder_load(sd, **shape)

    @staticmethod
    def _resize_for_points(rgb, depth_shape):
        """generic.py:98-104: create_pcd brings the colour image to the depth size with cv2.resize(..., INTER_AREA)"""
        import cv2
        return cv2.resize(rgb, (int(depth_shape[1]), int(depth_shape[0])), interpolation=cv2.INTER_AREA)

    # ------------------------------------------------------------------ build (graph.py:262-415)
    def create_feature_map(self, save_path=None):
        if self.dataset is None:
            print("No dataset loaded")          # graph.py:267-269
            return
        import torch
        self._ensure_encoder()
        eng, p = self.engine, self.cfg.pipeline
        g = lambda k, dflt: getattr(p, k, dflt) if not isinstance(p, dict) else p.get(k, dflt)
        skip = int(p.skip_frames)
        ids = list(range(0, len(self.dataset), skip))
        nF, FB = len(ids), int(self.frame_batch)
        # ---- pass 1 (graph.py:339-345): frames -> resident HBM store, uploaded batch-wise on the copy stream
        first = self.dataset[ids[0]]
        H, W = np.array(first[1]).shape
        eng.scene_begin(H, W, np.asarray(self.dataset.depth_intrinsics, dtype=np.float64), float(self.dataset.scale), float(p.voxel_size), nF)
        rgbs = []
        resized = False                          # rgb and depth sizes differ somewhere: two resized colour images per frame
        for b0 in range(0, nF, FB):
            chunk = ids[b0:b0 + FB]
            dd = np.empty((len(chunk), H, W), np.uint16); cc = np.empty((len(chunk), H, W, 3), np.uint8); pp = np.empty((len(chunk), 16), np.float64)
            for k, i in enumerate(chunk):
                rgb_image, depth_image, pose, _, _ = self.dataset[i]
                depth = np.array(depth_image).astype(np.uint16)
                rgb = np.array(rgb_image).astype(np.uint8)
                cc[k] = rgb if rgb.shape[:2] == depth.shape[:2] else self._resize_for_points(rgb, depth.shape)
                if rgb.shape[:2] != depth.shape[:2]:
                    # graph.py:378-379: the feature pass (SAM + crops) sees PIL `rgb_image.resize(depth_image.size)` (bicubic);
                    # the point colours of create_pcd come from a cv2 INTER_AREA resize (generic.py:98-104).  Geometry is built
                    # from the second, then hmsg_scene_put_rgb swaps in the first for the crops.
                    from PIL import Image
                    rgb = np.asarray(Image.fromarray(rgb).resize((depth.shape[1], depth.shape[0])))
                    resized = True
                dd[k], pp[k] = depth, np.asarray(pose, dtype=np.float64).reshape(16)
                rgbs.append(rgb)
            eng.put_frames_host(b0, dd, cc, pp)
            # datasets whose create_pcd reads the intrinsics per frame (dataloader/iphone.py:325: frames[image_id - 1]["K"])
            ds = self.dataset
            if hasattr(ds, "frame_K"):
                eng.set_intrinsics(b0, np.stack([np.asarray(ds.frame_K(i), dtype=np.float64) for i in chunk]))
            elif hasattr(ds, "frames") and hasattr(ds, "indices"):
                eng.set_intrinsics(b0, np.stack([np.asarray(ds.frames[ds.indices[i] - 1]["K"], dtype=np.float64) for i in chunk]))
        eng.set_num_frames(nF)
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
        main = self.cfg.get("main") if isinstance(self.cfg, dict) else getattr(self.cfg, "main", None)
        sp = (main.get("save_path") if isinstance(main, dict) else getattr(main, "save_path", None)) if main is not None else None
        if sp:
            self.save_full_pcd(path=sp)         # graph.py:359
        # ---- pass 2 (graph.py:373-411)
        if resized:
            for b0 in range(0, nF, FB):
                eng.put_rgb_host(b0, np.stack(rgbs[b0:b0 + FB]))
        d = self.clip_feat_dim
        eng.features_begin(d)
        merge_type = g("merge_type", "sequential")
        if merge_type not in ("sequential", "hierarchical"):
            raise ValueError("pipeline.merge_type=%r (graph.py:425-443 knows 'hierarchical' and 'sequential')" % (merge_type,))
        max_mask_distance = float(g("max_mask_distance", float("inf")))
        vs = float(p.voxel_size)
        exact = bool(self.exact_mask_sums) and merge_type == "sequential" and self.merge_objects
        if exact:
            eng.objects_begin(float(g("init_overlap_thresh", 0.75)), vs, float(g("iou_thresh", 0.05)))   # seq_merge args, graph.py:437-442
        eng.mask_store_reset()
        self.frames_feats, self.frame_global_feats = [], {}
        dev = f"cuda:{eng.device}"
        bufs, events = [None, None], [None, None]
        kept = []                                 # (b0, n, M, counts, Fp device, Fg device)
        for bi, b0 in enumerate(range(0, nF, FB)):
            chunk = ids[b0:b0 + FB]
            n = len(chunk)
            all_masks = [self.mask_generator.generate(rgbs[b0 + k]) for k in range(n)]   # SAM: outside the hot path
            counts = np.array([len(m) for m in all_masks], np.int32)
            M = int(counts.max()) if n else 0
            if M == 0:
                kept.append((b0, n, 0, counts, None, None))
                continue
            # dense masks -> pinned staging (double buffered: the GPU works on batch i while the host fills batch i+1)
            slot = bi & 1
            need = n * M * H * W
            if bufs[slot] is None or bufs[slot].numel() < need:
                bufs[slot] = torch.empty(need, dtype=torch.uint8).pin_memory()
            if events[slot] is not None:
                events[slot].synchronize()
            seg = bufs[slot][:need].view(n, M, H, W)
            segn = seg.numpy()
            boxes = np.zeros((n, M, 4), np.int32)
            boxes[:, :, 2:] = 1                     # padded slots: an empty mask on a 1-pixel box
            for k, ms in enumerate(all_masks):
                for j, m in enumerate(ms):
                    segn[k, j] = m["segmentation"]
                    boxes[k, j] = m["bbox"]
                if len(ms) < M:
                    segn[k, len(ms):] = 0
            eng.masks_dense(b0, seg)
            events[slot] = torch.cuda.Event()
            events[slot].record(eng.torch_stream())
            eng.masks_counts(b0, counts)            # ragged SAM output: the softmax of extractor.py:168-172 runs over the frame's own masks
            feats = torch.empty((n * (2 * M + 1), d), dtype=torch.float32, device=dev)
            eng.encode_crops(b0, n, M, boxes, int(p.clip_bbox_margin), feats)     # crops + preprocess + encoder, fused
            Fp = eng.fuse_scatter(b0, n, M, feats.view(n, 2 * M + 1, d), float(p.clip_masked_weight),
                                  Fp_out=torch.empty((n, M, d), dtype=torch.float32, device=dev))
            kept.append((b0, n, M, counts, Fp, feats.view(n, 2 * M + 1, d)[:, 2 * M].clone()))
            # create_3d_masks for every frame of the batch (graph.py:391-402) -> HBM mask store
            eng.mask_nodes_batch(b0, n, vs, max_mask_distance, keep=True)
            if exact:
                for k in range(n):                # + one seq_merge iteration per frame with ordered sums (graph.py:437-442)
                    eng.objects_add_frame(b0 + k, vs, max_mask_distance)
        # one pass of device -> host copies after the loop (no per-batch sync)
        eng.torch_wait()
        for (b0, n, M, counts, Fp, Fg) in kept:
            Fp_h = Fp.cpu() if Fp is not None else None
            Fg_h = Fg.cpu().numpy() if Fg is not None else None
            for k in range(n):
                self.frames_feats.append(Fp_h[k, :counts[k]] if Fp_h is not None else torch.zeros((0, d)))
                if Fg_h is not None:
                    self.frame_global_feats[ids[b0 + k]] = Fg_h[k]
        self.frames_pcd = _LazyFramesPcd(eng, range(nF))
        # ---- graph.py:413-415
        self.full_feats_array = eng.node_feats_finalize()
        if not self.merge_objects:
            return self.full_feats_array
        # ---- graph.py:424-448: merge the 3-D masks into object instances, drop those with < 10 points -> self.mask_pcds
        if merge_type == "hierarchical":
            self._hierarchical_merge(nF, float(g("init_overlap_thresh", 0.75)), float(g("overlap_thresh_factor", 0.025)), vs, float(g("iou_thresh", 0.05)))
        else:
            if not exact:
                eng.objects_begin(float(g("init_overlap_thresh", 0.75)), vs, float(g("iou_thresh", 0.05)))
                eng.objects_merge_stored(0, nF)
            eng.objects_finish(10)
        off, ox, oc = eng.objects_read()
        self.mask_pcds = [to_o3d(PointCloud(ox[off[j]:off[j + 1]], oc[off[j]:off[j + 1]])) for j in range(len(off) - 1)]
        # ---- graph.py:451-488: one feature per object (cosine-DBSCAN largest-cluster mean) -> self.mask_feats
        feats = eng.object_feats(self.full_feats_array, vs, 0.8, 0.01, 100) if len(self.mask_pcds) else np.zeros((0, d), np.float32)
        self.mask_feats = [feats[j] for j in range(len(self.mask_pcds))]
        return self.full_feats_array

    def _hierarchical_merge(self, nF, th, th_factor, down_size, proxy_th):
        """graph.py:425-433 on the frames of the mask store"""
        hierarchical_merge(self.engine, [self.engine.mask_store_read(f)[:3] for f in range(nF)], th, th_factor, down_size, proxy_th)
        self.engine.objects_finish(10)

    # ------------------------------------------------------------------ retrieval plumbing
    def _text(self, queries: List[str], query_feats=None):
        if query_feats is not None:
            return np.asarray(query_feats, dtype=np.float32).reshape(len(queries), -1)
        fn = getattr(self, "text_feats_fn", None)       # dropin_graph_class: the reference's own get_text_feats_multiple_templates
        if fn is not None:
            return np.float32(fn(list(queries)))
        return np.float32(get_text_feats_multiple_templates(queries, self.clip_model, self.clip_feat_dim))

    def _set_index(self, key, rows):
        """object_embs = np.array([obj.embedding ...]) (graph.py:3126) -> one HBM matrix, cached.  The matrix lives in the
        engine, so the tag of what it currently holds is kept ON the engine: another Graph (or a direct index_set) using
        the same engine invalidates it.  Call invalidate_index() after editing embeddings in place."""
        E = np.ascontiguousarray(np.asarray(rows, dtype=np.float32))
        # the tag carries a digest of the matrix itself: equal lengths / recycled object ids can never alias a stale
        # device copy (hashing a few MB is far cheaper than the upload it saves)
        import hashlib
        tag = (id(self), self._index_epoch, key, E.shape, hashlib.blake2b(E.tobytes(), digest_size=16).digest())
        if getattr(self.engine, "_graph_index_tag", None) != tag:
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
        eng = self._set_index(("labels",), text_feats)
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
            embs = room_name_feats if room_name_feats is not None else self._text([r.name for r in rooms_list])     # :3198-3203
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
        # :3259-3272: the reference builds a dict keyed by the trailing integer of the room id, so rooms "0_1" and "1_1"
        # (floor_id = -1 on a multi-floor graph) collapse to one key that keeps its first position
        out = list(dict.fromkeys(int(str(rooms_list[r].room_id).split("_")[-1]) for r in order))
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
        eng = self._set_index(("views", len(embs)), np.stack(embs))
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

    # graph.py:2216-2258
    def query_floor(self, query, query_method="clip", floor_name_feats=None, query_feats=None, zero_level_order_ids=None):
        """A number in text form selects by floor level order (:2231-2237); otherwise the clip branch matches the query
        against the text embeddings of "floor i" (:2239-2252; pass `floor_name_feats` / `query_feats` when the text tower
        is not attached).  Second positional argument may also be the floor-name embedding matrix (round-1 signature)."""
        if not isinstance(query_method, str):
            floor_name_feats, query_method = query_method, "clip"
        n_names = len(self.floors) if floor_name_feats is None else len(floor_name_feats)
        if zero_level_order_ids is None and self.floors and n_names == len(self.floors) and all(hasattr(f, "floor_zero_level") for f in self.floors):
            zero_level_order_ids = np.argsort([f.floor_zero_level for f in self.floors])           # :2231-2233
        try:
            i = int(query) - 1
            return zero_level_order_ids[i] if zero_level_order_ids is not None else i
        except BaseException:
            pass
        if query_method != "clip":
            raise NotImplementedError("query_floor: only the 'clip' branch runs on the device (the 'gpt' branch is an LLM call)")
        q = self._text([query], query_feats)
        if floor_name_feats is None:
            floor_name_feats = self._text(["floor " + str(i) for i in range(len(self.floors))])
        eng = self._set_index(("floors",), floor_name_feats)
        sim = eng.query_scores(q[:1])[0]
        i = int(np.argsort(sim)[::-1][0])             # :2251 (ties -> the higher index, like the reference's argsort()[::-1])
        return zero_level_order_ids[i] if zero_level_order_ids is not None else i

    # graph.py:3277-3359 (view-embedding branch: per-room max, top 3)
    def query_room(self, query: str, floor_id: int = -1, query_method: str = "view_embedding", query_feats=None, room_name_feats=None):
        is_room_text_valid = query is not None and query != "" and "unknown" not in query.lower()
        if query_method == "label" and is_room_text_valid:      # same label branch as query_hmsg_room (:3300-3334)
            return self.query_hmsg_room(query, floor_id, "label", query_feats, room_name_feats)
        # view-embedding branch, also taken for an "unknown" room text: three highest-ranking rooms (:3345-3359)
        return self.query_hmsg_room("unknown", floor_id, "view_embedding", self._text([query], query_feats))[:3]


class B200Standalone:
    """What the stand-alone `Graph` adds on top of the mixin: the artefact I/O (graph.py:1892-1987, :3769-3990) and the
    fast-path glue of query_hierarchy_protected (graph.py:3484-3716) with an injectable instruction parser.  In drop-in
    mode (`dropin_graph_class`) these stay the reference's own methods."""

    # ------------------------------------------------------------------ N4: graphs built by the unmodified reference
    def load_hmsg_graph(self, path):
        """graph.py:1892-1987 (metadata only): floors / rooms / objects JSON -> node lists; the object
        embeddings are packed into one device matrix on the first query."""
        from holoagent_b200.memory.hmsg.graph.store import load_graph_nodes
        self.floors, self.rooms, self.objects, self.views = load_graph_nodes(path)
        self.graph_path = path
        self.invalidate_index()
        return self

    load_graph = load_hmsg_graph

    # ------------------------------------------------------------------ room names (graph.py:2129-2186; Appendix A: room.py:160-168, :303-306)
    def set_room_names(self, room_names: List[str]):
        """graph.py:2129-2144: one name per room, room_center_pos = mean of the room's vertices"""
        assert len(room_names) == len(self.rooms), "The length of room_names should be the same as the number of rooms in the graph"
        for room, name in zip(self.rooms, room_names):
            room.name = name
            v = np.asarray(getattr(room, "vertices", []), dtype=np.float64)
            if v.size:
                room.room_center_pos = np.mean(v.reshape(-1, v.shape[-1]), axis=0)

    def generate_room_names(self, generate_method: str = "label", default_room_types: List[str] = None, room_type_feats=None):
        """graph.py:2146-2186.  "view_embedding": every stored view embedding of a room votes for its best room type, the
        type with most votes names the room (room.py:148-168; ties -> the lower type index, as np.unique + argmax give);
        "obj_embedding": cosine-DBSCAN(eps 0.02, min 2) mean of the room's object embeddings against the type texts
        (room.py:286-303, graph_utils.py:682-728).  The dot products run on the device like every other retrieval site;
        `room_type_feats` [len(default_room_types), d] replaces the text tower.  "label" asks an LLM about the object
        names (llm_utils.infer_room_type_from_object_list_chat): outside the hot path - use the reference Graph through
        `dropin_graph_class` for it."""
        if generate_method == "label":
            raise NotImplementedError('generate_room_names("label") calls an LLM (room.py:262-283); use dropin_graph_class(<reference Graph>)')
        if generate_method not in ("obj_embedding", "view_embedding"):
            return NotImplementedError                        # sic: the reference returns the class (graph.py:2185)
        assert default_room_types is not None, "You should provide a list of default room types"
        text = self._text(list(default_room_types), room_type_feats)
        for room in self.rooms:
            if generate_method == "view_embedding":
                if len(room.embeddings) == 0:
                    print("empty embeddings")                  # room.py:143-145: the name stays as it was
                    continue
                eng = self._set_index(("roomtypes", len(default_room_types)), text)
                sim = eng.query_scores(np.asarray(np.stack(room.embeddings), dtype=np.float32))      # [views, types]
                col_ids = np.argmax(sim, axis=1)
                unique, counts = np.unique(col_ids, return_counts=True)
                room.name = default_room_types[int(unique[int(np.argmax(counts))])]
            else:
                embs = [np.asarray(o.embedding, dtype=np.float64) for o in room.objects]
                if not embs:
                    continue
                rep = _feats_denoise_dbscan(np.stack(embs)).reshape(1, -1)
                eng = self._set_index(("roomtypes", len(default_room_types)), text)
                room.name = default_room_types[int(np.argmax(eng.query_scores(np.asarray(rep, dtype=np.float32))))]

    def build_hier_multimodal_scene_graph(self, save_path=None):
        """graph.py:2033-2127: floor / room segmentation, room views, navigation graph - once-per-scene CPU heuristics and
        model calls outside SURVEY §8's path.  In drop-in mode this is the reference's own method."""
        raise NotImplementedError("floors / rooms / navigation graph construction is outside the B200 hot path: build the class with "
                                  "dropin_graph_class(<reference Graph>) to keep the reference's build_hier_multimodal_scene_graph")

    build_graph = build_hier_multimodal_scene_graph

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

    # graph.py:3593-3716 / :3484-3591 (fast path; the slow `use_gpt` branch is LLM / VLM reasoning, out of scope)
    BACKGROUND_LABELS = ["background", "divider", "ledge", "pillar", "tape", "stairs", "door", "doors", "stair", "window", "glass", "railing",
                         "glass doors", "whiteboard", "sliding door", "carpet", "ceiling", "curtain"]

    def _parse_hier(self, query_instruction, icra=False):
        """The reference parses "object X in room Y on floor Z" with an LLM prompt (llm_utils); a callable
        `self.hier_query_parser(instruction) -> (floor_query, room_query, object_query)` is injected here."""
        parser = getattr(self, "hier_query_parser", None)
        if parser is None:
            raise RuntimeError("query_hierarchy_protected: set graph.hier_query_parser (the LLM parse of the instruction is outside the hot path)")
        return parser(query_instruction)

    def _hier_result(self, floor_id, room_ids, object_ids, res_dict):
        return (self.floors[floor_id] if floor_id != -1 else None,
                [self.floors[floor_id].rooms[k] for k in room_ids] if floor_id != -1 else [self.rooms[k] for k in room_ids],
                [self.objects[i] for i in object_ids], res_dict)

    def query_hierarchy_protected(self, query_instruction: str, top_k: int = 1, use_gpt: bool = False):
        """-> (Floor | None, [Room], [Object], res_dict): floor by number / clip, room by label match, objects by
        embedding kNN with the reference's negative labels (graph.py:3608-3716)."""
        import time
        if use_gpt:
            raise NotImplementedError("use_gpt=True is the slow VLM reasoning path (graph.py:3665-3680), outside the B200 hot path")
        negative_labels = self.BACKGROUND_LABELS + ["monitor", "wall", "speaker"]
        t0 = time.time()
        floor_query, room_query, object_query = self._parse_hier(query_instruction)
        res_dict = {"object_query": object_query, "room_query": room_query, "negative_labels": negative_labels, "LLM_Parse_Time": time.time() - t0}
        floor_id = self.query_floor(floor_query) if floor_query is not None else -1
        room_ids = self.query_hmsg_room(room_query, floor_id=floor_id, query_method="label") if room_query is not None else []
        if object_query is not None:
            object_ids, room_ids, object_scores = self.query_hmsg_object(object_query, floor_id=floor_id, room_ids=room_ids, top_k=top_k,
                                                                         negative_prompt=negative_labels)
        else:
            object_ids, room_ids, object_scores = [], [], []
        res_dict["object_scores"] = object_scores
        return self._hier_result(floor_id, room_ids, object_ids, res_dict)

    def query_hierarchy_protected_icra(self, query_instruction: str, top_k: int = 1, use_gpt: bool = False):
        """graph.py:3484-3591: same flow with negative_labels = ["background"] (["wall"] for exhibition rooms)."""
        import time
        if use_gpt:
            raise NotImplementedError("use_gpt=True is the slow VLM reasoning path, outside the B200 hot path")
        negative_labels = ["background"]
        t0 = time.time()
        floor_query, room_query, object_query = self._parse_hier(query_instruction, icra=True)
        llm = time.time() - t0
        if "Exhibition" in room_query:
            negative_labels = ["wall"]
        floor_id = self.query_floor(floor_query) if floor_query is not None else -1
        room_ids = self.query_hmsg_room(room_query, floor_id=floor_id, query_method="label") if room_query is not None else []
        if object_query is not None:
            object_ids, room_ids, _ = self.query_hmsg_object(object_query, floor_id=floor_id, room_ids=room_ids, top_k=top_k, negative_prompt=negative_labels)
        else:
            object_ids, room_ids = [], []
        res_dict = {"room_query": room_query, "object_query": object_query, "negative_labels": negative_labels, "LLM_Parse_Time": llm,
                    "FastMatching": 0.0, "ObjectInImageCheck": 0.0, "VLM_Rethinking": 0.0, "Re_Matching": 0.0, "Total_Time": 0.0}
        return self._hier_result(floor_id, room_ids, object_ids, res_dict)


class Graph(B200HotPath, B200Standalone):
    """Stand-alone drop-in (no reference checkout needed): same constructor keywords as the mixin needs."""

    def __init__(self, cfg, dataset=None, clip_model=None, preprocess=None, mask_generator=None, engine=None, clip_feat_dim=512):
        """cfg: the reference's hydra config (dict / OmegaConf-like) - keys read here are the ones
        the reference reads on the hot path: pipeline.voxel_size,
        pipeline.skip_frames, pipeline.clip_bbox_margin, pipeline.clip_masked_weight,
        pipeline.max_mask_distance, pipeline.merge_type.  dataset: an RGBDDataset-like object yielding
        (rgb_image, depth_image, pose, rgb_intrinsics, depth_intrinsics) with .scale and
        .depth_intrinsics (generic.py:19-33)."""
        self.cfg = _ns(cfg) if isinstance(cfg, dict) else cfg
        self.dataset = dataset
        self.clip_model = clip_model
        self.preprocess = preprocess
        self.mask_generator = mask_generator
        self.full_pcd = PointCloud()
        self.full_feats_array = None
        self.mask_feats = []
        self.mask_pcds = []
        self.frames_feats = []
        self.frames_pcd = []
        self.objects = []
        self.rooms = []
        self.floors = []
        self.views = []
        self._b200_init(engine, clip_feat_dim)


def dropin_graph_class(reference_graph_cls, text_feats_fn=None):
    """`class Graph(B200HotPath, <reference Graph>)`: the reference's constructor, graph building, LLM glue
    and I/O stay untouched; create_feature_map and every retrieval core resolve to libhmsg_b200.so first.  Usage on the
    reference side (nav_agent/sem_nav_ctr/goal_pose_publisher.py:39, application scripts):

        from hmsg.graph.graph import Graph as RefGraph
        from holoagent_b200.memory.hmsg.graph.graph import dropin_graph_class
        Graph = dropin_graph_class(RefGraph)

    text_feats_fn(texts, clip_model, clip_feat_dim) -> [n, d]: the reference's get_text_feats_multiple_templates (its CLIP
    text tower stays where it is); defaults to the one the reference module imported."""

    class Graph(B200HotPath, reference_graph_cls):
        def __init__(self, *args, engine=None, **kwargs):
            reference_graph_cls.__init__(self, *args, **kwargs)
            self._b200_init(engine, None)
            fn = text_feats_fn
            if fn is None:
                import sys
                mod = sys.modules.get(reference_graph_cls.__module__)
                fn = getattr(mod, "get_text_feats_multiple_templates", None)
            if fn is not None:
                self.text_feats_fn = lambda texts, _fn=fn: _fn(texts, self.clip_model, self.clip_feat_dim)

    Graph.__name__ = "Graph"
    Graph.__qualname__ = "Graph"
    return Graph
