"""HmsgEngine: thin host-side owner of one hmsg_ctx (one per GPU).  Every method maps
one to one onto a C-ABI entry point of include/hmsg_b200.h; numpy arrays are host buffers,
torch CUDA tensors are passed as device pointers (on_device=1)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import VitDesc, ptr

VIT_BLOB_ORDER_HEAD = ["conv1.weight", "class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias"]
VIT_BLOB_ORDER_LAYER = ["ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                        "attn.out_proj.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                        "mlp.c_proj.weight", "mlp.c_proj.bias"]
VIT_BLOB_ORDER_TAIL = ["ln_post.weight", "ln_post.bias", "proj"]


def pack_vit_blob(state_dict, layers: int) -> np.ndarray:
    """Flatten an open_clip ``VisionTransformer.state_dict()`` (or the synthetic one from
    holoagent_b200.synth.make_vit_weights) into the float32 "encoder blob" hmsg_encoder_load
    expects (order documented in DESIGN.md)."""
    parts = []

    def get(k):
        v = state_dict[k]
        if hasattr(v, "detach"):
            v = v.detach().float().cpu().numpy()
        return np.ascontiguousarray(v, dtype=np.float32).reshape(-1)

    for k in VIT_BLOB_ORDER_HEAD:
        parts.append(get(k))
    for i in range(layers):
        for k in VIT_BLOB_ORDER_LAYER:
            parts.append(get(f"transformer.resblocks.{i}.{k}"))
    for k in VIT_BLOB_ORDER_TAIL:
        parts.append(get(k))
    return np.concatenate(parts)


def _wrap_dev(ptr_value: int, numel: int, dtype, device):
    """Zero-copy torch view of library-owned device memory (plumbing for torch.distributed)."""
    import torch
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Holder:
        __cuda_array_interface__ = {"shape": (int(numel),), "typestr": {torch.int32: "<i4", torch.float64: "<f8", torch.float32: "<f4"}[dtype],
                                    "data": (int(ptr_value), False), "version": 2, "strides": None}
    return torch.as_tensor(_Holder(), device=device)


def _is_dev(a) -> bool:
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


class HmsgError(RuntimeError):
    pass


class HmsgEngine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.hmsg_ctx_create(device, C.byref(h))
        if rc != 0:
            raise HmsgError(self.lib.hmsg_last_error(None).decode())
        self.h = h
        self.device = device
        self.H = self.W = 0
        self.d = 0
        self.n_voxels = self.n_nodes = 0
        self.vit = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.hmsg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HmsgError(f"[{rc}] " + self.lib.hmsg_last_error(self.h).decode())

    # ------------------------------------------------------------------ torch stream plumbing
    def torch_stream(self):
        """The ctx-owned CUDA stream as a torch ExternalStream (for events / ordering)."""
        import torch
        if getattr(self, "_ts", None) is None:
            self._ts = torch.cuda.ExternalStream(self.stream, device=self.device)
        return self._ts

    def wait_torch(self):
        """Order the ctx stream after everything queued on torch's current stream."""
        import torch
        self.torch_stream().wait_stream(torch.cuda.current_stream(self.device))

    def torch_wait(self):
        """Order torch's current stream after everything queued on the ctx stream."""
        import torch
        torch.cuda.current_stream(self.device).wait_stream(self.torch_stream())

    # ------------------------------------------------------------------ misc
    def sync(self):
        self._ck(self.lib.hmsg_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.hmsg_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.hmsg_launch_count(self.h))

    def set_option(self, key: str, value: int):
        self._ck(self.lib.hmsg_set_option(self.h, key.encode(), int(value)))

    PROF = {"gemm": 0, "attn": 1, "eltwise": 2, "knn": 3, "nn": 4, "scatter": 5, "geom": 6, "crops": 7, "mask3d": 8, "comm": 9}

    def prof_enable(self, *classes):
        mask = 0
        for c in classes:
            mask |= 1 << self.PROF[c]
        self._ck(self.lib.hmsg_prof_enable(self.h, mask))

    def prof_read(self, cls):
        ms, n, w = C.c_double(), C.c_int64(), C.c_double()
        self._ck(self.lib.hmsg_prof_read(self.h, self.PROF[cls], C.byref(ms), C.byref(n), C.byref(w)))
        return {"ms": ms.value, "launches": n.value, "work": w.value}

    # ------------------------------------------------------------------ scene
    def scene_begin(self, H, W, K, depth_scale, voxel_size, frame_capacity):
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        self._ck(self.lib.hmsg_scene_begin(self.h, H, W, ptr(K), float(depth_scale), float(voxel_size), int(frame_capacity)))
        self.H, self.W = H, W
        self.n_voxels = self.n_nodes = 0          # tables of the previous scene are gone

    def add_frames(self, depth, rgb, poses):
        dev = _is_dev(depth)
        if not dev:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
            poses = np.ascontiguousarray(poses, dtype=np.float64)
            n = depth.shape[0] if depth.ndim == 3 else 1
        else:
            n = depth.shape[0]
            self.wait_torch()
        self._ck(self.lib.hmsg_scene_add_frames(self.h, ptr(depth), ptr(rgb), ptr(poses), int(n), 1 if dev else 0))
        if dev:
            self.torch_wait()     # torch may recycle the (possibly temporary) input tensors only after our copy

    def set_intrinsics(self, frame_begin, K):
        """per-frame depth intrinsics: K float64 [n,3,3] for frames frame_begin.. (iphone.py:325 reads K per frame)"""
        K = np.ascontiguousarray(K, dtype=np.float64).reshape(-1, 9)
        self._ck(self.lib.hmsg_scene_set_intrinsics(self.h, int(frame_begin), int(len(K)), ptr(K)))

    @property
    def num_frames(self):
        return int(self.lib.hmsg_scene_num_frames(self.h))

    def unproject_frame(self, frame):
        hw = self.H * self.W
        xyz = np.empty((hw, 3), np.float64); rgb = np.empty((hw, 3), np.float64); valid = np.empty(hw, np.uint8)
        self._ck(self.lib.hmsg_unproject_frame(self.h, int(frame), ptr(xyz), ptr(rgb), ptr(valid)))
        return xyz, rgb, valid.astype(bool)

    def voxel_build(self):
        n = C.c_int64(0)
        mb = np.zeros(3, np.float64)
        self._ck(self.lib.hmsg_voxel_build(self.h, C.byref(n), ptr(mb)))
        self.n_voxels = n.value
        return n.value, mb

    def voxel_build_staged(self, ranges):
        """single-process walk through the staged entry points (no collectives): must equal voxel_build()"""
        mm = np.array([np.inf] * 3 + [-np.inf] * 3)
        for (b0, n) in ranges:
            t = np.zeros(6)
            self._ck(self.lib.hmsg_voxel_bounds(self.h, int(b0), int(n), ptr(t)))
            mm[:3] = np.minimum(mm[:3], t[:3]); mm[3:] = np.maximum(mm[3:], t[3:])
        self._ck(self.lib.hmsg_voxel_grid_set(self.h, ptr(mm)))
        for (b0, n) in ranges:
            self._ck(self.lib.hmsg_voxel_mark(self.h, int(b0), int(n)))
        nv = C.c_int64()
        self._ck(self.lib.hmsg_voxel_scan(self.h, C.byref(nv)))
        for (b0, n) in ranges:
            self._ck(self.lib.hmsg_voxel_accumulate(self.h, int(b0), int(n)))
        self._ck(self.lib.hmsg_voxel_finalize(self.h))
        self.n_voxels = nv.value
        return nv.value, mm[:3]

    def voxel_build_sharded(self, ranges, world):
        """Stage-wise voxel build where this rank only touches its own frame ranges [(begin, n), ...];
        partial results are merged with torch.distributed collectives (NCCL)."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device)
        mm = np.array([np.inf] * 3 + [-np.inf] * 3)
        for (b0, n) in ranges:
            t = np.zeros(6)
            self._ck(self.lib.hmsg_voxel_bounds(self.h, int(b0), int(n), ptr(t)))
            mm[:3] = np.minimum(mm[:3], t[:3]); mm[3:] = np.maximum(mm[3:], t[3:])
        import os, sys
        dbg = os.environ.get("HMSG_DEBUG")
        if dbg:
            print(f"[rank {dist.get_rank()}] local bounds {mm} over {len(ranges)} ranges", file=sys.stderr, flush=True)
        lo = torch.from_numpy(mm[:3].copy()).to(dev); hi = torch.from_numpy(mm[3:].copy()).to(dev)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        mm = np.concatenate([lo.cpu().numpy(), hi.cpu().numpy()])
        if dbg:
            print(f"[rank {dist.get_rank()}] merged bounds {mm}", file=sys.stderr, flush=True)
        self._ck(self.lib.hmsg_voxel_grid_set(self.h, ptr(mm)))
        for (b0, n) in ranges:
            self._ck(self.lib.hmsg_voxel_mark(self.h, int(b0), int(n)))
        bp, nw = C.c_void_p(), C.c_int64()
        self._ck(self.lib.hmsg_voxel_bitmap(self.h, C.byref(bp), C.byref(nw)))
        mine = _wrap_dev(bp.value, nw.value, torch.int32, dev)
        gathered = torch.empty(world * nw.value, dtype=torch.int32, device=dev)
        self.torch_wait()
        dist.all_gather_into_tensor(gathered, mine)
        self.wait_torch()
        self._ck(self.lib.hmsg_voxel_bitmap_or(self.h, ptr(gathered), int(world)))
        nv = C.c_int64()
        self._ck(self.lib.hmsg_voxel_scan(self.h, C.byref(nv)))
        for (b0, n) in ranges:
            self._ck(self.lib.hmsg_voxel_accumulate(self.h, int(b0), int(n)))
        pa, pc = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.hmsg_voxel_acc(self.h, C.byref(pa), C.byref(pc), C.byref(nv)))
        acc = _wrap_dev(pa.value, nv.value * 6, torch.float64, dev)
        cnt = _wrap_dev(pc.value, nv.value, torch.int32, dev)
        self.torch_wait()
        dist.all_reduce(acc, op=dist.ReduceOp.SUM); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        self.wait_torch()
        self._ck(self.lib.hmsg_voxel_finalize(self.h))
        self.n_voxels = nv.value
        return nv.value, mm[:3]

    def voxels_read(self):
        n = self.n_voxels
        xyz = np.empty((n, 3), np.float64); rgb = np.empty((n, 3), np.float64)
        ijk = np.empty((n, 3), np.int32); cnt = np.empty(n, np.uint32)
        self._ck(self.lib.hmsg_voxels_read(self.h, ptr(xyz), ptr(rgb), ptr(ijk), ptr(cnt)))
        return xyz, rgb, ijk, cnt

    def radius_filter(self, nb_points=1000, radius=1.0):
        n = C.c_int64(0)
        self._ck(self.lib.hmsg_radius_filter(self.h, int(nb_points), float(radius), C.byref(n)))
        self.n_nodes = n.value
        return n.value

    def radius_counts(self):
        c = np.empty(self.n_voxels, np.uint32)
        self._ck(self.lib.hmsg_radius_counts_read(self.h, ptr(c)))
        return c

    def nodes_read(self):
        n = self.n_nodes
        xyz = np.empty((n, 3), np.float64); rgb = np.empty((n, 3), np.float64)
        ijk = np.empty((n, 3), np.int32); vox = np.empty(n, np.int64)
        self._ck(self.lib.hmsg_nodes_read(self.h, ptr(xyz), ptr(rgb), ptr(ijk), ptr(vox)))
        return xyz, rgb, ijk, vox

    def pixel_to_node(self, frame, want_dist=True):
        hw = self.H * self.W
        idx = np.empty(hw, np.int64)
        dist = np.empty(hw, np.float64) if want_dist else None
        self._ck(self.lib.hmsg_pixel_to_node(self.h, int(frame), ptr(idx), ptr(dist)))
        return idx, dist

    def points_to_node(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        idx = np.empty(len(xyz), np.int64); dist = np.empty(len(xyz), np.float64)
        self._ck(self.lib.hmsg_points_to_node(self.h, ptr(xyz), len(xyz), ptr(idx), ptr(dist)))
        return idx, dist

    # ------------------------------------------------------------------ features
    def features_begin(self, d):
        self._ck(self.lib.hmsg_features_begin(self.h, int(d)))
        self.d = d

    def masks_dense(self, frame_begin, seg):
        dev = _is_dev(seg)
        if not dev:
            seg = np.ascontiguousarray(seg, dtype=np.uint8)
        n, M = seg.shape[0], seg.shape[1]
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_masks_dense(self.h, int(frame_begin), int(n), int(M), ptr(seg), 1 if dev else 0))
        if dev:
            self.torch_wait()

    def masks_boxes(self, frame_begin, xywh):
        dev = _is_dev(xywh)
        if not dev:
            xywh = np.ascontiguousarray(xywh, dtype=np.int32)
        n, M = xywh.shape[0], xywh.shape[1]
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_masks_boxes(self.h, int(frame_begin), int(n), int(M), ptr(xywh), 1 if dev else 0))
        if dev:
            self.torch_wait()

    def masks_labels(self, frame_begin, labels, M):
        """instance-id image form: labels int8 [n,H,W] (numpy or torch CUDA), mask m = pixels with label m and depth > 0"""
        dev = _is_dev(labels)
        if not dev:
            labels = np.ascontiguousarray(labels, dtype=np.int8)
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_masks_labels(self.h, int(frame_begin), int(labels.shape[0]), int(M), ptr(labels), 1 if dev else 0))
        if dev:
            self.torch_wait()

    def masks_counts(self, frame_begin, counts):
        """ragged SAM output: counts[i] real masks in frame frame_begin + i of the batch just set (slots past it are padding)"""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        self._ck(self.lib.hmsg_masks_counts(self.h, int(frame_begin), int(len(counts)), ptr(counts)))

    def fuse_scatter(self, frame_begin, n, M, feats, maskedd_weight, Fp_out=None):
        dev = _is_dev(feats)
        if not dev:
            feats = np.ascontiguousarray(feats, dtype=np.float32)
            if Fp_out is None:
                Fp_out = np.empty((n, M, self.d), np.float32)
        else:
            self.wait_torch()
        self._ck(self.lib.hmsg_fuse_scatter(self.h, int(frame_begin), int(n), int(M), ptr(feats), float(maskedd_weight), ptr(Fp_out),
                                            1 if dev else 0))
        if dev:
            self.torch_wait()
        return Fp_out

    def node_feats_finalize(self, out=None):
        dev = out is not None and _is_dev(out)
        if out is None:
            out = np.empty((self.n_nodes, self.d), np.float32)
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_node_feats_finalize(self.h, ptr(out), 1 if dev else 0))
        if dev:
            self.torch_wait()
        return out

    def node_feats_raw(self):
        s = np.empty((self.n_nodes, self.d), np.float32); c = np.empty(self.n_nodes, np.float32)
        self._ck(self.lib.hmsg_node_feats_raw(self.h, ptr(s), ptr(c)))
        return s, c

    def node_feats_device(self):
        ps, pc = C.c_void_p(), C.c_void_p()
        n, d = C.c_int64(), C.c_int32()
        self._ck(self.lib.hmsg_node_feats_device(self.h, C.byref(ps), C.byref(pc), C.byref(n), C.byref(d)))
        return ps.value, pc.value, n.value, d.value

    def mask_nodes(self, frame, down_size, M):
        off = np.zeros(M + 1, np.int64)
        self._ck(self.lib.hmsg_mask_nodes(self.h, int(frame), float(down_size), ptr(off), None, None, None))
        tot = int(off[M])
        xyz = np.empty((tot, 3), np.float64); rgb = np.empty((tot, 3), np.float64); ijk = np.empty((tot, 3), np.int32)
        if tot:
            self._ck(self.lib.hmsg_mask_nodes(self.h, int(frame), float(down_size), ptr(off), ptr(xyz), ptr(rgb), ptr(ijk)))
        return off, xyz, rgb, ijk

    def mask_nodes_batch(self, frame_begin, n_frames, down_size, filter_distance=float("inf"), keep=True):
        """A7 for frames of the current mask batch, results stay in HBM (keep: appended to the mask store)"""
        self._ck(self.lib.hmsg_mask_nodes_batch(self.h, int(frame_begin), int(n_frames), float(down_size), float(filter_distance), 1 if keep else 0))

    def mask_store_reset(self):
        self._ck(self.lib.hmsg_mask_store_reset(self.h))

    def mask_store_count(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.hmsg_mask_store_count(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def mask_store_read(self, frame, max_masks=4096):
        """-> (offsets [n_masks+1], xyz, rgb, ijk) of one stored frame (host)"""
        nm = C.c_int32()
        off = np.zeros(max_masks + 1, np.int64)
        self._ck(self.lib.hmsg_mask_store_read(self.h, int(frame), C.byref(nm), ptr(off), None, None, None))
        off = off[:nm.value + 1].copy()
        tot = int(off[-1])
        xyz = np.empty((tot, 3), np.float64); rgb = np.empty((tot, 3), np.float64); ijk = np.empty((tot, 3), np.int32)
        if tot:
            self._ck(self.lib.hmsg_mask_store_read(self.h, int(frame), C.byref(nm), ptr(off), ptr(xyz), ptr(rgb), ptr(ijk)))
        return off, xyz, rgb, ijk

    def objects_merge_stored(self, frame_begin, n_frames):
        """seq_merge iterations over stored frames (graph_utils.py:1015-1038), fed from HBM"""
        self._ck(self.lib.hmsg_objects_merge_stored(self.h, int(frame_begin), int(n_frames)))

    # ------------------------------------------------------------------ N1 object instances
    def objects_begin(self, overlap_thresh=0.75, down_size=0.05, iou_thresh=0.05):
        """seq_merge(frames_pcd, th, down_size, proxy_th) state reset (graph_utils.py:1015-1021)"""
        self._ck(self.lib.hmsg_objects_begin(self.h, float(overlap_thresh), float(down_size), float(iou_thresh)))

    def objects_add_masks(self, off, xyz, rgb=None):
        """one seq_merge iteration: ragged frame masks (host arrays: off int64 [n+1], xyz/rgb float64 [P,3])"""
        off = np.ascontiguousarray(off, dtype=np.int64)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        if rgb is not None:
            rgb = np.ascontiguousarray(rgb, dtype=np.float64).reshape(-1, 3)
        self._ck(self.lib.hmsg_objects_add_masks(self.h, int(len(off) - 1), ptr(off), ptr(xyz) if len(xyz) else None,
                                                 ptr(rgb) if (rgb is not None and len(rgb)) else None, 0))

    def objects_add_frame(self, frame, down_size, filter_distance=float("inf")):
        """one seq_merge iteration fed from the device: create_3d_masks(frame) -> merge, nothing visits the host"""
        self._ck(self.lib.hmsg_objects_add_frame(self.h, int(frame), float(down_size), float(filter_distance)))

    def objects_finish(self, min_points=10):
        n, p = C.c_int64(), C.c_int64()
        self._ck(self.lib.hmsg_objects_finish(self.h, int(min_points), C.byref(n), C.byref(p)))
        return n.value, p.value

    def objects_count(self):
        n, p, g = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.hmsg_objects_count(self.h, C.byref(n), C.byref(p), C.byref(g)))
        return n.value, p.value, g.value

    def objects_read(self):
        n, p, _ = self.objects_count()
        off = np.zeros(n + 1, np.int64); xyz = np.empty((p, 3), np.float64); rgb = np.empty((p, 3), np.float64)
        self._ck(self.lib.hmsg_objects_read(self.h, ptr(off), ptr(xyz) if p else None, ptr(rgb) if p else None))
        return off, xyz, rgb

    def object_feats(self, full_feats, voxel_size, max_dist=0.8, eps=0.01, min_points=100):
        """graph.py:451-488: one feature row per object of objects_finish() (host numpy in / out)"""
        full_feats = np.ascontiguousarray(full_feats, dtype=np.float32)
        n, _, _ = self.objects_count()
        out = np.zeros((n, full_feats.shape[1]), np.float32)
        self._ck(self.lib.hmsg_object_feats(self.h, ptr(full_feats), int(full_feats.shape[1]), float(voxel_size), float(max_dist), float(eps),
                                            int(min_points), ptr(out), 0))
        return out

    # ------------------------------------------------------------------ encoder
    def encoder_load(self, state_dict, image=224, patch=32, width=768, layers=12, heads=12, mlp=3072, out_dim=512, quick_gelu=False):
        blob = pack_vit_blob(state_dict, layers)
        desc = VitDesc(image, patch, width, layers, heads, mlp, out_dim, 1 if quick_gelu else 0)
        self._ck(self.lib.hmsg_encoder_load(self.h, C.byref(desc), ptr(blob), blob.size))
        self.vit = desc

    def encode_images(self, x, normalize=True, out=None):
        if self.vit is None:
            raise HmsgError("[3] encoder: call hmsg_encoder_load first")     # same status / text as the library's own check
        dev = _is_dev(x)
        B = x.shape[0]
        if not dev:
            x = np.ascontiguousarray(x, dtype=np.float32)
            out = np.empty((B, self.vit.out_dim), np.float32)
        elif out is None:
            import torch
            out = torch.empty((B, self.vit.out_dim), dtype=torch.float32, device=x.device)
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_encode_images(self.h, ptr(x), int(B), ptr(out), 1 if normalize else 0, 1 if dev else 0))
        if dev:
            self.torch_wait()
        return out

    def encode_images_ptr(self, x_ptr: int, B: int, out, normalize=True):
        """device pointer in (e.g. the ctx crop buffer), torch CUDA tensor out"""
        self._ck(self.lib.hmsg_encode_images(self.h, C.c_void_p(x_ptr), int(B), ptr(out), 1 if normalize else 0, 1))
        self.torch_wait()
        return out

    # ------------------------------------------------------------------ crops (A8 / N3)
    def has_device_crops(self) -> bool:
        return True

    def make_crops(self, frame_begin, n, M, xywh, bbox_margin) -> int:
        """-> device pointer of [n, 2M+1, 3, 224, 224] float32 (masked, plain, full-frame order)"""
        dev = _is_dev(xywh)
        if not dev:
            xywh = np.ascontiguousarray(xywh, dtype=np.int32)
        else:
            self.wait_torch()
        out = C.c_void_p()
        self._ck(self.lib.hmsg_make_crops(self.h, int(frame_begin), int(n), int(M), ptr(xywh), int(bbox_margin), 1 if dev else 0, C.byref(out)))
        if dev:
            self.torch_wait()
        return out.value

    def encode_crops(self, frame_begin, n, M, xywh, bbox_margin, feats_out):
        """fused crops -> patch matrix -> encoder; feats_out: torch CUDA tensor [n*(2M+1), d]"""
        dev = _is_dev(xywh)
        if not dev:
            xywh = np.ascontiguousarray(xywh, dtype=np.int32)
        else:
            self.wait_torch()
        self._ck(self.lib.hmsg_encode_crops(self.h, int(frame_begin), int(n), int(M), ptr(xywh), int(bbox_margin), 1 if dev else 0, ptr(feats_out)))
        self.torch_wait()
        return feats_out

    def crops_read(self, n_crops):
        a = np.empty((n_crops, 3, 224, 224), np.float32)
        self._ck(self.lib.hmsg_crops_read(self.h, int(n_crops), ptr(a)))
        return a

    # ------------------------------------------------------------------ host-buffer variants used by bench e2e / multi-GPU
    def scene_reset_frames(self):
        self._ck(self.lib.hmsg_scene_reset_frames(self.h))

    def add_frames_host(self, depth_t, rgb_t, poses_np):
        """pinned torch CPU tensors (depth int16-viewed uint16, rgb uint8) + float64 poses"""
        poses_np = np.ascontiguousarray(poses_np, dtype=np.float64)
        self._ck(self.lib.hmsg_scene_add_frames(self.h, ptr(depth_t), ptr(rgb_t), ptr(poses_np), int(depth_t.shape[0]), 0))

    def set_num_frames(self, n):
        """declare the scene length when frames were put at explicit ids (ranks hold only their own)"""
        self._ck(self.lib.hmsg_scene_set_num_frames(self.h, int(n)))

    def put_frames_host(self, frame_begin, depth_t, rgb_t, poses_np):
        poses_np = np.ascontiguousarray(poses_np, dtype=np.float64)
        self._ck(self.lib.hmsg_scene_put_frames(self.h, int(frame_begin), ptr(depth_t), ptr(rgb_t), ptr(poses_np), int(depth_t.shape[0]), 0))

    def put_rgb_host(self, frame_begin, rgb):
        """swap the colour images of stored frames (geometry stays valid): hmsg_scene_put_rgb"""
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        self._ck(self.lib.hmsg_scene_put_rgb(self.h, int(frame_begin), ptr(rgb), int(rgb.shape[0]), 0))
        self.sync()          # pageable source: the copy must be done before `rgb` can go away

    def node_feats_finalize_host(self, out_t):
        self._ck(self.lib.hmsg_node_feats_finalize(self.h, ptr(out_t), 0))

    def pack_partials(self, dst, Fp_rows, fp_floats):
        self.wait_torch()
        self._ck(self.lib.hmsg_node_feats_pack(self.h, ptr(dst), ptr(Fp_rows), int(fp_floats)))
        self.torch_wait()

    def merge_partials(self, gathered, world, stride):
        self._ck(self.lib.hmsg_node_feats_merge(self.h, ptr(gathered), int(world), int(stride)))
        self.torch_wait()

    # ------------------------------------------------------------------ multi-GPU through the C-ABI (NCCL inside the library)
    def comm_init_torch(self):
        """Give the ctx its own NCCL communicator; the 128-byte unique id travels over the already
        initialised torch.distributed group (plumbing only - every data-path collective runs inside
        libhmsg_b200.so on the ctx stream)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = np.zeros(128, np.uint8)
        if rank == 0:
            rc = self.lib.hmsg_comm_unique_id(ptr(uid))
            if rc != 0:
                raise HmsgError("hmsg_comm_unique_id failed (NCCL not loadable)")
        t = torch.from_numpy(uid)
        if dist.get_backend() == "nccl":
            t = t.to(torch.device("cuda", self.device))
        dist.broadcast(t, 0)
        uid = t.cpu().numpy().copy()
        self._ck(self.lib.hmsg_comm_init(self.h, ptr(uid), int(rank), int(world)))
        self.comm_rank, self.comm_world = rank, world

    def comm_init_local(self, uid=None, rank=0, world=1):
        """communicator without torch.distributed: uid = bytes from comm_unique_id() of rank 0 (world 1: made here)"""
        if uid is None:
            uid = self.comm_unique_id()
        uid = np.ascontiguousarray(np.frombuffer(bytes(uid), dtype=np.uint8))
        self._ck(self.lib.hmsg_comm_init(self.h, ptr(uid), int(rank), int(world)))
        self.comm_rank, self.comm_world = rank, world

    def comm_unique_id(self) -> bytes:
        uid = np.zeros(128, np.uint8)
        if self.lib.hmsg_comm_unique_id(ptr(uid)) != 0:
            raise HmsgError("hmsg_comm_unique_id failed (NCCL not loadable)")
        return uid.tobytes()

    def comm_info(self):
        r, w, b = C.c_int32(), C.c_int32(), C.c_double()
        self._ck(self.lib.hmsg_comm_info(self.h, C.byref(r), C.byref(w), C.byref(b)))
        return r.value, w.value, b.value

    def voxel_build_sharded_c(self, ranges, comm=None):
        """hmsg_voxel_build_sharded: this rank's frame ranges [(begin, n), ...], collectives inside the library"""
        rg = np.ascontiguousarray(np.asarray(ranges, dtype=np.int64).reshape(-1, 2))
        nv = C.c_int64(); mb = np.zeros(3, np.float64)
        self._ck(self.lib.hmsg_voxel_build_sharded(self.h, C.c_void_p(comm) if comm else None, ptr(rg) if len(rg) else None, int(len(rg)), C.byref(nv), ptr(mb)))
        self.n_voxels = nv.value
        return nv.value, mb

    def radius_filter_sharded(self, nb_points=1000, radius=1.0, comm=None):
        n = C.c_int64(0)
        self._ck(self.lib.hmsg_radius_filter_sharded(self.h, C.c_void_p(comm) if comm else None, int(nb_points), float(radius), C.byref(n)))
        self.n_nodes = n.value
        return n.value

    def allgather_nodes(self, Fp_local=None, fp_floats=0, Fp_all=None, fp_stride=0, comm=None):
        if Fp_all is not None:
            self.wait_torch()
        self._ck(self.lib.hmsg_allgather_nodes(self.h, C.c_void_p(comm) if comm else None, ptr(Fp_local), int(fp_floats), ptr(Fp_all), int(fp_stride)))
        if Fp_all is not None:
            self.torch_wait()

    def gemm_debug(self, A_f16, W_f16, C_f32, M, N, K):
        self.wait_torch()
        self._ck(self.lib.hmsg_gemm_f16_debug(self.h, ptr(A_f16), ptr(W_f16), ptr(C_f32), M, N, K))

    # ------------------------------------------------------------------ retrieval
    def index_set(self, E, borrow=False):
        dev = _is_dev(E)
        if not dev:
            E = np.ascontiguousarray(E, dtype=np.float32)
        self._E_keepalive = E if borrow else None
        self._graph_index_tag = None          # whatever a Graph mirror cached in this engine is gone now
        self.index_N, self.index_d = E.shape
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_index_set(self.h, ptr(E), int(E.shape[0]), int(E.shape[1]), (2 if borrow else 1) if dev else 0))
        if dev:
            self.torch_wait()

    def query_topk(self, Q, k, row_mask=None, ids=None, scores=None):
        dev = _is_dev(Q)
        nq = Q.shape[0]
        if not dev:
            Q = np.ascontiguousarray(Q, dtype=np.float32)
            if row_mask is not None:
                row_mask = np.ascontiguousarray(row_mask, dtype=np.uint8)
            ids = np.empty((nq, k), np.int64); scores = np.empty((nq, k), np.float32)
        elif ids is None:
            import torch
            ids = torch.empty((nq, k), dtype=torch.int64, device=Q.device)
            scores = torch.empty((nq, k), dtype=torch.float32, device=Q.device)
        if dev:
            self.wait_torch()
        self._ck(self.lib.hmsg_query_topk(self.h, ptr(Q), int(nq), int(k), ptr(row_mask), ptr(ids), ptr(scores), 1 if dev else 0))
        if dev:
            self.torch_wait()
        return ids, scores

    def query_scores(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, self.index_d)
        out = np.empty((Q.shape[0], self.index_N), np.float32)
        self._ck(self.lib.hmsg_query_scores(self.h, ptr(Q), Q.shape[0], ptr(out), 0))
        return out

    def pixel_feature_map(self, frame):
        out = np.empty((self.H * self.W, self.d), np.float16)
        self._ck(self.lib.hmsg_pixel_feature_map(self.h, int(frame), ptr(out)))
        return out

    def query_object(self, Q, query_id, k, row_mask=None):
        """Q [n_req, Qp, d] host float32.  Returns ids [n_req,k], scores [n_req,k], n_found [n_req]."""
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        n_req, Qp = Q.shape[0], Q.shape[1]
        if row_mask is not None:
            row_mask = np.ascontiguousarray(row_mask, dtype=np.uint8)
        ids = np.empty((n_req, k), np.int64); scores = np.empty((n_req, k), np.float32); nf = np.empty(n_req, np.int32)
        self._ck(self.lib.hmsg_query_object(self.h, ptr(Q), n_req, Qp, int(query_id), int(k), ptr(row_mask), ptr(ids), ptr(scores), ptr(nf), 0))
        return ids, scores, nf
