"""Frame-batched HMSG ingest driver: the loop body of Graph.create_feature_map
(fsr_vln/memory/hmsg/graph/graph.py:339-415) expressed over C-ABI calls, for one rank.

Multi-GPU (option A): rank r owns the frame batches b with b % world == r for BOTH
phases.  Geometry: local bounds / occupancy / accumulation over the rank's frames merged by three
tiny collectives (min-max of 6 doubles, all-gather + OR of the bitmap, sum of the f64 voxel
accumulators) so that every rank ends with the identical voxel / node table; the radius filter runs
replicated on the merged table.  Features: crops -> encoder -> fusion -> scatter on the rank's frames,
then ONE NCCL all-gather carries each rank's packed [partial sum_features | counter | F_p rows]; the
partials are summed in rank order by hmsg_node_feats_merge (deterministic).  (The fp64-pipe-bound
geometry passes cost ~15 us/frame; replicating them on every rank would cap 8-GPU efficiency at ~88 %.)"""
from __future__ import annotations

import numpy as np


def shard_batches(n_frames: int, frame_batch: int, world: int, rank: int):
    """Frame batches (begin, count) of the whole job and the ones rank `rank` owns (b % world == rank)."""
    batches = [(b0, min(frame_batch, n_frames - b0)) for b0 in range(0, n_frames, frame_batch)]
    return batches, [b for i, b in enumerate(batches) if i % world == rank]


def packed_layout(n_nodes: int, d: int, n_frames: int, frame_batch: int, world: int, M: int):
    """Per-rank send buffer of the single all-gather: [sum_features n*d | counter n | F_p rows].
    Returns (partial_floats, stride_floats) - stride is padded to the largest per-rank F_p block."""
    n_batches = -(-n_frames // frame_batch)
    nmax_local = -(-n_batches // world) * frame_batch
    part = n_nodes * d + n_nodes
    return part, part + nmax_local * M * d


class IngestJob:
    def __init__(self, eng, n_frames, frame_batch, M, d, boxes_dev, rank=0, world=1, crops="auto", maskedd_weight=0.4418, bbox_margin=50,
                 nb_points=1000, radius=1.0):
        import torch
        self.torch = torch
        self.eng, self.F, self.FB, self.M, self.d = eng, n_frames, frame_batch, M, d
        self.rank, self.world = rank, world
        self.boxes_dev = boxes_dev
        self.w, self.margin, self.nb, self.radius = maskedd_weight, bbox_margin, nb_points, radius
        self.batches, self.my_batches = shard_batches(n_frames, frame_batch, world, rank)
        self.n_local = sum(n for _, n in self.my_batches)
        dev = boxes_dev.device
        B = frame_batch * (2 * M + 1)
        self.feats = torch.empty((B, d), dtype=torch.float32, device=dev)
        self.Fp_local = torch.empty((max(self.n_local, 1), M, d), dtype=torch.float32, device=dev)
        self.crops_mode = crops
        if crops in ("auto", "device"):
            if eng.has_device_crops():
                self.crops_mode = "device"
            elif crops == "device":
                raise RuntimeError("device crops requested but hmsg_make_crops is not available")
            else:
                self.crops_mode = "synthetic"
        self.syn_crops = None
        if self.crops_mode == "synthetic":
            g = torch.Generator(device=dev).manual_seed(1)
            self.syn_crops = torch.randn((B, 3, 224, 224), generator=g, device=dev, dtype=torch.float32)
        self.full_feats = None
        self.gather_buf = None
        self.host = None
        self.h2d_bytes = self.d2h_bytes = 0

    # ------------------------------------------------------------------
    def _features_pass(self, boxes_host=None):
        eng, M, d = self.eng, self.M, self.d
        off = 0
        for (b0, n) in self.my_batches:
            B = n * (2 * M + 1)
            if boxes_host is not None:
                eng.masks_boxes(b0, boxes_host[b0:b0 + n])          # host XYWH -> staged H2D inside
            else:
                eng.masks_boxes(b0, self.boxes_dev[b0:b0 + n])
            if self.crops_mode == "device":
                eng.encode_crops(b0, n, M, (boxes_host if boxes_host is not None else self.boxes_dev)[b0:b0 + n], self.margin, self.feats)
            else:
                eng.encode_images(self.syn_crops[:B], out=self.feats)
            eng.fuse_scatter(b0, n, M, self.feats[:B].view(n, 2 * M + 1, d), self.w, Fp_out=self.Fp_local[off:off + n])
            off += n

    def _merge(self):
        """all-gather of packed partials + deterministic rank-order sum (world > 1 only)."""
        if self.world == 1:
            return
        import torch.distributed as dist
        torch = self.torch
        eng = self.eng
        ps, pc, n, d = eng.node_feats_device()
        part, stride = packed_layout(n, d, self.F, self.FB, self.world, self.M)
        if self.gather_buf is None or self.gather_buf.numel() != self.world * stride:
            self.gather_buf = torch.empty(self.world * stride, dtype=torch.float32, device=self.boxes_dev.device)
            self.send_buf = torch.zeros(stride, dtype=torch.float32, device=self.boxes_dev.device)
        eng.pack_partials(self.send_buf, self.Fp_local, self.n_local * self.M * d)
        eng.torch_wait()
        dist.all_gather_into_tensor(self.gather_buf, self.send_buf)
        eng.wait_torch()
        eng.merge_partials(self.gather_buf, self.world, stride)

    def _geometry(self):
        """voxel table: single GPU = all frames; world > 1 = this rank's frame batches + tiny collectives"""
        if self.world == 1:
            self.eng.voxel_build()
        else:
            self.eng.voxel_build_sharded(self.my_batches, self.world)

    def step_device(self):
        """One whole build with frames resident in HBM."""
        eng = self.eng
        self._geometry()
        eng.radius_filter(self.nb, self.radius)
        eng.features_begin(self.d)
        self._features_pass()
        self._merge()
        if self.full_feats is None or self.full_feats.shape[0] != eng.n_nodes:
            self.full_feats = self.torch.empty((eng.n_nodes, self.d), dtype=self.torch.float32, device=self.boxes_dev.device)
        eng.node_feats_finalize(self.full_feats)

    # ------------------------------------------------------------------
    def bind_host(self, host_depth, host_rgb, poses, boxes_np):
        """host_depth / host_rgb: pinned tensors holding THIS RANK's frames in my_batches order
        ([n_local,H,W] / [n_local,H,W,3]); poses / boxes: full [F,...] arrays."""
        self.host = (host_depth, host_rgb, np.ascontiguousarray(poses, dtype=np.float64), np.ascontiguousarray(boxes_np, dtype=np.int32))
        self.h2d_bytes = host_depth.numel() * 2 + host_rgb.numel() + self.n_local * 16 * 8 + self.n_local * self.M * 16
        self.host_out = None

    def step_host(self):
        """Same build through host buffers: pinned frames -> H2D, node features -> D2H."""
        eng = self.eng
        hd, hr, poses, boxes = self.host
        eng.scene_reset_frames()
        off = 0
        for (b0, n) in self.my_batches:       # only this rank's frames cross PCIe
            eng.put_frames_host(b0, hd[off:off + n], hr[off:off + n], poses[b0:b0 + n])
            off += n
        eng.set_num_frames(self.F)
        self._geometry()
        eng.radius_filter(self.nb, self.radius)
        eng.features_begin(self.d)
        self._features_pass(boxes_host=boxes)
        self._merge()
        if self.host_out is None or self.host_out.shape[0] != eng.n_nodes:
            self.host_out = self.torch.empty((eng.n_nodes, self.d), dtype=self.torch.float32).pin_memory()
        eng.node_feats_finalize_host(self.host_out)
        self.d2h_bytes = self.host_out.numel() * 4

    def release(self):
        self.syn_crops = None
        self.gather_buf = None
        self.feats = None
