"""Frame-batched HMSG ingest driver: the loop body of Graph.create_feature_map
(fsr_vln/memory/hmsg/graph/graph.py:339-415) expressed over C-ABI calls, for one rank.

Per frame batch (graph.py:373-411): masks -> 2M+1 crops -> encoder (A8/A9) -> mask-feature fusion (A5) ->
pixel->node NN + winner election + node feature scatter (A4/A6) -> per-mask 3-D node sets (A7, create_3d_masks,
graph.py:391-402: kept in the HBM mask store = the reference's `frames_pcd`).

Multi-GPU (option A): rank r owns one contiguous block of ceil/floor(F / world) frames for BOTH
phases (per-frame sharding: the ranks differ by at most one frame; batches inside the block are `frame_batch`
frames with a ragged tail).  A rank's scene holds ONLY its own block, numbered 0 .. n_local-1 (the C-ABI never
sees global frame ids; 50 k x 1280x720 frames = 230 GB do not fit one GPU, 6 250 per rank do).  Geometry: local bounds / occupancy / accumulation over the rank's frames merged by three
tiny collectives so that every rank ends with the identical voxel table; the radius filter counts neighbours for a
slice of the voxel table per rank.  Features: every rank scatters its own frames into a dense partial, then
hmsg_allgather_nodes exchanges row slices all-to-all, sums them in rank order and all-gathers the result.
All data-path collectives run inside libhmsg_b200.so (NCCL on the ctx stream, `collective="c"`); the older
torch.distributed form (`collective="torch"`: one all_gather_into_tensor of dense partials) is kept for A/B runs."""
from __future__ import annotations

import numpy as np


def frame_block(n_frames: int, world: int, rank: int):
    """Contiguous frame block (begin, count) of rank `rank`: the first n_frames % world ranks hold one frame more."""
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def shard_batches(n_frames: int, frame_batch: int, world: int, rank: int):
    """Frame batches (begin, count) of the whole job (rank blocks in rank order) and the ones rank `rank` owns."""
    allb, mine = [], []
    for r in range(world):
        b0, cnt = frame_block(n_frames, world, r)
        bs = [(f0, min(frame_batch, b0 + cnt - f0)) for f0 in range(b0, b0 + cnt, frame_batch)]
        allb += bs
        if r == rank:
            mine = bs
    return allb, mine


def packed_layout(n_nodes: int, d: int, n_frames: int, frame_batch: int, world: int, M: int):
    """collective="torch": per-rank send buffer of the single all-gather: [sum_features n*d | counter n | F_p rows].
    Returns (partial_floats, stride_floats) - stride is padded to the largest per-rank F_p block."""
    nmax_local = -(-n_frames // world)
    part = n_nodes * d + n_nodes
    return part, part + nmax_local * M * d


class IngestJob:
    """One rank's share of the build.  `n_frames` = the rank's LOCAL frame count (its scene holds frames 0..n_frames-1,
    `boxes_dev` [n_frames, M, 4] are their mask rectangles); `total_frames` = the job size over all ranks."""

    def __init__(self, eng, n_frames, frame_batch, M, d, boxes_dev, rank=0, world=1, crops="auto", maskedd_weight=0.4418, bbox_margin=50,
                 nb_points=1000, radius=1.0, a7=True, voxel_size=0.05, max_mask_distance=10000.0, collective="c", gather_fp=False,
                 total_frames=None, labels_dev=None):
        import torch
        self.torch = torch
        self.eng, self.F, self.FB, self.M, self.d = eng, n_frames, frame_batch, M, d
        self.rank, self.world = rank, world
        self.total = total_frames if total_frames is not None else n_frames
        self.boxes_dev = boxes_dev
        self.w, self.margin, self.nb, self.radius = maskedd_weight, bbox_margin, nb_points, radius
        self.a7, self.vs, self.max_mask_distance = a7, voxel_size, max_mask_distance
        self.collective, self.gather_fp = collective, gather_fp
        self.labels_dev = labels_dev          # int8 [n_frames,H,W] instance-id images: masks come from them instead of the rectangles
        self.my_batches = [(f0, min(frame_batch, n_frames - f0)) for f0 in range(0, n_frames, frame_batch)]
        self.n_local = n_frames
        self.nmax_local = -(-self.total // world)
        dev = boxes_dev.device
        B = frame_batch * (2 * M + 1)
        self.feats = torch.empty((B, d), dtype=torch.float32, device=dev)
        self.Fp_local = torch.empty((max(self.nmax_local, 1), M, d), dtype=torch.float32, device=dev)
        self.Fp_all = None
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
        if world > 1 and collective == "c" and getattr(eng, "comm_world", None) != world:
            eng.comm_init_torch()

    # ------------------------------------------------------------------
    def _features_pass(self, boxes_host=None):
        eng, M, d = self.eng, self.M, self.d
        off = 0
        if self.a7:
            eng.mask_store_reset()
        for i, (b0, n) in enumerate(self.my_batches):
            B = n * (2 * M + 1)
            if self.labels_dev is not None:
                eng.masks_labels(b0, self.labels_dev[b0:b0 + n], M)
            elif boxes_host is not None:
                eng.masks_boxes(b0, boxes_host[b0:b0 + n])          # host XYWH -> staged H2D inside
            else:
                eng.masks_boxes(b0, self.boxes_dev[b0:b0 + n])
            if self.crops_mode == "device":
                eng.encode_crops(b0, n, M, (boxes_host if boxes_host is not None else self.boxes_dev)[b0:b0 + n], self.margin, self.feats)
            else:
                eng.encode_images(self.syn_crops[:B], out=self.feats)
            eng.fuse_scatter(b0, n, M, self.feats[:B].view(n, 2 * M + 1, d), self.w, Fp_out=self.Fp_local[off:off + n])
            if self.a7:      # create_3d_masks for every frame of the batch (graph.py:391-402), kept in the HBM mask store
                eng.mask_nodes_batch(b0, n, self.vs, self.max_mask_distance, keep=True)
            off += n

    def _merge(self):
        """node-embedding merge across ranks (world > 1 only)."""
        if self.world == 1:
            return
        torch, eng = self.torch, self.eng
        if self.collective == "c":
            if self.gather_fp:
                stride = self.nmax_local * self.M * self.d
                if self.Fp_all is None:
                    self.Fp_all = torch.empty((self.world, stride), dtype=torch.float32, device=self.boxes_dev.device)
                eng.allgather_nodes(self.Fp_local, self.n_local * self.M * self.d, self.Fp_all, stride)
            else:
                eng.allgather_nodes()
            return
        import torch.distributed as dist
        ps, pc, n, d = eng.node_feats_device()
        part, stride = packed_layout(n, d, self.total, self.FB, self.world, self.M)
        if self.gather_buf is None or self.gather_buf.numel() != self.world * stride:
            self.gather_buf = torch.empty(self.world * stride, dtype=torch.float32, device=self.boxes_dev.device)
            self.send_buf = torch.zeros(stride, dtype=torch.float32, device=self.boxes_dev.device)
        eng.pack_partials(self.send_buf, self.Fp_local, self.n_local * self.M * d)
        eng.torch_wait()
        dist.all_gather_into_tensor(self.gather_buf, self.send_buf)
        eng.wait_torch()
        eng.merge_partials(self.gather_buf, self.world, stride)

    def _geometry(self):
        """voxel + node table: single GPU = all frames; world > 1 = this rank's frame block + tiny collectives"""
        eng = self.eng
        if self.world == 1:
            eng.voxel_build()
            eng.radius_filter(self.nb, self.radius)
        elif self.collective == "c":
            eng.voxel_build_sharded_c([(0, self.F)] if self.F else [])
            eng.radius_filter_sharded(self.nb, self.radius)
        else:
            eng.voxel_build_sharded([(0, self.F)] if self.F else [], self.world)
            eng.radius_filter(self.nb, self.radius)

    def step_device(self):
        """One whole build with frames resident in HBM."""
        eng = self.eng
        self._geometry()
        eng.features_begin(self.d)
        self._features_pass()
        self._merge()
        if self.full_feats is None or self.full_feats.shape[0] != eng.n_nodes:
            self.full_feats = self.torch.empty((eng.n_nodes, self.d), dtype=self.torch.float32, device=self.boxes_dev.device)
        eng.node_feats_finalize(self.full_feats)

    # ------------------------------------------------------------------
    def bind_host(self, host_depth, host_rgb, poses, boxes_np):
        """host_depth / host_rgb: pinned tensors holding THIS RANK's frames ([n_local,H,W] / [n_local,H,W,3]);
        poses [n_local,16] / boxes [n_local,M,4]: this rank's rows."""
        self.host = (host_depth, host_rgb, np.ascontiguousarray(poses, dtype=np.float64), np.ascontiguousarray(boxes_np, dtype=np.int32))
        self.h2d_bytes = (host_depth[:self.n_local].numel() * 2 + host_rgb[:self.n_local].numel() + self.n_local * 16 * 8 +
                          self.n_local * self.M * 16)
        self.host_out = None

    def step_host(self):
        """Same build through host buffers: pinned frames -> H2D, node features -> D2H.

        The uploads are queued on the ctx's copy stream up front (hmsg_scene_put_frames is asynchronous, one event
        per batch).  Crops + encoder of a batch need only that batch's pixels, not the node table, so they run
        first, batch by batch, each waiting for its own upload only: PCIe transfer of later batches overlaps the
        tensor-core work of earlier ones.  Geometry (needs every frame) follows, then fusion / NN / scatter / A7 per
        batch from the stored encoder outputs ((2M+1) x d floats per frame)."""
        eng, torch = self.eng, self.torch
        hd, hr, poses, boxes = self.host
        M, d = self.M, self.d
        eng.scene_reset_frames()
        for (b0, n) in self.my_batches:       # only this rank's frames cross PCIe
            eng.put_frames_host(b0, hd[b0:b0 + n], hr[b0:b0 + n], poses[b0:b0 + n])
        eng.set_num_frames(self.F)
        if self.crops_mode == "device":
            if getattr(self, "feats_all", None) is None:
                self.feats_all = torch.empty((max(self.F, 1) * (2 * M + 1), d), dtype=torch.float32, device=self.boxes_dev.device)
            for (b0, n) in self.my_batches:
                eng.masks_boxes(b0, boxes[b0:b0 + n])
                eng.encode_crops(b0, n, M, boxes[b0:b0 + n], self.margin, self.feats_all[b0 * (2 * M + 1):(b0 + n) * (2 * M + 1)])
        self._geometry()
        eng.features_begin(d)
        if self.a7:
            eng.mask_store_reset()
        for (b0, n) in self.my_batches:
            B = n * (2 * M + 1)
            eng.masks_boxes(b0, boxes[b0:b0 + n])
            if self.crops_mode == "device":
                fe = self.feats_all[b0 * (2 * M + 1):(b0 + n) * (2 * M + 1)]
            else:
                fe = eng.encode_images(self.syn_crops[:B], out=self.feats)[:B]
            eng.fuse_scatter(b0, n, M, fe.view(n, 2 * M + 1, d), self.w, Fp_out=self.Fp_local[b0:b0 + n])
            if self.a7:
                eng.mask_nodes_batch(b0, n, self.vs, self.max_mask_distance, keep=True)
        self._merge()
        if self.rank == 0:                     # the build's result (full_feats_array) lands on ONE host process, as in the reference
            if self.host_out is None or self.host_out.shape[0] != eng.n_nodes:
                self.host_out = torch.empty((eng.n_nodes, d), dtype=torch.float32).pin_memory()
            eng.node_feats_finalize_host(self.host_out)
            self.d2h_bytes = self.host_out.numel() * 4
        else:                                  # the other ranks keep their (identical) copy in HBM for retrieval
            if self.full_feats is None or self.full_feats.shape[0] != eng.n_nodes:
                self.full_feats = torch.empty((eng.n_nodes, d), dtype=torch.float32, device=self.boxes_dev.device)
            eng.node_feats_finalize(self.full_feats)
            self.d2h_bytes = 0

    def release(self):
        self.syn_crops = None
        self.gather_buf = None
        self.feats = None
        self.Fp_all = None
        self.feats_all = None
