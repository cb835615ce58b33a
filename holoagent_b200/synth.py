"""Synthetic posed RGB-D frames, SAM-like masks and kNN tables.

The reference consumes posed RGB-D sequences from disk (uint16 depth in mm, uint8
RGB, 4x4 camera-to-world poses; fsr_vln/memory/hmsg/dataloader/horizon.py:217-268,
hm3dsem.py:129-155).  No dataset is available offline, so benchmarks and parity
tests use a seeded analytic scene: a camera moving inside an axis-aligned
12 x 3 x 9 m room that contains 6 boxes.  Depth is the pin-hole Z-depth of the
first ray/surface hit, quantised to uint16 millimetres exactly like the
reference's data writers do ((depth_m * 1000).astype(uint16)); 5 % of the pixels
are zeroed ("holes") and everything beyond ``depth_cut`` is zeroed
(horizon.py:258-261).  Intrinsics follow the hfov-90-degree formula of
hm3dsem.py:147-155.  Poses are an analytic closed trajectory (ellipse, rotating
yaw, small pitch); the reference's HM3D pose files are data under
/root/reference and cannot travel to the GPU box, so they are not used.

Written with torch ops only, so the same code generates frames on the CPU (tests,
bit-reproducible) or directly in HBM (large benchmark runs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

ROOM = (12.0, 3.0, 9.0)  # x, y (height), z extents in metres
# 6 boxes: (xmin, ymin, zmin, xmax, ymax, zmax)
BOXES = (
    (1.0, 0.0, 1.0, 2.2, 1.1, 2.5),
    (4.0, 0.0, 6.5, 6.5, 0.8, 7.6),
    (9.2, 0.0, 1.2, 10.4, 2.0, 2.0),
    (8.0, 0.0, 6.0, 9.0, 1.4, 7.5),
    (5.2, 0.0, 0.4, 6.8, 0.9, 1.3),
    (2.5, 0.0, 4.0, 3.3, 0.5, 4.9),
)


def intrinsics(H: int, W: int) -> np.ndarray:
    """hfov = 90 deg pin-hole matrix (hm3dsem.py:147-155), float64 [3,3]."""
    hfov = 90 * np.pi / 180
    vfov = 2 * math.atan(np.tan(hfov / 2) * H / W)
    fx = W / (2.0 * np.tan(hfov / 2.0))
    fy = H / (2.0 * np.tan(vfov / 2.0))
    return np.array([[fx, 0.0, W / 2], [0.0, fy, H / 2], [0.0, 0.0, 1.0]], dtype=np.float64)


def poses(frame_ids: np.ndarray) -> np.ndarray:
    """Camera-to-world 4x4 float64 poses for the given frame ids ([F,4,4]).

    Camera axes: x right, y down, z forward (the convention create_pcd assumes,
    generic.py:122-124).  World: x, y(up), z of the room.
    """
    f = np.asarray(frame_ids, dtype=np.float64)
    t = f * 0.013
    cx, cz = ROOM[0] / 2, ROOM[2] / 2
    px = cx + 3.4 * np.cos(t)
    pz = cz + 2.3 * np.sin(t)
    py = 1.5 + 0.12 * np.sin(0.37 * f)
    yaw = 0.11 * f
    pitch = 0.15 * np.sin(0.05 * f)
    cy_, sy_ = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    # camera basis in world coordinates
    fwd = np.stack([cy_ * cp, -sp, sy_ * cp], -1)          # z_cam
    right = np.stack([-sy_, np.zeros_like(yaw), cy_], -1)   # x_cam
    down = np.cross(fwd, right)                             # y_cam (points down)
    T = np.zeros((len(f), 4, 4), dtype=np.float64)
    T[:, :3, 0] = right
    T[:, :3, 1] = down
    T[:, :3, 2] = fwd
    T[:, 0, 3] = px
    T[:, 1, 3] = py
    T[:, 2, 3] = pz
    T[:, 3, 3] = 1.0
    return T


def _hash_u32(a: torch.Tensor) -> torch.Tensor:
    """32-bit integer mix on int64 tensors (values kept in [0, 2^32))."""
    m = 0xFFFFFFFF
    a = a & m
    a = ((a ^ (a >> 16)) * 0x45D9F3B) & m
    a = ((a ^ (a >> 16)) * 0x45D9F3B) & m
    a = (a ^ (a >> 16)) & m
    return a


@torch.no_grad()
def make_frames(frame_ids, H: int, W: int, depth_cut: float = 10.0, device="cpu",
                hole_frac: float = 0.05, return_labels: bool = False):
    """Returns (depth uint16 [F,H,W] as int16-viewed torch.uint16, rgb uint8 [F,H,W,3],
    poses float64 [F,4,4] numpy, K float64 [3,3] numpy) and, with return_labels, the surface every pixel sees as
    int8 [F,H,W]: 0..5 the six room planes, 6..11 the boxes, -1 where depth is 0 (SAM-like instance masks for the
    object layer: the same surface keeps its label across frames)."""
    frame_ids = np.asarray(frame_ids, dtype=np.int64)
    F = len(frame_ids)
    K = intrinsics(H, W)
    T = poses(frame_ids)
    dev = torch.device(device)
    Tt = torch.from_numpy(T).to(dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float64),
                            torch.arange(W, device=dev, dtype=torch.float64), indexing="ij")
    dcam = torch.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], torch.ones_like(xs)], -1)  # [H,W,3]
    depth_out = torch.empty((F, H, W), dtype=torch.int32, device=dev)
    label_out = torch.empty((F, H, W), dtype=torch.int8, device=dev) if return_labels else None
    rgb_out = torch.empty((F, H, W, 3), dtype=torch.uint8, device=dev)
    pix = (torch.arange(H, device=dev).view(H, 1) * W + torch.arange(W, device=dev).view(1, W)).to(torch.int64)
    lo_room = torch.zeros(3, dtype=torch.float64, device=dev)
    hi_room = torch.tensor(ROOM, dtype=torch.float64, device=dev)
    boxes = torch.tensor(BOXES, dtype=torch.float64, device=dev)
    for n in range(F):
        R = Tt[n, :3, :3]
        o = Tt[n, :3, 3]
        d = dcam @ R.T                              # world ray dir per unit Z-depth
        inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
        # room: camera is inside -> exit distance
        t1 = (lo_room - o) * inv
        t2 = (hi_room - o) * inv
        ex = torch.maximum(t1, t2).min(-1)
        s = ex.values
        if return_labels:
            lab = (ex.indices * 2 + (torch.gather(t2, -1, ex.indices[..., None])[..., 0] >= torch.gather(t1, -1, ex.indices[..., None])[..., 0]).long())
        for b in range(boxes.shape[0]):
            ta = (boxes[b, :3] - o) * inv
            tb = (boxes[b, 3:] - o) * inv
            tn = torch.minimum(ta, tb).max(-1).values
            tf = torch.maximum(ta, tb).min(-1).values
            hit = (tn < tf) & (tn > 1e-6)
            if return_labels:
                lab = torch.where(hit & (tn < s), torch.full_like(lab, 6 + b), lab)
            s = torch.where(hit & (tn < s), tn, s)
        mm = torch.round(s * 1000.0)
        mm = torch.where((s > depth_cut) | (mm > 65535) | (mm < 1), torch.zeros_like(mm), mm)
        fid = int(frame_ids[n])
        h = _hash_u32(pix * 2654435761 + fid * 97 + 12345)
        hole = (h % 10000) < int(hole_frac * 10000)
        mm = torch.where(hole, torch.zeros_like(mm), mm)
        depth_out[n] = mm.to(torch.int32)
        if return_labels:
            label_out[n] = torch.where(mm > 0, lab, torch.full_like(lab, -1)).to(torch.int8)
        h2 = _hash_u32(h + 0x9E3779B9)
        rgb_out[n, :, :, 0] = (h2 & 255).to(torch.uint8)
        rgb_out[n, :, :, 1] = ((h2 >> 8) & 255).to(torch.uint8)
        rgb_out[n, :, :, 2] = ((h2 >> 16) & 255).to(torch.uint8)
    depth_u16 = depth_out.to(torch.uint16) if hasattr(torch, "uint16") else depth_out
    if return_labels:
        return depth_u16, rgb_out, T, K, label_out
    return depth_u16, rgb_out, T, K


N_INSTANCES = 12


@torch.no_grad()
def instance_masks(labels: torch.Tensor):
    """labels int8 [F,H,W] (make_frames(return_labels=True)) -> (seg uint8 [F,12,H,W], xywh int32 [F,12,4]): one dense
    mask per visible surface (what SAM's "segmentation" / "bbox" carry); an invisible surface is an empty mask on a
    1-pixel box."""
    F, H, W = labels.shape
    ids = torch.arange(N_INSTANCES, device=labels.device, dtype=labels.dtype).view(1, -1, 1, 1)
    seg = (labels[:, None] == ids)
    rows = seg.any(-1); cols = seg.any(-2)                                  # [F,12,H], [F,12,W]
    ar_h = torch.arange(H, device=labels.device); ar_w = torch.arange(W, device=labels.device)
    y0 = torch.where(rows, ar_h, H).amin(-1); y1 = torch.where(rows, ar_h, -1).amax(-1)
    x0 = torch.where(cols, ar_w, W).amin(-1); x1 = torch.where(cols, ar_w, -1).amax(-1)
    empty = y1 < 0
    xywh = torch.stack([torch.where(empty, 0, x0), torch.where(empty, 0, y0), torch.where(empty, 1, x1 - x0 + 1), torch.where(empty, 1, y1 - y0 + 1)], -1)
    return seg.to(torch.uint8), xywh.to(torch.int32)


def make_frames_np(frame_ids, H, W, depth_cut=10.0):
    """CPU/numpy convenience wrapper: depth uint16 [F,H,W], rgb uint8 [F,H,W,3], poses, K."""
    d, c, T, K = make_frames(frame_ids, H, W, depth_cut, device="cpu")
    return d.view(torch.int16).numpy().view(np.uint16).copy(), c.numpy().copy(), T, K


def make_masks(frame_id: int, depth: np.ndarray, M: int = 32):
    """M seeded rectangular SAM-like mask dicts for one frame.

    Each dict carries the keys the reference reads: "segmentation" bool [H,W]
    (extractor.py:184; generic.py:169; sam_utils.py:159), "bbox" XYWH
    (sam_utils.py:143,158) and "predicted_iou" (extractor.py:124).
    segmentation = rectangle AND (depth > 0).
    """
    H, W = depth.shape
    bb = make_mask_boxes(frame_id, H, W, M)
    valid = depth > 0
    out = []
    for (x, y, w, h) in bb:
        seg = np.zeros((H, W), dtype=bool)
        seg[y:y + h, x:x + w] = valid[y:y + h, x:x + w]
        out.append({"segmentation": seg, "bbox": [int(x), int(y), int(w), int(h)], "predicted_iou": 0.9})
    return out


def make_mask_boxes(frame_id: int, H: int, W: int, M: int = 32) -> np.ndarray:
    """int32 [M,4] XYWH rectangles, seed = 1234 + frame_id."""
    rs = np.random.RandomState(1234 + int(frame_id))
    w = np.minimum(rs.randint(32, 257, size=M), W)
    h = np.minimum(rs.randint(32, 257, size=M), H)
    x = (rs.rand(M) * (W - w + 1)).astype(np.int64)
    y = (rs.rand(M) * (H - h + 1)).astype(np.int64)
    return np.stack([x, y, w, h], 1).astype(np.int32)


def make_knn_tables(N: int, Q: int, d: int = 512, device="cpu", scale: float = 0.8):
    """E [N,d] unit rows (seed 7) and Q [Q,d] unit rows * 0.8 (seed 11); SURVEY §8d."""
    g = torch.Generator(device=device)
    g.manual_seed(7)
    E = torch.empty((N, d), dtype=torch.float32, device=device)
    step = 1 << 18
    for s in range(0, N, step):
        e = torch.randn((min(step, N - s), d), generator=g, device=device, dtype=torch.float32)
        E[s:s + e.shape[0]] = e / e.norm(dim=-1, keepdim=True)
    g.manual_seed(11)
    q = torch.randn((Q, d), generator=g, device=device, dtype=torch.float32)
    q = q / q.norm(dim=-1, keepdim=True) * scale
    return E, q


@dataclass
class VitB32Shape:
    """open_clip ViT-B/32 visual tower (graph.py:112-119; constants.py:3-7 => d=512)."""
    image: int = 224
    patch: int = 32
    width: int = 768
    layers: int = 12
    heads: int = 12
    mlp: int = 3072
    out_dim: int = 512

    @property
    def tokens(self):
        return (self.image // self.patch) ** 2 + 1


# the reference's default `models.clip.type: ViT-L/14` (graph.py:98-104, clip_feat_dim 768): 257 tokens, width 1024
VIT_L14 = VitB32Shape(image=224, patch=14, width=1024, layers=24, heads=16, mlp=4096, out_dim=768)


def make_vit_weights(shape: VitB32Shape = VitB32Shape(), seed: int = 0) -> dict:
    """Seeded N(0,0.02) weights rounded to fp16 (stored as float32 of the rounded
    values), named like open_clip's ``VisionTransformer.state_dict()``."""
    g = torch.Generator().manual_seed(seed)
    w = shape.width

    def rn(*s, std=0.02, mean=0.0):
        return (torch.randn(*s, generator=g) * std + mean).half().float()

    sd = {
        "conv1.weight": rn(w, 3, shape.patch, shape.patch),
        "class_embedding": rn(w),
        "positional_embedding": rn(shape.tokens, w),
        "ln_pre.weight": rn(w, mean=1.0), "ln_pre.bias": rn(w),
        "ln_post.weight": rn(w, mean=1.0), "ln_post.bias": rn(w),
        "proj": rn(w, shape.out_dim, std=w ** -0.5),
    }
    for i in range(shape.layers):
        p = f"transformer.resblocks.{i}."
        sd[p + "ln_1.weight"] = rn(w, mean=1.0)
        sd[p + "ln_1.bias"] = rn(w)
        sd[p + "attn.in_proj_weight"] = rn(3 * w, w, std=w ** -0.5)
        sd[p + "attn.in_proj_bias"] = rn(3 * w)
        sd[p + "attn.out_proj.weight"] = rn(w, w, std=w ** -0.5)
        sd[p + "attn.out_proj.bias"] = rn(w)
        sd[p + "ln_2.weight"] = rn(w, mean=1.0)
        sd[p + "ln_2.bias"] = rn(w)
        sd[p + "mlp.c_fc.weight"] = rn(shape.mlp, w, std=w ** -0.5)
        sd[p + "mlp.c_fc.bias"] = rn(shape.mlp)
        sd[p + "mlp.c_proj.weight"] = rn(w, shape.mlp, std=shape.mlp ** -0.5)
        sd[p + "mlp.c_proj.bias"] = rn(w)
    return sd
