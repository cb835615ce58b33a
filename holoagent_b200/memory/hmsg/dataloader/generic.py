"""Drop-in for fsr_vln/memory/hmsg/dataloader/generic.py (RGBDDataset.create_pcd /
create_3d_masks, generic.py:74-190) backed by libhmsg_b200.so.  Same argument names, defaults
and return types; point clouds are open3d objects when open3d is installed, otherwise the
minimal holoagent_b200.runtime.PointCloud (``.points`` / ``.colors`` float64 arrays)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from holoagent_b200.runtime import PointCloud, get_engine


class RGBDDataset(ABC):
    def __init__(self, cfg):
        self.root_dir = cfg["root_dir"]
        self.transforms = cfg["transforms"]
        self.depth_cut = cfg["depth_cut"]
        self.rgb_intrinsics = None
        self.depth_intrinsics = None
        self.scale = None
        self.data_list = self._get_data_list()
        self._engine = None

    @abstractmethod
    def _get_data_list(self):
        pass

    def __len__(self):
        return len(self.data_list)

    # -- engine plumbing -----------------------------------------------------------------
    def _scratch_engine(self, H, W):
        """A private 1-frame scene for stand-alone create_pcd calls."""
        from holoagent_b200.engine import HmsgEngine
        if self._engine is None or (self._engine.H, self._engine.W) != (H, W):
            self._engine = HmsgEngine(0)
            self._engine.scene_begin(H, W, np.asarray(self.depth_intrinsics, dtype=np.float64), float(self.scale), 0.05, 1)
        return self._engine

    def create_pcd(self, rgb, depth, camera_pose=None, idx=None, mask_img=False, filter_distance=np.inf):
        """generic.py:74-138.  rgb: image (or bool mask when mask_img), depth: uint16 image."""
        import cv2
        rgb = np.array(rgb)
        depth = np.array(depth)
        if mask_img:
            depth = (depth * rgb.astype(bool)).astype(np.uint16)        # :115-116 depth * mask
            rgb = np.zeros(depth.shape + (3,), np.uint8)
        else:
            rgb = rgb.astype(np.uint8)
            if rgb.shape[0] != depth.shape[0] or rgb.shape[1] != depth.shape[1]:
                rgb = cv2.resize(rgb, (depth.shape[1], depth.shape[0]), interpolation=cv2.INTER_AREA)   # :98-104
        H, W = depth.shape
        eng = self._scratch_engine(H, W)
        eng.scene_reset_frames()
        pose = np.eye(4) if camera_pose is None else np.asarray(camera_pose, dtype=np.float64)
        eng.add_frames(depth.astype(np.uint16)[None], rgb[None], pose.reshape(1, 16))
        xyz, col, valid = eng.unproject_frame(0)
        pts = xyz[valid]
        if len(pts):
            # :126-127 mean camera-frame depth filter (Z of the float32-rounded depth)
            z = (depth.astype(np.float32) / self.scale)[valid.reshape(H, W)]
            if z.mean() > filter_distance:
                return _wrap(PointCloud())
        return _wrap(PointCloud(pts, None if mask_img else col[valid]))

    def create_3d_masks(self, masks, depth, full_pcd, full_pcd_tree, camera_pose, idx=None, down_size=0.02, filter_distance=None):
        """generic.py:140-190.  ``full_pcd`` must be the node table of a holoagent_b200 Graph
        (the engine already holds it together with its spatial index; ``full_pcd_tree`` is
        accepted for signature compatibility and ignored)."""
        eng = getattr(full_pcd, "_hmsg_engine", None)
        if eng is None:
            raise NotImplementedError("create_3d_masks needs full_pcd produced by holoagent_b200 Graph.create_feature_map "
                                      "(its node index lives on the GPU); arbitrary clouds are not indexed")
        if idx is None:
            raise ValueError("create_3d_masks: pass the frame id `idx` (graph.py:397 does)")
        frame = full_pcd._hmsg_frame_of(idx)
        seg = np.stack([np.asarray(m["segmentation"]).astype(np.uint8) for m in masks])
        eng.masks_dense(frame, seg[None])
        off, xyz, rgb, _ = eng.mask_nodes(frame, down_size, len(masks))
        return [_wrap(PointCloud(xyz[off[i]:off[i + 1]], rgb[off[i]:off[i + 1]])) for i in range(len(masks))]


def _wrap(pc):
    from holoagent_b200.runtime import to_o3d
    return to_o3d(pc)
