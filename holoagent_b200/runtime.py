"""Process-wide default engine (one hmsg_ctx per GPU) for the drop-in modules."""
from __future__ import annotations

_engines = {}


def get_engine(device: int = 0):
    from .engine import HmsgEngine
    if device not in _engines:
        _engines[device] = HmsgEngine(device)
    return _engines[device]


class PointCloud:
    """Minimal stand-in for open3d.geometry.PointCloud (points/colors as float64 arrays) used
    when open3d is not importable.  Carries an optional token tying it to an engine node table."""

    def __init__(self, points=None, colors=None, token=None):
        import numpy as np
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)
        self.colors = np.zeros((0, 3)) if colors is None else np.asarray(colors, dtype=np.float64)
        self._hmsg_token = token

    def is_empty(self):
        return len(self.points) == 0

    def __len__(self):
        return len(self.points)

    def __iadd__(self, other):
        import numpy as np
        self.points = np.concatenate([self.points, other.points]); self.colors = np.concatenate([self.colors, other.colors])
        self._hmsg_token = None
        return self


def to_o3d(pc: PointCloud):
    """open3d PointCloud if open3d is installed, else the stand-in itself."""
    try:
        import open3d as o3d
    except Exception:
        return pc
    out = o3d.geometry.PointCloud()
    out.points = o3d.utility.Vector3dVector(pc.points)
    if len(pc.colors):
        out.colors = o3d.utility.Vector3dVector(pc.colors)
    return out
