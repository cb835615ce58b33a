"""Import harness that lets the UNMODIFIED reference sources under /root/reference/fsr_vln run in this
container (no open3d / faiss / open_clip / segment_anything / omegaconf / GPU here).

Used ONLY by tests/golden/make_reference_golden.py to generate fixtures from the reference's own
code.  Nothing under holoagent_b200/ imports it, and it cannot run on the GPU box (/root/reference is
absent there) - the fixtures it produced are committed instead.

Three kinds of stand-ins, kept strictly apart:
  * absent modules nothing on the executed path calls (matplotlib, pyvista, oss2, ...) -> inert dummies;
  * `open3d` / `faiss` -> functional shims for the handful of methods the executed path calls.  Their
    arithmetic is the oracle's restatement of Open3D 0.18 / faiss IndexFlatL2 (cited there): fixtures
    made through them pin the REFERENCE GLUE (loops, index_put, counters, dtype casts, cv2 / PIL /
    torch / scipy / sklearn calls, thresholds, ordering) - not Open3D's internals;
  * `.cuda()` / device="cuda" -> no-ops (placement only, no arithmetic).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/fsr_vln"
INERT = {"matplotlib", "omegaconf", "open_clip", "segment_anything", "torchmetrics", "oss2", "pyvista", "skfmm"}


class _Meta(type):
    def __getattr__(cls, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy

    def __getitem__(cls, k):
        return _Dummy


class _Dummy(metaclass=_Meta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in INERT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _InertModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ---------------------------------------------------------------------------------------------
# open3d shim
# ---------------------------------------------------------------------------------------------
def _lazy(n):
    if n.startswith("__"):
        raise AttributeError(n)
    return _Dummy


def _oracle():
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from oracle import hmsg_oracle as O
    return O


class _AABB:
    def __init__(self, lo, hi):
        self._lo, self._hi = lo, hi

    def get_min_bound(self):
        return self._lo

    def get_max_bound(self):
        return self._hi


class PointCloud:
    """open3d.geometry.PointCloud stand-in: float64 [n,3] points / colors."""

    def __init__(self, points=None, colors=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64).reshape(-1, 3)
        self.colors = np.zeros((0, 3)) if colors is None else np.asarray(colors, dtype=np.float64).reshape(-1, 3)

    def __setattr__(self, k, v):
        if k in ("points", "colors"):
            v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
        object.__setattr__(self, k, v)

    def _has_colors(self):
        return len(self.colors) == len(self.points) and len(self.points) > 0

    def __iadd__(self, o):                       # PointCloud::operator+= : append; colours kept only if both have them
        if len(o.points) == 0:
            return self
        keep_c = (len(self.points) == 0 or self._has_colors()) and o._has_colors()
        old_c = self.colors if len(self.points) else np.zeros((0, 3))
        self.points = np.concatenate([self.points, o.points], 0)
        self.colors = np.concatenate([old_c, o.colors], 0) if keep_c else np.zeros((0, 3))
        return self

    def __add__(self, o):
        r = PointCloud(self.points.copy(), self.colors.copy())
        r += o
        return r

    def is_empty(self):
        return len(self.points) == 0

    def has_points(self):
        return len(self.points) > 0

    def transform(self, T):
        self.points = _oracle().transform_points(self.points, T)
        return self

    def voxel_down_sample(self, voxel_size):
        O = _oracle()
        p, c, _, _ = O.voxel_down_sample(self.points, self.colors if self._has_colors() else None, voxel_size)
        return PointCloud(p, c)

    def remove_radius_outlier(self, nb_points, radius):
        ind = _oracle().radius_outlier_keep(self.points, nb_points, radius)
        return self.select_by_index(ind), [int(i) for i in ind]

    def select_by_index(self, ind):
        ind = np.asarray(ind, dtype=np.int64)
        return PointCloud(self.points[ind], self.colors[ind] if self._has_colors() else None)

    def cluster_dbscan(self, eps, min_points, print_progress=False):
        return [int(v) for v in _oracle().cluster_dbscan(self.points, eps, min_points)]

    def get_min_bound(self):
        return self.points.min(axis=0) if len(self.points) else np.zeros(3)

    def get_max_bound(self):
        return self.points.max(axis=0) if len(self.points) else np.zeros(3)

    def get_axis_aligned_bounding_box(self):
        return _AABB(self.get_min_bound(), self.get_max_bound())


def _make_open3d():
    o3d = types.ModuleType("open3d")
    geometry = types.ModuleType("open3d.geometry")
    geometry.PointCloud = PointCloud
    geometry.AxisAlignedBoundingBox = _AABB
    geometry.__getattr__ = _lazy
    utility = types.ModuleType("open3d.utility")
    utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64).reshape(-1, 3)
    utility.__getattr__ = _lazy
    io = types.ModuleType("open3d.io")
    io.write_point_cloud = lambda *a, **k: True
    io.__getattr__ = _lazy
    o3d.geometry, o3d.utility, o3d.io = geometry, utility, io
    o3d.__getattr__ = _lazy
    o3d.__path__ = []
    for m in (o3d, geometry, utility, io):
        sys.modules[m.__name__] = m
    for sub in ("visualization", "pipelines", "core", "t"):
        sm = _InertModule("open3d." + sub)
        sm.__path__ = []
        sys.modules[sm.__name__] = sm
        setattr(o3d, sub, sm)


# ---------------------------------------------------------------------------------------------
# faiss shim (IndexFlatL2.add / search k=1 only)
# ---------------------------------------------------------------------------------------------
class IndexFlatL2:
    def __init__(self, d):
        self.d = d
        self.x = np.zeros((0, d), np.float32)

    def add(self, x):
        self.x = np.concatenate([self.x, np.ascontiguousarray(x, np.float32)], 0)

    def search(self, q, k=1):
        D, I = _oracle().flat_l2_nn(np.ascontiguousarray(q, np.float32), self.x)
        return D.reshape(-1, 1), I.reshape(-1, 1)


def _make_faiss():
    f = types.ModuleType("faiss")
    f.IndexFlatL2 = IndexFlatL2
    f.__getattr__ = _lazy
    sys.modules["faiss"] = f


def install():
    """Make `import memory.hmsg...` / `import perception...` resolve to the reference sources."""
    import torch
    if not os.path.isdir(REF):
        raise RuntimeError("reference sources not present (this harness only runs in the build container)")
    sys.meta_path.insert(0, _Finder())
    _make_open3d()
    _make_faiss()
    sys.path.insert(0, REF)
    # device placement no-ops (no GPU in the container)
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "ones", "empty", "tensor"):
        orig = getattr(torch, name)

        def wrap(*a, __orig=orig, **k):
            k.pop("device", None) if k.get("device") == "cuda" else None
            return __orig(*a, **k)
        setattr(torch, name, wrap)
