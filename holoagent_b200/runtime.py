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


# ---------------------------------------------------------------------------------------------
# .ply point clouds (graph.py:3769-3795, :3880-3990 go through o3d.io.{write,read}_point_cloud).
# open3d is used when it is importable; otherwise the same file subset is written / read here:
# binary_little_endian 1.0, `double x y z` + `uchar red green blue` (what Open3D 0.18 writes for a
# coloured PointCloud), and float / double positions, optional normals / alpha, ascii or binary on read.
# ---------------------------------------------------------------------------------------------
_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8",
              "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def paint_uniform_color(pc, rgb):
    """PointCloud.paint_uniform_color for both the open3d object and the stand-in."""
    import numpy as np
    if hasattr(pc, "paint_uniform_color"):
        pc.paint_uniform_color(rgb)
    else:
        pc.colors = np.tile(np.asarray(rgb, dtype=np.float64).reshape(1, 3), (len(pc.points), 1))
    return pc


def write_point_cloud(path, pc):
    import numpy as np
    try:
        import open3d as o3d
        if isinstance(pc, o3d.geometry.PointCloud):
            return o3d.io.write_point_cloud(path, pc)
    except ImportError:
        pass
    pts = np.asarray(pc.points, dtype=np.float64).reshape(-1, 3)
    cols = np.asarray(pc.colors, dtype=np.float64).reshape(-1, 3)
    has_c = len(cols) == len(pts) and len(pts) > 0
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")] + ([("red", "u1"), ("green", "u1"), ("blue", "u1")] if has_c else [])
    rec = np.zeros(len(pts), dtype=fields)
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    if has_c:
        c8 = np.round(np.clip(cols * 255.0, 0.0, 255.0)).astype(np.uint8)      # Open3D: round(min(255, max(0, c * 255)))
        rec["red"], rec["green"], rec["blue"] = c8[:, 0], c8[:, 1], c8[:, 2]
    head = ["ply", "format binary_little_endian 1.0", "comment Created by holoagent_b200", f"element vertex {len(pts)}",
            "property double x", "property double y", "property double z"]
    if has_c:
        head += ["property uchar red", "property uchar green", "property uchar blue"]
    head.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        f.write(rec.tobytes())
    return True


def read_point_cloud(path):
    import numpy as np
    try:
        import open3d as o3d
        return o3d.io.read_point_cloud(path)
    except ImportError:
        pass
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            a = np.loadtxt(f, max_rows=n, ndmin=2) if n else np.zeros((0, len(props)))
            col = {name: a[:, i] for i, (name, _) in enumerate(props)}
        elif fmt in ("binary_little_endian", "binary_big_endian"):
            e = "<" if fmt == "binary_little_endian" else ">"
            rec = np.frombuffer(f.read(n * sum(np.dtype(t).itemsize for _, t in props)), dtype=[(nm, e + t) for nm, t in props], count=n)
            col = {nm: rec[nm] for nm, _ in props}
        else:
            raise ValueError(f"{path}: unknown PLY format {fmt}")
    pts = np.stack([col["x"], col["y"], col["z"]], axis=1).astype(np.float64) if n else np.zeros((0, 3))
    cols = np.zeros((0, 3))
    if n and all(k in col for k in ("red", "green", "blue")):
        cols = np.stack([col["red"], col["green"], col["blue"]], axis=1).astype(np.float64) / 255.0
    return PointCloud(pts, cols)
