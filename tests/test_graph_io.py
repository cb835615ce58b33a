"""CPU: the artefact I/O half of the Graph mirror (graph.py:3769-3990) and the PLY subset it falls back to when
open3d is not importable (binary_little_endian double xyz + uchar rgb = what Open3D writes for a PointCloud)."""
import os

import numpy as np
import pytest
import torch

from holoagent_b200.memory.hmsg.graph.graph import Graph
from holoagent_b200.runtime import PointCloud, read_point_cloud, write_point_cloud


class _NoEngine:
    """Graph only stores the engine in its constructor; nothing in this file may touch the GPU."""
    def __getattr__(self, name):
        raise AssertionError(f"I/O path touched the engine ({name})")


def _graph():
    return Graph({"pipeline": {}}, engine=_NoEngine(), clip_feat_dim=512)


def _cloud(rs, n):
    return PointCloud(rs.randn(n, 3) * 3.0, rs.rand(n, 3))


def test_ply_roundtrip(tmp_path):
    rs = np.random.RandomState(0)
    pc = _cloud(rs, 1000)
    fn = os.path.join(tmp_path, "a.ply")
    write_point_cloud(fn, pc)
    head = open(fn, "rb").read(200)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0") and b"property double x" in head and b"property uchar red" in head
    back = read_point_cloud(fn)
    assert np.array_equal(np.asarray(back.points), pc.points)                      # float64 positions survive exactly
    assert np.array_equal(np.asarray(back.colors), np.round(pc.colors * 255.0) / 255.0)   # colours are 8-bit on disk
    # empty cloud and a cloud without colours
    write_point_cloud(fn, PointCloud())
    assert len(np.asarray(read_point_cloud(fn).points)) == 0
    write_point_cloud(fn, PointCloud(pc.points[:5]))
    b = read_point_cloud(fn)
    assert np.array_equal(np.asarray(b.points), pc.points[:5]) and len(np.asarray(b.colors)) == 0


def test_ply_reader_accepts_float32_and_ascii(tmp_path):
    pts = np.array([[0.5, -1.25, 2.0], [3.0, 4.0, -5.5]], np.float32)
    fn = os.path.join(tmp_path, "f.ply")
    rec = np.zeros(2, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1"), ("alpha", "u1")])
    rec["x"], rec["y"], rec["z"] = pts.T
    rec["red"], rec["green"], rec["blue"] = [255, 0], [128, 64], [0, 255]
    with open(fn, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                b"property float nx\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar alpha\n"
                b"element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        f.write(rec.tobytes())
    pc = read_point_cloud(fn)
    assert np.array_equal(np.asarray(pc.points), pts.astype(np.float64))
    assert np.allclose(np.asarray(pc.colors), [[1.0, 128 / 255, 0.0], [0.0, 64 / 255, 1.0]])
    with open(fn, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 2\nproperty double x\nproperty double y\nproperty double z\nend_header\n0.5 -1.25 2\n3 4 -5.5\n")
    assert np.array_equal(np.asarray(read_point_cloud(fn).points), pts.astype(np.float64))


def test_full_pcd_and_feats_files(tmp_path):
    rs = np.random.RandomState(1)
    g = _graph()
    g.full_pcd = _cloud(rs, 300)
    g.full_feats_array = rs.randn(300, 512).astype(np.float32)
    g.mask_pcds = [_cloud(rs, 40), PointCloud(), _cloud(rs, 25)]
    g.mask_feats = [rs.randn(512).astype(np.float32) for _ in range(3)]
    p = str(tmp_path)
    g.save_full_pcd(p)
    g.save_full_pcd_feats(p)
    assert len(g.mask_pcds) == 2 and g.mask_feats.shape == (2, 512)               # the empty object is dropped (graph.py:3806-3812)
    assert sorted(os.listdir(p)) == ["full_feats.pt", "full_pcd.ply", "mask_feats.pt"]
    h = _graph()
    assert np.array_equal(np.asarray(h.load_full_pcd(p).points), g.full_pcd.points)
    ff = h.load_full_pcd_feats(p, full_feats=True, normalize=False)
    assert np.array_equal(ff, g.full_feats_array) and h.full_feats_array is ff
    mf = h.load_full_pcd_feats(p)                                                 # defaults: mask feats, L2-normalised
    ref = torch.nn.functional.normalize(torch.from_numpy(g.mask_feats), p=2, dim=-1).numpy()
    assert np.array_equal(mf, ref) and np.allclose(np.linalg.norm(h.mask_feats, axis=1), 1.0, atol=1e-6)
    assert h.load_full_pcd(os.path.join(p, "missing")) is None and h.load_full_pcd_feats(os.path.join(p, "missing")) is None


@pytest.mark.parametrize("state", ["both", "objects", "full"])
def test_masked_pcds_files(tmp_path, state):
    rs = np.random.RandomState(2)
    g = _graph()
    g.mask_pcds = [_cloud(rs, 30), _cloud(rs, 4), _cloud(rs, 12), PointCloud()]
    g.mask_feats = [rs.randn(512).astype(np.float32) for _ in range(4)]
    kept_pts = [g.mask_pcds[0].points.copy(), g.mask_pcds[2].points.copy()]
    p = str(tmp_path)
    g.save_masked_pcds(p, state)
    assert len(g.mask_pcds) == 2 and len(g.mask_feats) == 2                        # < 10 points and empty objects removed
    if state in ("both", "objects"):
        assert sorted(os.listdir(os.path.join(p, "objects"))) == ["pcd_0.ply", "pcd_1.ply"]
    if state in ("both", "full"):
        u = read_point_cloud(os.path.join(p, "masked_pcd.ply"))
        assert np.array_equal(np.asarray(u.points), np.concatenate(kept_pts))
        assert len(np.unique(np.asarray(u.colors), axis=0)) == 2                   # one random colour per object
    if state == "full":
        return
    h = _graph()
    assert h.load_masked_pcds_new(p) is None                                       # "load full pcd feats first"
    h.mask_feats = np.stack(g.mask_feats)
    os.remove(os.path.join(p, "objects", "pcd_0.ply"))
    open(os.path.join(p, "objects", "note.txt"), "w").write("x")                   # listdir count = 2, pcd_0 missing
    out = h.load_masked_pcds_new(p)
    assert len(out) == 1 and np.array_equal(np.asarray(out[0].points), kept_pts[1])
    assert h.mask_feats.shape == (1, 512) and np.array_equal(h.mask_feats[0], g.mask_feats[1])
