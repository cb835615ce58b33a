"""GPU parity: N1 3-D mask merging (seq_merge / merge_3d_masks, graph_utils.py:620-679, :827-956,
:1015-1038) through the C-ABI vs the fixture written by the reference's own run and vs the oracle."""
import os

import numpy as np
import pytest

from holoagent_b200 import synth
from oracle import hmsg_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ragged(masks):
    off = np.zeros(len(masks) + 1, np.int64)
    for i, (p, _) in enumerate(masks):
        off[i + 1] = off[i] + len(p)
    xyz = np.concatenate([p for p, _ in masks], 0) if len(masks) and off[-1] else np.zeros((0, 3))
    rgb = np.concatenate([c for _, c in masks], 0) if len(masks) and off[-1] else np.zeros((0, 3))
    return off, xyz, rgb


def _run_gpu(engine, frames_masks, th, down, iou, min_points):
    engine.objects_begin(th, down, iou)
    for fm in frames_masks:
        engine.objects_add_masks(*_ragged(fm))
    engine.objects_finish(min_points)
    return engine.objects_read()


def test_seq_merge_matches_reference_run(engine):
    """Per-frame 3-D masks of the fixture scene (oracle A7, pinned by the same reference run) ->
    CUDA seq_merge == the object point sets Graph.create_feature_map produced."""
    z = np.load(os.path.join(GOLD, "ref_build.npz"))
    H, W = int(z["H"]), int(z["W"])
    depth, rgb, T, K = synth.make_frames_np(z["frame_ids"], H, W)
    segs = np.unpackbits(z["segs"], axis=-1)[..., :W].astype(bool)
    tree = O.build_tree(z["node_xyz"])
    frames_masks = [[(p, c) for p, c, _ in O.create_3d_masks(segs[f], depth[f], K, 1000.0, T[f], z["node_xyz"], z["node_rgb"], tree, float(z["voxel_size"]))]
                    for f in range(len(depth))]
    off, xyz, col = _run_gpu(engine, frames_masks, 0.75, float(z["voxel_size"]), 0.05, 10)
    assert np.array_equal(off, z["obj_off"])
    assert np.array_equal(xyz, z["obj_pts"])
    # colours ride along with the points: check against the oracle's merge
    objs = [o for o in O.seq_merge(frames_masks, 0.75, float(z["voxel_size"]), 0.05) if len(o[0]) >= 10]
    assert np.array_equal(col, np.concatenate([o[1] for o in objs], 0))


def _blob_scene(seed, n_frames, n_obj, masks_per_frame):
    """Synthetic instance masks: each frame sees a random subset of blobs, each through a random
    sub-window of the blob's voxel-grid points (so masks of one blob overlap partially), plus clutter."""
    rs = np.random.RandomState(seed)
    vs = 0.05
    blobs = []
    for _ in range(n_obj):
        c = rs.rand(3) * [6, 6, 1.5]
        ext = 0.2 + rs.rand(3) * 0.5
        g = np.stack(np.meshgrid(*[np.arange(-e, e, vs) for e in ext], indexing="ij"), -1).reshape(-1, 3)
        g = g[(np.abs(g / ext) ** 2).sum(1) < 1.0] + c + rs.rand(3) * 0.01
        blobs.append(g)
    frames = []
    for _ in range(n_frames):
        fm = []
        for b in rs.choice(n_obj, size=masks_per_frame, replace=True):
            g = blobs[b]
            if rs.rand() < 0.15:
                fm.append((np.zeros((0, 3)), np.zeros((0, 3))))                 # an empty mask (filtered / no depth)
                continue
            lo = g.min(0) + (g.max(0) - g.min(0)) * rs.rand(3) * 0.3
            hi = g.max(0) - (g.max(0) - g.min(0)) * rs.rand(3) * 0.3
            sel = g[np.all((g >= lo) & (g <= hi), 1)]
            if rs.rand() < 0.3 and len(sel):                                    # stray far-away points (DBSCAN removes them)
                sel = np.concatenate([sel, sel[:3] + [1.5, 0, 0]])
            fm.append((sel, rs.rand(len(sel), 3)))
        frames.append(fm)
    return frames, vs


@pytest.mark.parametrize("seed,n_frames,n_obj,mpf", [(0, 5, 6, 4), (1, 8, 10, 6), (2, 3, 3, 8), (3, 1, 4, 5)])
def test_seq_merge_vs_oracle(engine, seed, n_frames, n_obj, mpf):
    frames, vs = _blob_scene(seed, n_frames, n_obj, mpf)
    ref = [o for o in O.seq_merge(frames, 0.75, vs, 0.05) if len(o[0]) >= 10]
    off, xyz, col = _run_gpu(engine, frames, 0.75, vs, 0.05, 10)
    assert len(off) - 1 == len(ref)
    for i, (p, c) in enumerate(ref):
        assert np.array_equal(xyz[off[i]:off[i + 1]], p), i
        assert np.array_equal(col[off[i]:off[i + 1]], c), i


def test_objects_edge_cases(engine):
    engine.objects_begin(0.75, 0.05, 0.05)
    assert engine.objects_finish(10) == (0, 0)                                # no frames at all
    engine.objects_begin(0.75, 0.05, 0.05)
    engine.objects_add_masks(np.zeros(1, np.int64), np.zeros((0, 3)))         # a frame without masks
    engine.objects_add_masks(np.array([0, 0, 0], np.int64), np.zeros((0, 3)))  # two empty masks
    assert engine.objects_finish(10) == (0, 0)
