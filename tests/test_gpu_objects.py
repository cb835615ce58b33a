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


# ------------------------------------------------------------------------------------------------
# fixture scene helpers (the scene of tests/golden/ref_build.npz)
# ------------------------------------------------------------------------------------------------
def _fixture_scene(engine):
    z = np.load(os.path.join(GOLD, "ref_build.npz"))
    H, W = int(z["H"]), int(z["W"])
    depth, rgb, T, K = synth.make_frames_np(z["frame_ids"], H, W)
    segs = np.unpackbits(z["segs"], axis=-1)[..., :W].astype(bool)
    vs = float(z["voxel_size"])
    engine.scene_begin(H, W, K, 1000.0, vs, len(depth))
    engine.add_frames(depth, rgb, T.reshape(len(depth), 16))
    engine.voxel_build()
    engine.radius_filter(1000, 1.0)
    nxyz, nrgb, _, _ = engine.nodes_read()
    assert nxyz.shape == z["node_xyz"].shape
    return z, depth, rgb, T, K, segs, vs, nxyz, nrgb


def test_add_frame_chained_on_device(engine):
    """create_3d_masks + seq_merge without leaving HBM == the oracle run on the GPU's own node table,
    bit for bit (voxel sums follow the row-major pixel order like Open3D's sequential accumulation)."""
    z, depth, rgb, T, K, segs, vs, nxyz, nrgb = _fixture_scene(engine)
    tree = O.build_tree(nxyz)
    frames_masks = []
    for f in range(len(depth)):
        pts, _, _ = O.create_pcd(rgb[f], depth[f], K, 1000.0, T[f])
        gidx, _ = engine.pixel_to_node(f, want_dist=False)                      # stage-wise: exact NN ties are implementation-defined
        full = np.asarray(gidx)
        fm = []
        for seg in segs[f]:
            sel = (seg & (depth[f] > 0)).reshape(-1)
            idx = full[sel]
            p, c, _, _ = O.voxel_down_sample(nxyz[idx], nrgb[idx], vs) if sel.any() else (np.zeros((0, 3)), np.zeros((0, 3)), None, None)
            fm.append((p, c))
        frames_masks.append(fm)
    ref = [o for o in O.seq_merge(frames_masks, 0.75, vs, 0.05) if len(o[0]) >= 10]
    engine.objects_begin(0.75, vs, 0.05)
    for f in range(len(depth)):
        engine.masks_dense(f, segs[f][None].astype(np.uint8))
        engine.objects_add_frame(f, vs, 6.0)
    engine.objects_finish(10)
    off, xyz, col = engine.objects_read()
    assert len(off) - 1 == len(ref)
    for i, (p, c) in enumerate(ref):
        assert np.array_equal(xyz[off[i]:off[i + 1]], p), i
        assert np.array_equal(col[off[i]:off[i + 1]], c), i
    # and within float64 noise of what the reference run itself produced (its node centroids were summed sequentially)
    assert np.array_equal(off, z["obj_off"]) and np.allclose(xyz, z["obj_pts"], rtol=1e-11, atol=1e-11)
    # filter_distance: every mask farther than 0.5 m on average is dropped -> nothing survives
    engine.objects_begin(0.75, vs, 0.05)
    for f in range(len(depth)):
        engine.masks_dense(f, segs[f][None].astype(np.uint8))
        engine.objects_add_frame(f, vs, 0.5)
    assert engine.objects_finish(10) == (0, 0)


def test_object_feats_vs_oracle(engine):
    """N2: voxel_down_sample -> NN gate -> nan_to_num -> cosine DBSCAN largest-cluster mean (graph.py:451-488)."""
    z, depth, rgb, T, K, segs, vs, nxyz, nrgb = _fixture_scene(engine)
    off = z["obj_off"]
    objs = [z["obj_pts"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    # extra objects: a tiny one (< min_points rows), one far from every node (all rows gated out), an empty one
    objs.append(nxyz[:40] + 1e-3)
    objs.append(nxyz[:30] + np.array([50.0, 0, 0]))
    engine.objects_begin(0.75, vs, 0.05)
    o2 = np.zeros(len(objs) + 1, np.int64)
    for i, p in enumerate(objs):
        o2[i + 1] = o2[i] + len(p)
    # a single frame whose masks are the objects; th = 2 > any ratio: nothing merges, DBSCAN(0.1, 10) may trim
    engine.objects_begin(2.0, vs, 0.05)
    engine.objects_add_masks(o2, np.concatenate(objs, 0))
    engine.objects_finish(1)
    goff, gxyz, _ = engine.objects_read()
    kept = [gxyz[goff[i]:goff[i + 1]] for i in range(len(goff) - 1)]
    n, d = len(nxyz), 256
    rs = np.random.RandomState(4)
    base = rs.randn(7, d).astype(np.float32); base /= np.linalg.norm(base, axis=1, keepdims=True)
    region = (np.floor(nxyz[:, 0] / 0.7).astype(np.int64) + 3 * np.floor(nxyz[:, 1] / 0.9).astype(np.int64)) % 7
    # per-region direction and scale + small noise: pairwise cosine distance inside a region <= 0.005 (clear of eps)
    full = base[region] * (0.5 + region[:, None].astype(np.float32) / 7.0) + 0.0015 * rs.randn(n, d).astype(np.float32)
    scatter = rs.rand(n) < 0.15                         # rows that belong to no cluster
    full[scatter] = rs.randn(int(scatter.sum()), d).astype(np.float32)
    full[rs.rand(n) < 0.02] = 0.0                       # nodes no pixel ever hit
    full[5, 3] = np.nan; full[9, 1] = np.inf
    full = full.astype(np.float32)
    got = engine.object_feats(full, vs, 0.8, 0.01, 100)
    tree = O.build_tree(nxyz)
    ref = O.object_feats(kept, nxyz, tree, full, vs, d)
    ref = np.stack([np.asarray(r, np.float32).reshape(-1) for r in ref])
    assert got.shape == ref.shape
    fin = np.isfinite(ref).all(1)
    # object points are averages of node centroids, so their nearest node is often an exact two-way tie that
    # cKDTree and the GPU may resolve differently (one row swapped for its neighbour): most objects agree to
    # float32 summation noise, none by more than one swapped row's weight
    err = np.abs(got[fin] - ref[fin]).max(1)
    assert np.median(err) < 2e-5 and err.max() < 3e-3, err
    assert np.all(got[-1] == 0)                         # the far object: no valid rows -> zeros


@pytest.mark.parametrize("seed,n_frames,n_obj,mpf", [(4, 6, 7, 4), (5, 5, 5, 5), (6, 1, 3, 4)])
def test_hierarchical_merge_vs_oracle(engine, seed, n_frames, n_obj, mpf):
    """pipeline.merge_type = "hierarchical" (graph.py:425-433 -> graph_utils.py:958-1012) on the same device kernels"""
    from holoagent_b200.memory.hmsg.graph.graph import hierarchical_merge
    frames, vs = _blob_scene(seed, n_frames, n_obj, mpf)
    ref = [o for o in O.hierarchical_merge(frames, 0.75, 0.025, vs, 0.05) if len(o[0]) >= 10]
    hierarchical_merge(engine, [_ragged(fm) for fm in frames], 0.75, 0.025, vs, 0.05)
    engine.objects_finish(10)
    off, xyz, col = engine.objects_read()
    assert len(off) - 1 == len(ref)
    for i, (p, c) in enumerate(ref):
        assert np.array_equal(xyz[off[i]:off[i + 1]], p), i
        assert np.array_equal(col[off[i]:off[i + 1]], c), i
