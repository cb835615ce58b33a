"""GPU parity: batched A7 (create_3d_masks for a whole frame batch, generic.py:140-190; graph.py:391-402),
ragged per-frame mask counts (extractor.py:168-172 softmax over the frame's own masks) and the stored-frame N1 merge."""
import numpy as np
import pytest

from oracle import hmsg_oracle as O
from holoagent_b200 import synth
from tests.scenes import scene, load_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nodes(engine):
    sc = scene()
    load_scene(engine, sc)
    engine.voxel_build()
    engine.radius_filter(150, 0.4)
    nxyz, nrgb, _, _ = engine.nodes_read()
    return sc, nxyz, nrgb, O.build_tree(nxyz)


def _oracle_masks(sc, f, segs, nxyz, nrgb, tree, engine):
    """oracle create_3d_masks fed the GPU's own pixel->node map (exact NN ties are implementation-defined, see test_gpu_geometry)"""
    gidx, _ = engine.pixel_to_node(f, want_dist=False)
    out = []
    for seg in segs:
        sel = (seg & (sc["depth"][f] > 0)).reshape(-1)
        if not sel.any():
            out.append((np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3), np.int32)))
            continue
        idx = gidx[sel]
        p, c, k, _ = O.voxel_down_sample(nxyz[idx], nrgb[idx], sc["vs"])
        out.append((p, c, k))
    return out


@pytest.mark.parametrize("dense,M", [(True, 12), (False, 40)])
def test_mask_nodes_batch_vs_oracle(engine, nodes, dense, M):
    sc, nxyz, nrgb, tree = nodes
    b0, nb = 2, 5
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][b0 + i]), sc["H"], sc["W"], M) for i in range(nb)])
    segs = []
    for i in range(nb):
        valid = sc["depth"][b0 + i] > 0
        s = np.zeros((M, sc["H"], sc["W"]), bool)
        for m, (x, y, w, h) in enumerate(boxes[i]):
            s[m, y:y + h, x:x + w] = valid[y:y + h, x:x + w] if not dense else True     # dense masks may cover depth == 0 pixels
        segs.append(s)
    if dense:
        engine.masks_dense(b0, np.stack(segs).astype(np.uint8))
    else:
        engine.masks_boxes(b0, boxes)
    engine.mask_store_reset()
    engine.mask_nodes_batch(b0, nb, sc["vs"], keep=True)
    nf, nm, npts = engine.mask_store_count()
    assert (nf, nm) == (nb, nb * M)
    tot = 0
    for i in range(nb):
        off, xyz, rgb, ijk = engine.mask_store_read(b0 + i)
        ref = _oracle_masks(sc, b0 + i, segs[i], nxyz, nrgb, tree, engine)
        assert len(off) == M + 1
        for m in range(M):
            p, c, k = ref[m]
            a, b = off[m], off[m + 1]
            assert b - a == len(p), (i, m)
            assert np.array_equal(ijk[a:b], k)                       # voxel keys in canonical order: exact
            assert np.allclose(xyz[a:b], p, rtol=1e-12, atol=1e-12)   # k_n * c_n by node vs Open3D's per-pixel sequential sum
            assert np.allclose(rgb[a:b], c, rtol=1e-12, atol=1e-12)
        tot += off[-1]
    assert tot == npts
    # the one-frame host API (hmsg_mask_nodes) reads the same batch result
    off1, xyz1, rgb1, ijk1 = engine.mask_nodes(b0 + 1, sc["vs"], M)
    off2, xyz2, rgb2, ijk2 = engine.mask_store_read(b0 + 1)
    assert np.array_equal(off1, off2) and np.array_equal(xyz1, xyz2) and np.array_equal(ijk1, ijk2)
    # determinism: a second pass gives the same bits
    engine.mask_store_reset()
    engine.mask_nodes_batch(b0, nb, sc["vs"], keep=True)
    off3, xyz3, _, _ = engine.mask_store_read(b0 + 1)
    assert np.array_equal(xyz3, xyz2)


def test_filter_distance_and_ragged_counts(engine, nodes):
    sc, nxyz, nrgb, tree = nodes
    b0, nb, M = 0, 4, 8
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][b0 + i]), sc["H"], sc["W"], M) for i in range(nb)])
    counts = np.array([8, 3, 0, 5], np.int32)
    engine.masks_boxes(b0, boxes)
    engine.masks_counts(b0, counts)
    # mean depth per mask as the reference computes it (float32 mean of depth / scale)
    engine.mask_store_reset()
    thr = 2.5
    engine.mask_nodes_batch(b0, nb, sc["vs"], filter_distance=thr, keep=True)
    assert engine.mask_store_count()[:2] == (nb, int(counts.sum()))
    for i in range(nb):
        off, xyz, _, _ = engine.mask_store_read(b0 + i)
        assert len(off) == counts[i] + 1                              # padded slots are not list entries
        dep = sc["depth"][b0 + i].astype(np.float32) / np.float32(sc["scale"])
        for m in range(counts[i]):
            x, y, w, h = boxes[i, m]
            z = dep[y:y + h, x:x + w]
            z = z[z > 0]
            if len(z) and abs(float(z.mean()) - thr) < 1e-4:
                continue                                               # float32 pairwise mean vs integer-sum mean: 1e-7 apart
            empty = len(z) == 0 or z.mean() > thr
            assert (off[m + 1] - off[m] == 0) == empty, (i, m)


def test_fuse_with_ragged_counts(engine, nodes):
    """padded mask slots take no part in the softmax (real SAM output is ragged)"""
    sc, nxyz, _, _ = nodes
    d, M, nb, b0 = 256, 6, 3, 1
    rs = np.random.RandomState(3)
    feats = rs.randn(nb, 2 * M + 1, d).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=-1, keepdims=True)
    counts = np.array([6, 2, 4], np.int32)
    boxes = np.stack([synth.make_mask_boxes(int(sc["ids"][b0 + i]), sc["H"], sc["W"], M) for i in range(nb)])
    for i in range(nb):
        boxes[i, counts[i]:] = (0, 0, 1, 1)
    engine.features_begin(d)
    engine.masks_boxes(b0, boxes)
    engine.masks_counts(b0, counts)
    Fp = engine.fuse_scatter(b0, nb, M, feats, 0.4418)
    for i in range(nb):
        k = counts[i]
        ref = O.fuse_mask_feats(feats[i, :k], feats[i, M:M + k], feats[i, 2 * M:2 * M + 1], 0.4418)
        assert np.allclose(Fp[i, :k], ref, rtol=0, atol=2e-6), i
    # and differs from the padded softmax for the short frames (the bug this guards against)
    engine.features_begin(d)
    engine.masks_boxes(b0, boxes)
    Fp_pad = engine.fuse_scatter(b0, nb, M, feats, 0.4418)
    assert np.abs(Fp_pad[1, :2] - Fp[1, :2]).max() > 1e-4


def test_merge_stored_equals_per_frame_chain(engine):
    """N1 fed from the mask store (batched A7) == N1 fed frame by frame (ordered sums): same objects, points within 1e-9"""
    from tests.test_gpu_objects import _fixture_scene
    z, depth, rgb, T, K, segs, vs, nxyz, nrgb = _fixture_scene(engine)
    F = len(depth)
    engine.objects_begin(0.75, vs, 0.05)
    for f in range(F):
        engine.masks_dense(f, segs[f][None].astype(np.uint8))
        engine.objects_add_frame(f, vs, 6.0)
    engine.objects_finish(10)
    off_a, xyz_a, col_a = engine.objects_read()
    engine.mask_store_reset()
    M = max(len(s) for s in segs)
    for b0 in range(0, F, 4):
        n = min(4, F - b0)
        seg = np.zeros((n, M) + depth[0].shape, np.uint8)
        cnt = np.zeros(n, np.int32)
        for i in range(n):
            seg[i, :len(segs[b0 + i])] = segs[b0 + i]
            cnt[i] = len(segs[b0 + i])
        engine.masks_dense(b0, seg)
        engine.masks_counts(b0, cnt)
        engine.mask_nodes_batch(b0, n, vs, filter_distance=6.0, keep=True)
    engine.objects_begin(0.75, vs, 0.05)
    engine.objects_merge_stored(0, F)
    engine.objects_finish(10)
    off_b, xyz_b, col_b = engine.objects_read()
    assert np.array_equal(off_a, off_b)
    assert np.allclose(xyz_a, xyz_b, rtol=1e-9, atol=1e-9) and np.allclose(col_a, col_b, rtol=1e-9, atol=1e-9)
