"""GPU parity at BASELINE.json's frame shapes (geometry / NN / scatter parity had only run at 320x240, the
d = 768 template instances of k_fuse / k_scatter_batch / k_sim_topk had no GPU test): A1-A6 vs the oracle on a few
640x480 (configs[1]) and 1280x720 (configs[3]) frames, and the d = 768 (ViT-L/14, the reference's default tower) widths."""
import numpy as np
import pytest
import torch

from oracle import hmsg_oracle as O
from holoagent_b200 import synth

pytestmark = pytest.mark.gpu


def _sq(a, b):
    d = a - b
    return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]


@pytest.mark.parametrize("H,W,F,M,d", [(480, 640, 3, 32, 512), (720, 1280, 2, 16, 768)])
def test_geometry_nn_scatter_vs_oracle_at_config_shapes(engine, H, W, F, M, d):
    ids = np.arange(F) * 7
    depth, rgb, T, K = synth.make_frames_np(ids, H, W)
    engine.scene_begin(H, W, K, 1000.0, 0.05, F)
    engine.add_frames(depth, rgb, T.reshape(F, 16))
    # ---- A1/A2: voxel table (keys, counts exact; centroids to the last ulps)
    nv, mb = engine.voxel_build()
    P, C = [], []
    for f in range(F):
        p, c, _ = O.create_pcd(rgb[f], depth[f], K, 1000.0, T[f])
        P.append(p); C.append(c)
    P = np.concatenate(P); C = np.concatenate(C)
    ovx, ovc, oijk, inv = O.voxel_down_sample(P, C, 0.05)
    vx, vc, vijk, cnt = engine.voxels_read()
    assert nv == len(ovx) and np.array_equal(mb, P.min(axis=0))
    assert np.array_equal(vijk, oijk) and np.array_equal(cnt, np.bincount(inv, minlength=nv))
    assert np.allclose(vx, ovx, rtol=1e-12, atol=1e-12) and np.allclose(vc, ovc, rtol=1e-12, atol=1e-12)
    # ---- A3: radius filter (a sample of the counts against cKDTree, the keep set against the counts)
    nb, radius = 120, 0.35
    n_nodes = engine.radius_filter(nb, radius)
    counts = engine.radius_counts()
    from scipy.spatial import cKDTree
    tree = cKDTree(vx)
    for i in np.random.RandomState(0).choice(nv, size=600, replace=False):
        js = np.asarray(tree.query_ball_point(vx[i], radius * (1 + 1e-9) + 1e-12))
        dl = vx[js] - vx[i]
        assert counts[i] == np.count_nonzero((dl[:, 0] * dl[:, 0] + dl[:, 1] * dl[:, 1]) + dl[:, 2] * dl[:, 2] < radius * radius)
    keep = np.nonzero(counts > nb)[0]
    nxyz, _, nijk, nvox = engine.nodes_read()
    assert n_nodes == len(keep) and np.array_equal(nvox, keep) and np.array_equal(nijk, vijk[keep])
    # ---- A4: pixel -> node == cKDTree (exact distance ties aside)
    nt = cKDTree(nxyz)
    gids = []
    for f in range(F):
        idx, dist = engine.pixel_to_node(f)
        p, _, m = O.create_pcd(rgb[f], depth[f], K, 1000.0, T[f])
        od, oi = nt.query(p, k=1)
        valid = m.reshape(-1)
        gi = idx[valid]
        bad = np.nonzero(gi != oi)[0]
        for b in bad:
            assert _sq(nxyz[gi[b]], p[b]) == _sq(nxyz[oi[b]], p[b]) and gi[b] < oi[b]
        assert len(bad) <= 1e-3 * len(gi)
        assert np.allclose(dist[valid], od, rtol=1e-12, atol=1e-14)
        gids.append(gi)
    # ---- A5/A6: fusion + last-writer-wins scatter (d = 512 and the d = 768 template instances)
    engine.features_begin(d)
    rs = np.random.RandomState(5)
    feats = rs.randn(F, 2 * M + 1, d).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=-1, keepdims=True)
    boxes = np.stack([synth.make_mask_boxes(int(ids[i]), H, W, M) for i in range(F)])
    engine.masks_boxes(0, boxes)
    Fp = engine.fuse_scatter(0, F, M, feats, 0.4418)
    sum_f = torch.zeros(n_nodes, d); cn = torch.zeros(n_nodes, 1)
    for f in range(F):
        oFp = O.fuse_mask_feats(feats[f, :M], feats[f, M:2 * M], feats[f, 2 * M:2 * M + 1], 0.4418)
        assert np.allclose(Fp[f], oFp, rtol=0, atol=2e-6)
        valid = depth[f] > 0
        segs = np.zeros((M, H, W), bool)
        for m, (x, y, w, h) in enumerate(boxes[f]):
            segs[m, y:y + h, x:x + w] = valid[y:y + h, x:x + w]
        O.ingest_frame(sum_f, cn, nt, n_nodes, depth[f], rgb[f], T[f], K, 1000.0, oFp, segs, idx=gids[f])
    gs, gc = engine.node_feats_raw()
    assert np.array_equal(gc, cn.numpy().reshape(-1))
    assert np.allclose(gs, sum_f.numpy(), rtol=0, atol=1e-3)
    assert np.mean(np.abs(gs - sum_f.numpy()) > 1e-6) < 1e-3


def test_knn_d768(engine):
    """k_sim_topk<d=768>: ids identical, scores ~1e-6 (graph.py:3127-3133 semantics on ViT-L/14 embeddings)"""
    N, d, k = 30000, 768, 5
    E, Q = synth.make_knn_tables(N, 24, d)
    E = E.numpy(); Q = Q.numpy()
    engine.index_set(E)
    ids, sc = engine.query_topk(Q, k)
    for r in range(len(Q)):
        oi, osc = O.query_topk(Q[r], E, k)
        assert np.array_equal(ids[r], oi), r
        assert np.allclose(sc[r], osc, atol=2e-6)
    # negative prompts (query_hmsg_object core) at the same width
    Qp = np.stack([Q[:3], Q[3:6]])
    gi, gs, nf = engine.query_object(Qp, 0, 4)
    for r in range(2):
        top, s = O.query_object_core(Qp[r], E, 0, 4, True)
        n = int(nf[r])
        assert n == min(4, len(top)) and np.array_equal(gi[r, :n], top[:n]) and np.allclose(gs[r, :n], s[:n], atol=2e-6)
