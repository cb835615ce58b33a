"""GPU: size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take minutes).

kNN 1 M x 512 (configs[2]): self-retrieval, sortedness, idempotence, linearity of the score, agreement of the
fused top-k with a top-k over the dense scores.  Ingest 640 x 480, M = 32 (configs[1] frame shape): batch-size
invariance, counter conservation (a frame adds exactly 1 to every node it touches - SURVEY H1), unit embeddings.
"""
import numpy as np
import pytest
import torch

from holoagent_b200 import ingest, synth

pytestmark = pytest.mark.gpu


def test_knn_1m_properties(engine):
    N, d, k = 1_000_000, 512, 5
    E, Q = synth.make_knn_tables(N, 64, d, device="cuda")
    torch.cuda.synchronize()
    engine.index_set(E, borrow=True)
    # (a) a stored row (scaled) retrieves itself with score 0.8 * |row|^2 = 0.8
    rows = torch.tensor([0, 1, 12345, 262143, 262144, 777777, N - 1], device="cuda")
    ids, sc = engine.query_topk((E[rows] * 0.8).contiguous(), k)
    assert torch.equal(ids[:, 0], rows)
    assert torch.allclose(sc[:, 0], torch.full((len(rows),), 0.8, device="cuda"), atol=1e-5)
    # (b) sorted, distinct, idempotent
    ids1, sc1 = engine.query_topk(Q, k)
    ids2, sc2 = engine.query_topk(Q, k)
    assert torch.equal(ids1, ids2) and torch.equal(sc1, sc2)
    assert bool((sc1[:, 1:] <= sc1[:, :-1]).all())
    assert all(len(set(r.tolist())) == k for r in ids1.cpu())
    # (c) batched passes (8 queries per pass) give what one query per call gives
    for i in (0, 7, 8, 63):
        a, b = engine.query_topk(Q[i:i + 1].contiguous(), k)
        assert torch.equal(a[0], ids1[i]) and torch.allclose(b[0], sc1[i], atol=1e-6)
    # (d) fused top-k == top-k of the dense scores (torch fp32 matmul as an independent checker, ties aside)
    dense = Q[:8] @ E.T
    tv, ti = dense.topk(k, dim=1)
    assert torch.allclose(sc1[:8], tv, atol=2e-5)
    assert (ids1[:8] == ti).float().mean() > 0.95
    # (e) linearity of the returned scores: s(q1 + q2) = s(q1) + s(q2) on the rows returned for q1 + q2
    q12 = (Q[0] + Q[1])[None].contiguous()
    i12, s12 = engine.query_topk(q12, k)
    assert torch.allclose(s12[0], dense[0][i12[0]] + dense[1][i12[0]], atol=2e-5)
    engine.index_set(E[:16].contiguous())     # drop the borrowed table before E is freed


def _ingest(engine, F, FB, boxes, d=512, M=32):
    job = ingest.IngestJob(engine, F, FB, M, d, boxes)
    job.step_device()
    engine.sync()
    s, c = engine.node_feats_raw()
    return job.full_feats.clone(), s, c


def test_ingest_640x480_properties(engine):
    F, H, W, M = 48, 480, 640, 32
    dpt, rgb, T, K = synth.make_frames(np.arange(F) * 3, H, W, device="cuda")
    engine.scene_begin(H, W, K, 1000.0, 0.05, F)
    engine.add_frames(dpt.view(torch.int16), rgb, torch.from_numpy(T.reshape(F, 16)).cuda())
    engine.sync()
    boxes = torch.from_numpy(np.stack([synth.make_mask_boxes(i, H, W, M) for i in range(F)])).cuda()
    engine.encoder_load(synth.make_vit_weights())
    full_a, sum_a, cnt_a = _ingest(engine, F, 16, boxes)
    n_nodes = engine.n_nodes
    # counter conservation: every frame adds exactly 1 to each node one of its valid pixels maps to
    expect = np.zeros(n_nodes, np.float32)
    valid = (dpt.cpu().numpy().reshape(F, -1) > 0)
    for b0 in range(0, F, 16):
        engine.masks_boxes(b0, boxes[b0:b0 + 16])
        for f in range(b0, b0 + 16):
            idx, _ = engine.pixel_to_node(f, want_dist=False)
            expect[np.unique(idx[valid[f]][idx[valid[f]] >= 0])] += 1
    assert np.array_equal(cnt_a, expect)
    assert cnt_a.max() <= F
    # batch-size invariance: the per-node sums walk the frames in the same order whatever the batching
    full_b, sum_b, cnt_b = _ingest(engine, F, 48, boxes)
    assert engine.n_nodes == n_nodes and np.array_equal(cnt_a, cnt_b)
    assert np.array_equal(sum_a, sum_b) and torch.equal(full_a, full_b)
    # finalize: sum / counter, untouched nodes stay exactly zero (counter 1e-5 in the reference)
    fa = full_a.cpu().numpy()
    assert np.isfinite(fa).all()
    assert np.all(fa[cnt_a == 0] == 0)
    t = cnt_a > 0
    assert np.allclose(fa[t], sum_a[t] / cnt_a[t][:, None], rtol=1e-6, atol=1e-7)
    # a node touched once carries one fp16-rounded unit feature (H7), or zero when its pixel lies in no mask
    once = cnt_a == 1
    if once.any():
        nrm = np.linalg.norm(fa[once], axis=1)
        assert np.all((nrm < 1e-6) | ((nrm > 0.995) & (nrm < 1.005)))
