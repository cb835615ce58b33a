"""GPU parity: A11 fused similarity + top-k vs the numpy oracle (graph.py:3126-3151)."""
import numpy as np
import pytest

from oracle import hmsg_oracle as O
from holoagent_b200 import synth

pytestmark = pytest.mark.gpu


def _check_topk(E, q, ids, scores, k):
    sim = E.astype(np.float64) @ q.astype(np.float64)
    oid, osc = O.query_topk(q, E, k)
    assert len(set(ids.tolist())) == k
    # scores are the true dot products (1e-3 relative contract; fp32 accumulation gives ~1e-6)
    assert np.allclose(scores, sim[ids], rtol=1e-3, atol=1e-6)
    assert np.all(np.diff(scores) <= 0)
    # the same rows as the oracle unless two scores are closer than fp32 summation noise
    if not np.array_equal(ids, oid):
        assert np.allclose(np.sort(sim[ids]), np.sort(sim[oid]), rtol=0, atol=2e-6)


@pytest.mark.parametrize("N,d", [(20000, 512), (777, 256), (5000, 1024)])
def test_query_topk(engine, N, d):
    E, Q = synth.make_knn_tables(N, 37, d)
    E, Q = E.numpy(), Q.numpy()
    engine.index_set(E)
    for nq, k in [(1, 5), (3, 1), (8, 24), (17, 10), (37, 32)]:
        ids, sc = engine.query_topk(Q[:nq], k)
        for i in range(nq):
            _check_topk(E, Q[i], ids[i], sc[i], k)


def test_row_mask_and_small_index(engine):
    E, Q = synth.make_knn_tables(300, 4, 512)
    E, Q = E.numpy(), Q.numpy()
    engine.index_set(E)
    mask = (np.arange(300) % 3 == 0).astype(np.uint8)
    ids, sc = engine.query_topk(Q, 5, row_mask=mask)
    sub = np.nonzero(mask)[0]
    for i in range(4):
        oid, osc = O.query_topk(Q[i], E[sub], 5)
        assert np.array_equal(ids[i], sub[oid])
    # k larger than the number of rows: padded with -1
    engine.index_set(E[:3])
    ids, sc = engine.query_topk(Q[:1], 5)
    assert set(ids[0][:3].tolist()) == {0, 1, 2} and np.all(ids[0][3:] == -1)


@pytest.mark.parametrize("Qp", [2, 5, 8, 22])
def test_query_object_negative_prompts(engine, Qp):
    N, d, k = 30000, 512, 5
    E, Qall = synth.make_knn_tables(N, 3 * Qp, d)
    E = E.numpy(); Qall = Qall.numpy().reshape(3, Qp, d)
    engine.index_set(E)
    for qid in (0, Qp - 1):
        ids, sc, nf = engine.query_object(Qall, qid, k)
        for r in range(3):
            top, osc = O.query_object_core(Qall[r], E, qid, k, True)
            sim = Qall[r].astype(np.float64) @ E.T.astype(np.float64)
            n = int(nf[r])
            assert n == min(k, int(np.sum(np.argmax(sim, axis=0) == qid)))
            if not np.array_equal(ids[r][:n], top[:n]):
                assert np.allclose(np.sort(sim[qid][ids[r][:n]]), np.sort(sim[qid][top[:n]]), atol=2e-6)
            assert np.allclose(sc[r][:n], sim[qid][ids[r][:n]], rtol=1e-3, atol=1e-6)


def test_large_k_ranked_path(engine):
    """k > 32 (callers rank a whole room: top_k = len(objects)) -> dense scores + key sort on the GPU."""
    N, d = 5000, 512
    E, Q = synth.make_knn_tables(N, 6, d)
    E, Q = E.numpy(), Q.numpy()
    engine.index_set(E)
    for k in (33, 100, N):
        ids, sc = engine.query_topk(Q[:2], k)
        for i in range(2):
            _check_topk(E, Q[i], ids[i], sc[i], k)
    mask = (np.arange(N) % 7 != 0).astype(np.uint8)
    ids, sc = engine.query_topk(Q[:1], 64, row_mask=mask)
    sub = np.nonzero(mask)[0]
    oid, _ = O.query_topk(Q[0], E[sub], 64)
    assert np.array_equal(ids[0], sub[oid])
    Qr = Q.reshape(2, 3, d)
    ids, sc, nf = engine.query_object(Qr, 1, 4000)
    for r in range(2):
        top, osc = O.query_object_core(Qr[r], E, 1, 4000, True)
        n = int(nf[r])
        assert n == len(top)
        assert np.array_equal(ids[r][:n], top) and np.all(ids[r][n:] == -1)
        assert np.allclose(sc[r][:n], osc, rtol=1e-3, atol=1e-6)
