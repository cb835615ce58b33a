"""Test double for HmsgEngine's retrieval calls (index_set / query_topk / query_object / query_scores), backed by
the numpy oracle.  TEST INFRASTRUCTURE ONLY: it lets the CPU suite exercise the host-side glue of the Graph mirror
(room selection, negative prompts, tie window, id mapping) without a GPU; the product never imports it."""
import numpy as np

from oracle import hmsg_oracle as O


class OracleRetrievalEngine:
    device = 0

    def __init__(self):
        self.E = None
        self.index_sets = 0

    def index_set(self, E, borrow=False):
        self.E = np.ascontiguousarray(np.asarray(E, dtype=np.float32))
        self._graph_index_tag = None
        self.index_N, self.index_d = self.E.shape
        self.index_sets += 1

    def query_topk(self, Q, k, row_mask=None, ids=None, scores=None):
        Q = np.asarray(Q, dtype=np.float32)
        out_i = np.full((len(Q), k), -1, np.int64); out_s = np.full((len(Q), k), -np.inf, np.float32)
        for r, q in enumerate(Q):
            i, s = O.query_topk(q, self.E, min(k, len(self.E)))
            out_i[r, :len(i)] = i; out_s[r, :len(i)] = s
        return out_i, out_s

    def query_object(self, Q, query_id, k, row_mask=None):
        Q = np.asarray(Q, dtype=np.float32)
        ids = np.full((len(Q), k), -1, np.int64); sc = np.zeros((len(Q), k), np.float32); nf = np.zeros(len(Q), np.int32)
        for r in range(len(Q)):
            sim = np.dot(Q[r], self.E.T)
            cls = np.argmax(sim, axis=0)
            n = min(k, int(np.sum(cls == query_id)))
            top, s = O.query_object_core(Q[r], self.E, query_id, k, True)
            ids[r, :n] = top[:n]; sc[r, :n] = s[:n]; nf[r] = n
        return ids, sc, nf

    def query_scores(self, Q):
        Q = np.asarray(Q, dtype=np.float32).reshape(-1, self.E.shape[1])
        return np.dot(Q, self.E.T).astype(np.float32)
