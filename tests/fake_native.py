"""Oracle-backed test double of the native layer (`_native.Gallery`, `_native.Context`) so that the HOST
logic of the Python mirrors (retrieval.py, detector.py, ref_bank.py, defenses.py) can run in a container
without a GPU - in particular under the reference's own orchestrator (tests/golden/pipeline_dropin.py).
TEST INFRASTRUCTURE ONLY: the product has no CPU path (tvc_ctx_create -> TVC_ERR_NO_DEVICE) and nothing
under multimodal_detection_consistency_b200/ imports this file or the oracle."""
from __future__ import annotations

import contextlib
import math

import numpy as np

from oracle import tvc_oracle as O


def _np(x, dtype=np.float32):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


class FakeGallery:
    def __init__(self, rows=None, dim=None, *, normalize=False, keep_master=True, global_row_offset=0, capacity=0,
                 ctx=None, device=None):
        self.normalize = bool(normalize)
        self.global_row_offset = int(global_row_offset)
        self.dim = int(dim if rows is None else np.asarray(rows).shape[1])
        self.rows = np.zeros((0, self.dim), np.float32)
        self.ctx = ctx if ctx is not None else FakeContext()
        if rows is not None:
            self.append(rows)

    def __len__(self):
        return int(self.rows.shape[0])

    ntotal = property(__len__)

    def append(self, rows):
        rows = _np(rows).reshape(-1, self.dim)
        self.rows = np.concatenate([self.rows, O.l2_normalize(rows) if self.normalize else rows])

    def truncate(self, n):
        self.rows = self.rows[:n]

    def move_row(self, src, dst):
        self.rows[dst] = self.rows[src]

    def get_rows(self, idx):
        return self.rows[_np(idx, np.int64)].copy()

    def search(self, queries, k, threshold=-math.inf, *, normalize_queries=False, skip_self=False):
        q = _np(queries)
        lead = q.shape[:-1] if q.ndim != 2 else None
        q = q.reshape(-1, self.dim)
        if normalize_queries:
            q = O.l2_normalize(q)
        s, i = O.search(q, self.rows, int(k), threshold=float(threshold), index_offset=self.global_row_offset,
                        skip_self=skip_self)
        if lead is not None and len(lead) != 1:
            s, i = s.reshape(*lead, k), i.reshape(*lead, k)
        return s, i

    def similarity_matrix(self, queries, *, normalize_queries=False):
        q = _np(queries).reshape(-1, self.dim)
        return O.similarity_matrix(O.l2_normalize(q) if normalize_queries else q, self.rows)

    def close(self):
        self.rows = self.rows[:0]


class FakeContext:
    def __init__(self):
        self.launches = 0

    def consistency_emb(self, params, img, txt, var=None, ret_gallery=None, ret_idx=None, gen=None, g_cnt=None,
                        gen_gallery=None, gen_idx=None, return_sims=False):
        p = params.as_dict()
        scores, flags, lists = O.consistency_emb(
            _np(img), _np(txt), _np(var), ret_rows=ret_gallery.rows if ret_gallery is not None else None,
            ret_idx=_np(ret_idx, np.int64).reshape(len(img), -1) if ret_idx is not None else None, gen=_np(gen),
            g_cnt=_np(g_cnt, np.int32), gen_rows=gen_gallery.rows if gen_gallery is not None else None,
            gen_idx=_np(gen_idx, np.int64).reshape(len(img), -1) if gen_idx is not None else None, params=p,
            ret_offset=ret_gallery.global_row_offset if ret_gallery is not None else 0,
            gen_offset=gen_gallery.global_row_offset if gen_gallery is not None else 0)
        self.launches += 1
        scores = scores.astype(np.float32)
        if not return_sims:
            return scores, flags

        def pad(rows, width):
            out = np.zeros((len(rows), width), np.float32)
            for r, vals in enumerate(rows):
                out[r, :len(vals)] = vals[:width]
            return out

        return scores, flags, (pad(lists[0], p["n_variants"]), pad(lists[1], p["n_retrieval"]),
                               pad(lists[2], p["n_generative"]))

    def consistency_sims(self, params, s0, sv=None, sr=None, r_cnt=None, sg=None, g_cnt=None, sxv=None):
        scores, flags = O.consistency_sims(_np(s0), _np(sv), _np(sr), _np(r_cnt, np.int32), _np(sg),
                                           _np(g_cnt, np.int32), _np(sxv), params=params.as_dict())
        return scores.astype(np.float32), flags

    def k_occurrence(self, idx, n_bins, idx_base=0, counts=None):
        c = O.k_occurrence(_np(idx, np.int64), int(n_bins), int(idx_base)).astype(np.int32)
        if counts is not None:
            counts += c
            return counts
        return c

    def retrieval_metrics(self, topk_idx, rel_ptr, rel_idx, k_values):
        ptr, idx = _np(rel_ptr, np.int64), _np(rel_idx, np.int64)
        relevant = [idx[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]
        return O.retrieval_metrics_from_topk(_np(topk_idx, np.int64), relevant, [int(k) for k in k_values]).astype(np.float32)

    def release_workspace(self):
        return 0


@contextlib.contextmanager
def installed():
    """Swap the test double in for the duration of the block (and restore the real bindings after)."""
    import multimodal_detection_consistency_b200._native as N
    from multimodal_detection_consistency_b200 import (defenses, detector, faiss_compat, hubness, metrics, ref_bank,
                                                       retrieval)
    ctx = FakeContext()
    saved = [(N, "Gallery", N.Gallery), (N.Context, "get", N.Context.__dict__["get"])]
    for mod in (retrieval, ref_bank, defenses, detector, hubness, metrics, faiss_compat):
        if "Gallery" in vars(mod):
            saved.append((mod, "Gallery", mod.Gallery))
    try:
        N.Gallery = FakeGallery
        N.Context.get = classmethod(lambda cls, device=None: ctx)
        for obj, name, _ in saved[2:]:
            setattr(obj, name, FakeGallery)
        yield ctx
    finally:
        for obj, name, val in saved:
            setattr(obj, name, val)
