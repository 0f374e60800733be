"""Retrieval evaluation on the device: mirrors `RetrievalMetrics` /
`RetrievalEvaluator` of src/utils/metrics.py:69-86,379-574.  The reference argsorts the full
[N_queries, N_candidates] similarity matrix and walks every query in Python; here the ranked lists
come from the exact top-k search (kernel a) and one kernel reduces them (tvc_retrieval_metrics).
Recall@K, Precision@K and NDCG@K are exact for every K up to the list length; MRR and mAP are those of
the full ranking whenever every relevant item lies inside the list (always true for matrices with at
most 64 candidates), otherwise they are the list-truncated values."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _native as N

MAX_LIST = 56          # TVC_MAX_K: longest ranked list the search kernel returns


@dataclass
class RetrievalMetrics:
    """src/utils/metrics.py:69-86."""
    recall_at_k: Dict[int, float]
    precision_at_k: Dict[int, float]
    map_score: float
    ndcg_at_k: Dict[int, float]
    mrr: float

    def to_dict(self) -> Dict[str, Any]:
        return {"recall_at_k": self.recall_at_k, "precision_at_k": self.precision_at_k, "map_score": self.map_score,
                "ndcg_at_k": self.ndcg_at_k, "mrr": self.mrr}


def _csr(relevant) -> (np.ndarray, np.ndarray):
    """list of per-query index lists, or a dense 0/1 matrix -> (rel_ptr [q+1], rel_idx)."""
    if isinstance(relevant, np.ndarray) and relevant.ndim == 2:
        rows, cols = np.nonzero(relevant)
        ptr = np.zeros(relevant.shape[0] + 1, np.int64)
        np.add.at(ptr, rows + 1, 1)
        return np.cumsum(ptr), cols.astype(np.int64)
    ptr = np.zeros(len(relevant) + 1, np.int64)
    ptr[1:] = np.cumsum([len(r) for r in relevant])
    idx = np.concatenate([np.asarray(r, np.int64).ravel() for r in relevant]) if len(relevant) else np.zeros(0, np.int64)
    return ptr, idx


class RetrievalEvaluator:
    """src/utils/metrics.py:379-574."""

    @staticmethod
    def from_topk(topk_idx, relevant, k_values: Sequence[int] = (1, 5, 10, 20, 50),
                  ctx: Optional[N.Context] = None) -> RetrievalMetrics:
        """Ranked lists [q, k] (numpy or torch cuda, the search output) + relevance (dense 0/1 matrix or one
        index list per query) -> RetrievalMetrics.  K values beyond the list length are clipped to it."""
        ctx = ctx or N.Context.get()
        k = int(topk_idx.shape[1])
        ks = [min(int(x), k) for x in k_values]
        ptr, idx = _csr(relevant)
        if N._is_torch(topk_idx):
            import torch
            ptr_t, idx_t = torch.from_numpy(ptr).to(topk_idx.device), torch.from_numpy(idx).to(topk_idx.device)
            per_q = ctx.retrieval_metrics(topk_idx, ptr_t, idx_t, ks).double().mean(0).cpu().numpy()
        else:
            per_q = ctx.retrieval_metrics(np.asarray(topk_idx), ptr, idx, ks).astype(np.float64).mean(0)
        nk = len(ks)
        return RetrievalMetrics(
            recall_at_k={int(kv): float(per_q[2 + t]) for t, kv in enumerate(k_values)},
            precision_at_k={int(kv): float(per_q[2 + nk + t]) for t, kv in enumerate(k_values)},
            map_score=float(per_q[1]),
            ndcg_at_k={int(kv): float(per_q[2 + 2 * nk + t]) for t, kv in enumerate(k_values)},
            mrr=float(per_q[0]))

    @staticmethod
    def compute_retrieval_metrics(similarities: np.ndarray, relevance: np.ndarray,
                                  k_values: List[int] = [1, 5, 10, 20, 50]) -> RetrievalMetrics:  # noqa: B006
        """The reference's signature (src/utils/metrics.py:386-459): a precomputed similarity matrix.  The
        ranked lists are a stable descending sort on the device (ties to the lower index; plumbing, not
        a kernel of ours - at gallery scale use evaluate(), which never builds the matrix); the reduction
        is tvc_retrieval_metrics."""
        import torch
        ctx = N.Context.get()
        sims = torch.as_tensor(np.ascontiguousarray(similarities, dtype=np.float32)).to(f"cuda:{ctx.device}")
        k = min(int(sims.shape[1]), 64)
        order = torch.sort(-sims, dim=1, stable=True).indices[:, :k].contiguous()
        return RetrievalEvaluator.from_topk(order, np.asarray(relevance), k_values, ctx=ctx)

    @staticmethod
    def evaluate(gallery: N.Gallery, query_features, relevant, k_values: Sequence[int] = (1, 5, 10, 20, 50),
                 normalize_queries: bool = False) -> RetrievalMetrics:
        """Search + metrics without the similarity matrix ever existing (what the evaluation loops of
        experiments/run_experiments.py need at gallery scale)."""
        k = min(max(int(x) for x in k_values), MAX_LIST, len(gallery))
        _, idx = gallery.search(query_features, k, normalize_queries=normalize_queries)
        return RetrievalEvaluator.from_topk(idx, relevant, k_values, ctx=gallery.ctx)


@dataclass
class SimilarityMetrics:
    """src/utils/metrics.py:88-105."""
    cosine_similarity: float = 0.0
    euclidean_distance: float = 0.0
    manhattan_distance: float = 0.0
    pearson_correlation: float = 0.0
    spearman_correlation: float = 0.0

    def to_dict(self) -> Dict[str, float]:
        return {"cosine_similarity": self.cosine_similarity, "euclidean_distance": self.euclidean_distance,
                "manhattan_distance": self.manhattan_distance, "pearson_correlation": self.pearson_correlation,
                "spearman_correlation": self.spearman_correlation}


def _host(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


class SimilarityCalculator:
    """src/utils/metrics.py:107-164.  `cosine_similarity` of two vectors is a few hundred flops and stays on
    the host (same zero-vector guard, same fp64 arithmetic as scipy's `1 - cosine`); `batch_cosine_similarity`
    is the [N, D] x [M, D] cosine matrix and runs as kernel (a)'s GEMM with the plain-store epilogue
    (tvc_similarity_matrix: bf16 operands, fp32 accumulate - within 2e-3 of the reference's fp32 `mm`)."""

    @staticmethod
    def cosine_similarity(x, y) -> float:
        a, b = _host(x).astype(np.float64).ravel(), _host(y).astype(np.float64).ravel()
        na, nb = np.linalg.norm(a), np.linalg.norm(b)
        if na == 0 or nb == 0:
            return 0.0
        return float(np.dot(a, b) / (na * nb))

    # the remaining scalar measures of the reference class (:166-268): two vectors in, one float out - host
    @staticmethod
    def euclidean_distance(x, y) -> float:
        return float(np.linalg.norm(_host(x).ravel() - _host(y).ravel()))

    @staticmethod
    def manhattan_distance(x, y) -> float:
        return float(np.sum(np.abs(_host(x).ravel() - _host(y).ravel())))

    @staticmethod
    def pearson_correlation(x, y) -> float:
        try:
            from scipy.stats import pearsonr
            corr = pearsonr(_host(x).ravel(), _host(y).ravel())[0]
            return float(corr) if not np.isnan(corr) else 0.0
        except Exception:  # noqa: BLE001
            return 0.0

    @staticmethod
    def spearman_correlation(x, y) -> float:
        try:
            from scipy.stats import spearmanr
            corr = spearmanr(_host(x).ravel(), _host(y).ravel())[0]
            return float(corr) if not np.isnan(corr) else 0.0
        except Exception:  # noqa: BLE001
            return 0.0

    @classmethod
    def compute_all_similarities(cls, x, y) -> "SimilarityMetrics":
        return SimilarityMetrics(cosine_similarity=cls.cosine_similarity(x, y),
                                 euclidean_distance=cls.euclidean_distance(x, y),
                                 manhattan_distance=cls.manhattan_distance(x, y),
                                 pearson_correlation=cls.pearson_correlation(x, y),
                                 spearman_correlation=cls.spearman_correlation(x, y))

    @staticmethod
    def batch_cosine_similarity(x, y) -> np.ndarray:
        rows = np.ascontiguousarray(_host(y), dtype=np.float32)
        queries = np.ascontiguousarray(_host(x), dtype=np.float32)
        gal = N.Gallery(rows, normalize=True)
        try:
            return np.asarray(gal.similarity_matrix(queries, normalize_queries=True))
        finally:
            gal.close()
