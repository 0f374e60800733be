"""tvc_retrieval_metrics / metrics.RetrievalEvaluator against the reference's own
RetrievalEvaluator.compute_retrieval_metrics (golden fixture) and the oracle restatement."""
from pathlib import Path

import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_reference_retrieval_evaluator_golden(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    z = np.load(GOLD / "retrieval_metrics.npz")
    for tag in ("small", "wide"):
        ks = [int(k) for k in z[f"{tag}_ks"]]
        m = tvc.RetrievalEvaluator.compute_retrieval_metrics(z[f"{tag}_sims"], z[f"{tag}_rel"], ks)
        for t, k in enumerate(ks):
            assert abs(m.recall_at_k[k] - z[f"{tag}_recall"][t]) < 1e-6, (tag, k)
            assert abs(m.precision_at_k[k] - z[f"{tag}_precision"][t]) < 1e-6, (tag, k)
            assert abs(m.ndcg_at_k[k] - z[f"{tag}_ndcg"][t]) < 1e-6, (tag, k)
        if tag == "small":            # 50 candidates: the list is the full ranking, so MRR / mAP are exact too
            assert abs(m.mrr - float(z["small_mrr"])) < 1e-6
            assert abs(m.map_score - float(z["small_map"])) < 1e-6
    # per query on the wide matrix: RR is exact when the first relevant item is inside the 64-long list
    sims, rel = z["wide_sims"], z["wide_rel"]
    order = np.argsort(-sims, axis=1, kind="stable")[:, :64]
    ptr = np.concatenate([[0], np.cumsum(rel.sum(1))]).astype(np.int64)
    idx = np.nonzero(rel)[1].astype(np.int64)
    per_q = tvc_ctx.retrieval_metrics(order, ptr, idx, [1, 10, 50])
    want = z["wide_per_query"]
    inside = want[:, 0] >= 1.0 / 64
    assert np.abs(per_q[inside, 0] - want[inside, 0]).max() < 1e-6
    assert (per_q[~inside, 0] == 0).all()


@pytest.mark.parametrize("q,n,k", [(1, 10, 10), (257, 5000, 56), (1000, 300, 20)])
def test_metrics_kernel_against_oracle(tvc_ctx, q, n, k):
    import torch
    rng = np.random.default_rng(q + k)
    topk = np.stack([rng.permutation(n)[:k] for _ in range(q)]).astype(np.int64)
    topk[rng.uniform(size=topk.shape) < 0.03] = -1                 # unused slots
    relevant = [list(rng.choice(n, size=rng.integers(0, 6), replace=False)) for _ in range(q)]
    for i in range(0, q, 3):                                        # make sure hits exist
        relevant[i] = list(set(relevant[i]) | {int(x) for x in topk[i, :3] if x >= 0})
    ks = sorted({1, min(5, k), k})
    ptr = np.concatenate([[0], np.cumsum([len(r) for r in relevant])]).astype(np.int64)
    idx = np.array([x for r in relevant for x in r], dtype=np.int64)
    want = O.retrieval_metrics_from_topk(topk, relevant, ks)
    got = tvc_ctx.retrieval_metrics(topk, ptr, idx, ks)
    assert np.abs(got - want).max() < 1e-6
    got_t = tvc_ctx.retrieval_metrics(torch.from_numpy(topk).cuda(), torch.from_numpy(ptr).cuda(),
                                      torch.from_numpy(idx).cuda(), ks)
    torch.cuda.synchronize()
    assert np.array_equal(got_t.cpu().numpy(), got)


def test_evaluate_searches_and_scores_without_the_matrix(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    g = O.synth_gallery(3000, 128, seed=2, clusters=50)
    rng = np.random.default_rng(0)
    pick = rng.integers(0, 3000, 200)
    q = O.l2_normalize(g[pick] + 0.02 * rng.standard_normal((200, 128)).astype(np.float32))
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    m = tvc.RetrievalEvaluator.evaluate(gal, q, [[int(p)] for p in pick], k_values=(1, 5, 10))
    _, idx = gal.search(q, 10)
    want = O.retrieval_metrics_from_topk(idx, [[int(p)] for p in pick], [1, 5, 10]).mean(0)
    assert abs(m.recall_at_k[10] - want[2 + 2]) < 1e-6 and abs(m.mrr - want[0]) < 1e-6
    assert m.recall_at_k[10] > 0.9
