"""Pins the NumPy oracle against outputs of the reference's own classes (tests/golden/*.npz, made by
tests/golden/make_golden.py from /root/reference).  CPU only."""
from pathlib import Path

import numpy as np
import pytest

from oracle import tvc_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return np.load(GOLD / name, allow_pickle=False)


# ------------------------------------------------------------------------------ ReferenceBank
def test_ref_bank_snapshot_queries():
    """The one fixture the reference ships (cache/ref_bank/references.json) through the reference's
    ReferenceBank.query_similar."""
    z = load("ref_bank.npz")
    vecs, qs = z["snap_vectors"], z["snap_queries"]
    hits = 0
    for i, q in enumerate(qs):
        idx, sim = O.ref_bank_query(vecs, q, top_k=10, similarity_threshold=None,
                                    config_threshold=float(z["snap_cfg_threshold"]))
        n = int((z["snap_idx"][i] >= 0).sum())
        assert len(idx) == n
        assert np.array_equal(idx, z["snap_idx"][i, :n])
        assert np.allclose(sim, z["snap_sim"][i, :n], rtol=0, atol=1e-12)
        hits += n
    assert hits > 0


def test_ref_bank_thresholds_and_falsy_zero():
    z = load("ref_bank.npz")
    bank, qs = z["bank"], z["queries"]
    cfg = float(z["cfg_threshold"])
    for ti, ta in enumerate(z["thr_args"]):
        thr = None if np.isnan(ta) else float(ta)
        for qi, q in enumerate(qs):
            idx, sim = O.ref_bank_query(bank, q, top_k=7, similarity_threshold=thr, config_threshold=cfg)
            n = int((z["idx"][ti, qi] >= 0).sum())
            assert np.array_equal(idx, z["idx"][ti, qi, :n])
            assert np.allclose(sim, z["sim"][ti, qi, :n], atol=1e-6)
    # explicit 0.0 behaves like None (src/ref_bank.py:191)
    assert np.array_equal(z["idx"][0], z["idx"][1])


# ------------------------------------------------------------------------------ ConsistencyChecker
@pytest.mark.parametrize("voting", ["simple", "weighted", "adaptive"])
@pytest.mark.parametrize("adaptive", [0, 1])
def test_consistency_checker(voting, adaptive):
    z = load("consistency_checker.npz")
    keys = [str(k) for k in z["keys"]]
    S = z["scores"]
    want = z[f"{voting}_{adaptive}"]
    col = {k: i for i, k in enumerate(keys)}
    params = dict(voting={"simple": 0, "weighted": 1, "adaptive": 2}[voting], cc_adaptive=adaptive)
    for i in range(S.shape[0]):
        got = cc_from_scores(S[i], col, params)
        assert np.allclose(got[:3], want[i, :3], atol=1e-12), (i, got, want[i])
        assert bool(got[3]) == bool(want[i, 3])


def cc_from_scores(row, col, params):
    """The oracle's ConsistencyChecker restatement on a reference-style score dict."""
    return O.consistency_from_scores(
        row[col["original_similarity"]], row[col["text_variant_consistency"]], row[col["text_variant_std"]],
        row[col["retrieval_consistency"]], row[col["retrieval_std"]], row[col["generative_consistency"]],
        row[col["generative_std"]], row[col["cross_modal_variance"]], params)


def test_consistency_checker_history():
    """Stateful threshold smoothing (consistency_checker.py:234-239) is host-side; the oracle's
    helper reproduces the reference's 40-decision run."""
    z = load("consistency_checker.npz")
    keys = [str(k) for k in z["keys"]]
    col = {k: i for i, k in enumerate(keys)}
    S, want = z["scores"], z["history"]
    hist = []
    for i in range(want.shape[0]):
        r = S[i]
        got = O.consistency_from_scores(
            r[col["original_similarity"]], r[col["text_variant_consistency"]], r[col["text_variant_std"]],
            r[col["retrieval_consistency"]], r[col["retrieval_std"]], r[col["generative_consistency"]],
            r[col["generative_std"]], r[col["cross_modal_variance"]], dict(voting=1, cc_adaptive=1),
            threshold_history=hist)
        hist.append(got[1])
        assert np.allclose(got[:3], want[i, :3], atol=1e-12), i
        assert bool(got[3]) == bool(want[i, 3])


# ------------------------------------------------------------------------------ similarity / top-k
def test_scalar_and_batch_cosine():
    z = load("similarity.npz")
    x, y = z["x"], z["y"]
    got = np.array([O.scalar_cosine(x[i], y[i]) for i in range(len(x))])
    assert np.allclose(got, z["pair"], atol=5e-7)  # scipy evaluates fp32 inputs in fp32
    assert got[5] == 0.0  # zero-vector guard, src/utils/metrics.py:137-139
    assert np.allclose(O.similarity_matrix(x[6:], y[6:], "cosine"), z["batch_np"], atol=2e-6)
    assert np.allclose(O.similarity_matrix(x[6:], y[6:], "cosine"), z["batch_t"], atol=2e-6)


def test_search_matches_reference_fallback():
    """MultiModalRetriever._search_index, sklearn branch (src/retrieval.py:669-671)."""
    z = load("similarity.npz")
    s, i = O.search(z["queries"], z["gallery"], 10, metric="cosine")
    assert np.array_equal(i, z["idx"])
    assert np.allclose(s, z["scores"], atol=2e-6)


def test_similarity_matrix_metrics():
    z = load("similarity.npz")
    q, g = z["queries"][:8], z["gallery"]
    assert np.allclose(O.similarity_matrix(q, g, "cosine"), z["mat_cos"], atol=2e-6)
    assert np.allclose(O.similarity_matrix(q, g, "dot_product"), z["mat_dot"], atol=2e-6)
    assert np.allclose(O.similarity_matrix(q, g, "euclidean"), z["mat_euc"], atol=2e-5)


def test_consistency_calculator_helpers():
    z = load("similarity.npz")
    d = O.similarity_distribution(z["scores"][0])
    assert np.allclose([d[k] for k in ["mean", "std", "min", "max", "median"]], z["dist"], atol=1e-7)
    got = [O.topk_overlap(z["idx"][i], z["idx"][i + 1], 10) for i in range(63)]
    assert np.array_equal(np.array(got), z["overlap"])


# ------------------------------------------------------------------------------ detectors
@pytest.mark.parametrize("mode_i", [0, 1, 2, 3])
def test_adversarial_detector_scores(mode_i):
    """src/detector.py detect_adversarial driven with table encoders."""
    z = load("detectors.npz")
    img, txt, var, gen, g_cnt = z["img"], z["txt"], z["var"], z["gen"], z["g_cnt"]
    want = z["det_scores"][mode_i]
    scores, flags, _ = O.consistency_emb(img, txt, var, gen=gen, g_cnt=g_cnt, params=dict(aggregation=mode_i))
    assert np.allclose(scores[:, O.S_DET_TV], want[:, 0], atol=2e-6)
    assert np.allclose(scores[:, O.S_DET_SD], want[:, 1], atol=2e-6)
    assert np.allclose(scores[:, O.S_DET_C], want[:, 2], atol=2e-6)
    assert np.allclose(scores[:, O.S_DET_AGG], want[:, 3], atol=2e-6)
    assert np.allclose(scores[:, O.S_TV_STD], want[:, 5], atol=2e-6)
    margin = np.abs(want[:, 3] - 0.5) > 1e-5
    assert np.array_equal((flags & O.FLAG_DET_ADV).astype(bool)[margin], want[margin, 4].astype(bool))
    if mode_i != 3:  # 'min' aggregation never crosses 0.5 on this data
        assert want[:, 4].any() and not want[:, 4].all()


def test_defense_detector_consistency_scores_and_dedup():
    """experiments/defenses/detector.py _deduplicate_references + _compute_consistency_scores."""
    z = load("detectors.npz")
    keys = [str(k) for k in z["cs_keys"]]
    scores, flags, (sv, sr, sg) = O.consistency_emb(z["img"], z["txt"], z["var"], ret_rows=z["gallery"],
                                                   ret_idx=z["cand"], gen=z["gen"], g_cnt=z["g_cnt"])
    col = dict(original_similarity=O.S_ORIGINAL, text_variant_consistency=O.S_TV_MEAN, text_variant_std=O.S_TV_STD,
               retrieval_consistency=O.S_RET_MEAN, retrieval_std=O.S_RET_STD,
               generative_consistency=O.S_GEN_MEAN, generative_std=O.S_GEN_STD,
               cross_modal_variance=O.S_CROSS_MODAL_VAR)
    for j, k in enumerate(keys):
        assert np.allclose(scores[:, col[k]], z["cs"][:, j], atol=3e-6), k
    assert np.array_equal(scores[:, O.S_N_RET].astype(np.int64), z["n_ret"])
    # kept reference indices identical (index repeat, exact duplicate row 7 of 3, near duplicate 9 of 4 dropped)
    for i in range(len(z["img"])):
        kept, _ = O.select_refs(z["img"][i], z["gallery"], z["cand"][i], 10, 0.95)
        assert kept == [int(x) for x in z["kept_idx"][i] if x >= 0]
        assert 7 not in kept and 9 not in kept


# ------------------------------------------------------------------------------ hubness
def test_hubness_top1_fraction():
    z = load("hubness.npz")
    for (ni, nq, d) in [(10, 5, 128), (50, 20, 256), (100, 50, 512)]:
        got = O.hubness_top1_fraction(z[f"bench_{ni}_{nq}_{d}_img"], z[f"bench_{ni}_{nq}_{d}_txt"])
        assert got == float(z[f"bench_{ni}_{nq}_{d}_score"])
        assert got > 0


def test_hubness_spec_k_occurrence():
    z = load("hubness.npz")
    counts, hub = O.hubness_spec(z["spec_features"], 10)
    assert np.allclose(hub, z["spec_hubness"], atol=0)
    assert counts.sum() == len(counts) * 10


# ------------------------------------------------------------------------------ oracle self-consistency
def test_topk_tie_rule_and_padding():
    s = np.array([[0.5, 0.9, 0.9, 0.1, 0.9], [0.2, 0.2, 0.2, 0.2, 0.2]], np.float32)
    v, i = O.topk_rows(s, 3)
    assert i.tolist() == [[1, 2, 4], [0, 1, 2]]
    v, i = O.topk_rows(s, 7)
    assert i[0].tolist() == [1, 2, 4, 0, 3, -1, -1] and np.isneginf(v[0, 5:]).all()
    v, i = O.topk_rows(s, 3, threshold=0.6)
    assert i.tolist() == [[1, 2, 4], [-1, -1, -1]]


def test_merge_equals_unsharded():
    rng = np.random.default_rng(0)
    g = O.l2_normalize(rng.standard_normal((1000, 32), dtype=np.float32))
    g[500] = g[10]
    q = O.l2_normalize(rng.standard_normal((40, 32), dtype=np.float32))
    full_s, full_i = O.search(q, g, 10)
    parts_s, parts_i = [], []
    for r in range(4):
        s, i = O.search(q, g[r * 250:(r + 1) * 250], 10, index_offset=r * 250)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = O.merge_topk(np.stack(parts_s, 1), np.stack(parts_i, 1), 10)
    assert np.array_equal(mi, full_i) and np.array_equal(ms, full_s)


def test_bf16_round():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.14159, 1e-30, 0.0], np.float32)
    import torch
    want = torch.from_numpy(x).bfloat16().float().numpy()
    assert np.array_equal(O.bf16_round(x), want)


def test_retrieval_metrics_oracle_matches_reference_evaluator():
    """oracle.retrieval_metrics_from_topk on the full ranking == RetrievalEvaluator.compute_retrieval_metrics
    (src/utils/metrics.py:386-574) run by tests/golden/make_golden.py."""
    z = np.load(GOLD / "retrieval_metrics.npz")
    for tag in ("small", "wide"):
        sims, rel = z[f"{tag}_sims"], z[f"{tag}_rel"]
        ks = [int(k) for k in z[f"{tag}_ks"]]
        order = np.argsort(-sims, axis=1, kind="stable")
        per_q = O.retrieval_metrics_from_topk(order, [list(np.flatnonzero(r)) for r in rel], ks)
        out, nk = per_q.mean(0), len(ks)
        assert np.abs(out[2:2 + nk] - z[f"{tag}_recall"]).max() < 1e-12
        assert np.abs(out[2 + nk:2 + 2 * nk] - z[f"{tag}_precision"]).max() < 1e-12
        assert np.abs(out[2 + 2 * nk:] - z[f"{tag}_ndcg"]).max() < 1e-12
        assert abs(out[0] - float(z[f"{tag}_mrr"])) < 1e-12 and abs(out[1] - float(z[f"{tag}_map"])) < 1e-12
        assert np.abs(per_q[:, :2] - z[f"{tag}_per_query"]).max() < 1e-12
