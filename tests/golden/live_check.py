"""Randomised oracle-vs-reference check (SURVEY.md §8c "reference-in-the-loop"): the reference's own
classes, imported from /root/reference exactly as tests/golden/make_golden.py imports them, are run
side by side with oracle/tvc_oracle.py on FRESH seeded inputs (seeds other than the fixtures').
Build container only; tests/test_oracle_live_reference.py runs this in a subprocess so the stub
modules never enter the test process.

    python tests/golden/live_check.py [seed ...]
"""
from __future__ import annotations

import re
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parents[1]))

import make_golden as MG  # noqa: E402
from oracle import tvc_oracle as O  # noqa: E402

KEYS = ["original_similarity", "text_variant_consistency", "text_variant_std", "retrieval_consistency",
        "retrieval_std", "generative_consistency", "generative_std", "cross_modal_variance"]


def check_ref_bank(mods, rng):
    """ReferenceBank.query_similar (src/ref_bank.py:172-224, 462-484) on un-normalised fp32/fp64 banks."""
    RB = mods["src.ref_bank"]
    n, d = int(rng.integers(20, 300)), int(rng.choice([16, 64, 200]))
    dt = rng.choice([np.float32, np.float64])
    bank_vecs = (rng.standard_normal((n, d)) * rng.uniform(0.2, 4.0)).astype(dt)
    bank_vecs[n // 2] = bank_vecs[1]                     # exact duplicate: a tie the (sim, index) rule must order
    qs = (bank_vecs[rng.integers(0, n, 10)] + rng.uniform(0.3, 2.0) * rng.standard_normal((10, d))).astype(dt)
    qs[0] = bank_vecs[1] * 1.5 + (0.01 * rng.standard_normal(d)).astype(dt)   # sees the duplicated pair at ranks 1, 2
    cfg_thr = float(rng.uniform(0.1, 0.8))
    checked = ties = 0
    with tempfile.TemporaryDirectory() as td:
        cfg = RB.ReferenceBankConfig(max_size=1000, similarity_threshold=cfg_thr, persistence_enabled=False,
                                     save_path=td, auto_clustering=False, feature_dim=d)
        bank = RB.ReferenceBank(cfg)
        for i, v in enumerate(bank_vecs):
            bank.references.append(RB.ReferenceItem(vector=v.copy(), metadata={"i": i}, timestamp=0.0))
        for thr in [None, 0.0, float(rng.uniform(-0.2, 0.9))]:
            for qi, q in enumerate(qs):
                k = int(rng.integers(2 if qi == 0 else 1, 12))
                want = bank.query_similar(q, top_k=k, similarity_threshold=thr)
                idx, sim = O.ref_bank_query(bank_vecs, q, top_k=k, similarity_threshold=thr, config_threshold=cfg_thr)
                w_idx = [it.metadata["i"] for it, _ in want]
                w_sim = np.array([s for _, s in want], np.float64)
                assert len(idx) == len(w_idx), (len(idx), len(w_idx))
                assert np.allclose(sim, w_sim, rtol=0, atol=1e-6)
                # the reference's argsort()[::-1] puts the HIGHER index of a tie first and may cut a tie at k;
                # north_star's rule is lower index first: an index may differ only between exactly tied rows
                all_sims = O.ref_bank_similarities(bank_vecs, q)
                for j in range(len(idx)):
                    if idx[j] != w_idx[j]:
                        assert all_sims[idx[j]] == all_sims[w_idx[j]], (j, idx, w_idx)
                        ties += 1
                order = sorted(range(len(idx)), key=lambda j: (-all_sims[idx[j]], idx[j]))
                assert order == list(range(len(idx))), idx      # (similarity desc, index asc)
                checked += len(idx)
    assert checked > 0 and ties > 0   # the planted duplicate must have exercised the tie rule
    return checked


def check_consistency_checker(mods, rng):
    """ConsistencyChecker.make_decision (experiments/defenses/consistency_checker.py:74-272), all votings,
    adaptive on/off, plus one stateful 30-decision history."""
    CC = mods["experiments.defenses.consistency_checker"]
    n = 120
    S = np.zeros((n, 8))
    S[:, 0] = rng.uniform(-0.3, 1.0, n)
    S[:, 1] = rng.uniform(-0.3, 1.0, n)
    S[:, 2] = rng.uniform(0, 0.6, n)
    S[:, 3] = rng.uniform(-0.1, 1.0, n) * (rng.uniform(size=n) > 0.2)
    S[:, 4] = rng.uniform(0, 0.6, n)
    S[:, 5] = rng.uniform(-0.1, 1.0, n) * (rng.uniform(size=n) > 0.2)
    S[:, 6] = rng.uniform(0, 0.6, n)
    S[:, 7] = rng.uniform(0, 0.3, n) * (rng.uniform(size=n) > 0.3)
    S[0] = 0.0                                            # everything absent
    for vi, voting in enumerate(["simple", "weighted", "adaptive"]):
        for adaptive in (False, True):
            for i in range(n):
                chk = CC.ConsistencyChecker(threshold=0.5, adaptive_threshold=adaptive, voting_strategy=voting)
                r = chk.make_decision({k: float(S[i, j]) for j, k in enumerate(KEYS)})
                got = O.consistency_from_scores(*S[i], dict(voting=vi, cc_adaptive=int(adaptive)))
                want = [float(r["overall_score"]), float(r["threshold"]), float(r["confidence"])]
                assert np.allclose(got[:3], want, rtol=0, atol=1e-12), (voting, adaptive, i, got, want)
                if abs(want[0] - want[1]) > 1e-9:
                    assert bool(got[3]) == bool(r["is_adversarial"])
    chk = CC.ConsistencyChecker(threshold=0.5, adaptive_threshold=True, voting_strategy="adaptive")
    hist = []
    for i in range(30):
        r = chk.make_decision({k: float(S[i, j]) for j, k in enumerate(KEYS)})
        got = O.consistency_from_scores(*S[i], dict(voting=2, cc_adaptive=1), threshold_history=hist)
        hist.append(got[1])
        assert np.allclose(got[:3], [float(r["overall_score"]), float(r["threshold"]), float(r["confidence"])],
                           rtol=0, atol=1e-12), i
    return n * 6 + 30


def check_search(mods, rng):
    """MultiModalRetriever._search_index, sklearn branch (src/retrieval.py:658-674) and
    compute_similarity_matrix (:682-722)."""
    R = mods["src.retrieval"]
    n, d, m = int(rng.integers(200, 1500)), int(rng.choice([32, 128, 512])), 24
    g, q = MG.unit(rng, n, d), MG.unit(rng, m, d)
    k = int(rng.integers(1, 20))
    ret = object.__new__(R.MultiModalRetriever)
    ret.config = R.RetrievalConfig(index_type="exact")
    ret.image_features, ret.text_features = g, q
    s, i = O.search(q, g, k, metric="cosine")
    for r in range(m):
        ii, ss = ret._search_index(None, q[r:r + 1], k)
        assert np.array_equal(i[r], ii), (r, i[r], ii)
        assert np.allclose(s[r], ss, rtol=0, atol=2e-6)
    for metric in ("cosine", "dot_product", "euclidean"):
        ret.config.similarity_metric = metric
        assert np.allclose(O.similarity_matrix(q, g, metric), R.MultiModalRetriever.compute_similarity_matrix(ret),
                           rtol=0, atol=2e-5)
    return m * k


def check_hubness(mods, rng):
    """HubnessAttack.compute_hubness (src/attacks/hubness_attack.py:464-498) and the k-occurrence pseudo-code of
    references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md:30-58."""
    H = mods["src.attacks.hubness_attack"]
    ni, nq, d = int(rng.integers(5, 80)), int(rng.integers(5, 60)), int(rng.choice([64, 256]))
    im = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((ni, d)).astype(np.float32)), dim=1)
    tx = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)), dim=1)
    im[0] = torch.nn.functional.normalize(tx[: max(1, nq // 3)].mean(0), dim=0)
    assert O.hubness_top1_fraction(im.numpy(), tx.numpy()) == float(H.HubnessAttack.compute_hubness(None, im, tx, 10))
    md = (MG.REF / "references" / "Adversarial_Hubness_Multi_Modal_Retrieval" / "README.md").read_text()
    block = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "def compute_hubness" in b][0]
    from sklearn.metrics.pairwise import cosine_similarity
    ns = {"np": np, "cosine_similarity": cosine_similarity}
    exec(block, ns)
    n, k = int(rng.integers(80, 400)), int(rng.integers(1, 12))
    cent = MG.unit(rng, 9, 48)
    f = (cent[rng.integers(0, 9, n)] + 0.1 * rng.standard_normal((n, 48))).astype(np.float32)
    counts, hub = O.hubness_spec(f, k)
    assert np.array_equal(hub, np.asarray(ns["compute_hubness"](f, k=k), np.float64))
    assert counts.sum() == n * k
    return n


def check_metrics(mods, rng):
    """RetrievalEvaluator.compute_retrieval_metrics (src/utils/metrics.py:386-574)."""
    M = mods["src.utils.metrics"]
    nq, nc = int(rng.integers(10, 60)), int(rng.integers(60, 300))
    sims = rng.permutation(nq * nc).reshape(nq, nc).astype(np.float64) / (nq * nc)
    rel = (rng.uniform(size=(nq, nc)) < 0.04).astype(np.int64)
    rel[0] = 0
    ks = [1, 5, 10, 20, 50]
    m = M.RetrievalEvaluator.compute_retrieval_metrics(sims, rel, ks)
    order = np.argsort(-sims, axis=1, kind="stable")
    out = O.retrieval_metrics_from_topk(order, [list(np.flatnonzero(r)) for r in rel], ks).mean(0)
    nk = len(ks)
    assert np.allclose(out[2:2 + nk], [m.recall_at_k[k] for k in ks], rtol=0, atol=1e-12)
    assert np.allclose(out[2 + nk:2 + 2 * nk], [m.precision_at_k[k] for k in ks], rtol=0, atol=1e-12)
    assert np.allclose(out[2 + 2 * nk:], [m.ndcg_at_k[k] for k in ks], rtol=0, atol=1e-12)
    assert abs(out[0] - m.mrr) < 1e-12 and abs(out[1] - m.map_score) < 1e-12
    return nq


def check_detectors(mods, rng):
    """AdversarialDetector.detect_adversarial (src/detector.py:441-682, 4 aggregations) and
    MultiModalDefenseDetector._deduplicate_references/_compute_consistency_scores
    (experiments/defenses/detector.py:228-325), driven by make_golden.gold_detectors with a fresh seed."""
    with tempfile.TemporaryDirectory() as td:
        MG.gold_detectors(mods, Path(td), seed=int(rng.integers(1 << 30)), nq=32)
        z = dict(np.load(Path(td) / "detectors.npz"))
    img, txt, var, gen, g_cnt = z["img"], z["txt"], z["var"], z["gen"], z["g_cnt"]
    for mode_i in range(4):
        want = z["det_scores"][mode_i]
        scores, flags, _ = O.consistency_emb(img, txt, var, gen=gen, g_cnt=g_cnt, params=dict(aggregation=mode_i))
        for col, j in ((O.S_DET_TV, 0), (O.S_DET_SD, 1), (O.S_DET_C, 2), (O.S_DET_AGG, 3), (O.S_TV_STD, 5)):
            assert np.allclose(scores[:, col], want[:, j], rtol=0, atol=2e-6), (mode_i, col)
        margin = np.abs(want[:, 3] - 0.5) > 1e-5
        assert np.array_equal((flags & O.FLAG_DET_ADV).astype(bool)[margin], want[margin, 4].astype(bool))
    scores, _, _ = O.consistency_emb(img, txt, var, ret_rows=z["gallery"], ret_idx=z["cand"], gen=gen, g_cnt=g_cnt)
    col = dict(original_similarity=O.S_ORIGINAL, text_variant_consistency=O.S_TV_MEAN, text_variant_std=O.S_TV_STD,
               retrieval_consistency=O.S_RET_MEAN, retrieval_std=O.S_RET_STD, generative_consistency=O.S_GEN_MEAN,
               generative_std=O.S_GEN_STD, cross_modal_variance=O.S_CROSS_MODAL_VAR)
    for j, k in enumerate(str(k) for k in z["cs_keys"]):
        assert np.allclose(scores[:, col[k]], z["cs"][:, j], rtol=0, atol=3e-6), k
    assert np.array_equal(scores[:, O.S_N_RET].astype(np.int64), z["n_ret"])
    for i in range(len(img)):
        kept, _ = O.select_refs(img[i], z["gallery"], z["cand"][i], 10, 0.95)
        assert kept == [int(x) for x in z["kept_idx"][i] if x >= 0]
    return len(img) * 5


def check_readme_sigma_rule(mods, rng):
    """The documented detection flow, README.md:225-319 (the reference's only executable statement of the
    TVC rule), exec'd from the markdown with table collaborators: its `img_std` over cos(query image, every
    retrieved / generated reference), the `img_std > adaptive_threshold` vote (recommended 0.30, README.md:846)
    and `cross_modal` are the oracle's S_REF_SIGMA, FLAG_SIGMA_ADV and S_ORIGINAL of the similarity-fed mode."""
    import types
    md = (MG.REF / "README.md").read_text()
    block = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "def detect_adversarial(image, text)" in b][0]
    d, n_gal, V, K, M = 48, 120, 3, 3, 1
    cent = MG.unit(rng, 5, d)
    gal = MG.unit(rng, n_gal, d) * 0.6 + cent[rng.integers(0, 5, n_gal)]
    gal = (gal / np.linalg.norm(gal, axis=1, keepdims=True)).astype(np.float32)
    table = {}
    checked = flagged = 0
    for i in range(24):
        text = f"t{i}"
        base = gal[rng.integers(n_gal)]
        table[text] = O.l2_normalize((base + 0.05 * rng.standard_normal(d))[None].astype(np.float32))[0]
        variants = [f"{text}/v{v}" for v in range(V)]
        for vtxt in variants:
            table[vtxt] = O.l2_normalize((table[text] + rng.uniform(0.02, 1.2) * rng.standard_normal(d))[None].astype(np.float32))[0]
        image = O.l2_normalize((base + rng.uniform(0.05, 1.5) * rng.standard_normal(d))[None].astype(np.float32))[0]
        sign = -1.0 if i % 2 else 0.5                                 # odd samples: generated references oppose the image
        sd = {vtxt: [O.l2_normalize((image * sign + 0.1 * rng.standard_normal(d))[None].astype(np.float32))[0] for _ in range(M)]
              for vtxt in variants}

        def retrieve(variant, top_k=5, **_):
            _, idx = O.search(table[variant][None], gal, K)
            return [("gal", int(j)) for j in idx[0]]

        ns = {
            "np": np,
            "text_augmenter": types.SimpleNamespace(generate_variants=lambda t, **kw: list(variants)),
            "retriever": types.SimpleNamespace(retrieve_images_by_text=retrieve),
            "sd_generator": types.SimpleNamespace(generate_reference_images=lambda prompt, **kw: [("sd", prompt, m) for m in range(M)]),
            "clip_model": types.SimpleNamespace(
                encode_image=lambda im: image if isinstance(im, str) else (gal[im[1]] if im[0] == "gal" else sd[im[1]][im[2]]),
                encode_text=lambda t: table[t]),
            "cosine_similarity": lambda a, b: float(O.scalar_cosine(a, b)),
            "adaptive_threshold": 0.30,
            "anomaly_classifier": types.SimpleNamespace(predict=lambda x: [False]),
        }
        exec(block, ns)
        want = ns["detect_adversarial"]("query-image", text)
        # the same references, in the pseudo-code's order, as similarity lists for the oracle
        sr, sg = [], []
        for vtxt in variants:
            sr += [O.scalar_cosine(image, gal[j]) for _, j in retrieve(vtxt)]
            sg += [O.scalar_cosine(image, r) for r in sd[vtxt]]
        s0 = O.scalar_cosine(image, table[text])
        row, flag = O.consistency_one(s0, [], sr, sg, params=dict(n_variants=0, n_retrieval=len(sr), n_generative=len(sg),
                                                                  sigma_threshold=0.30))
        cs = want["consistency_scores"]
        # the pseudo-code interleaves retrieved and generated references per variant; a population std does not
        # depend on the order
        assert abs(row[O.S_REF_SIGMA] - cs["image_std"]) <= 1e-12, (row[O.S_REF_SIGMA], cs["image_std"])
        assert abs(row[O.S_ORIGINAL] - cs["cross_modal"]) <= 1e-12
        n = len(sr) + len(sg)
        assert abs((row[O.S_RET_MEAN] * len(sr) + row[O.S_GEN_MEAN] * len(sg)) / n - cs["image_mean"]) <= 1e-12
        if abs(cs["image_std"] - 0.30) > 1e-9:
            assert bool(flag & O.FLAG_SIGMA_ADV) == bool(want["detection_votes"][0])
        flagged += bool(flag & O.FLAG_SIGMA_ADV)
        checked += 1
    assert 0 < flagged < checked, flagged                              # both sides of the 0.30 vote were exercised
    return checked


def main():
    seeds = [int(s) for s in sys.argv[1:]] or [1001, 1002, 1003]
    mods = MG.import_reference()
    for seed in seeds:
        rng = np.random.default_rng(seed)
        done = {f.__name__[6:]: f(mods, rng) for f in (check_ref_bank, check_consistency_checker, check_search,
                                                       check_hubness, check_metrics, check_detectors,
                                                       check_readme_sigma_rule)}
        print(f"seed {seed}: " + ", ".join(f"{k} {v}" for k, v in done.items()))
    print("live reference check ok")


if __name__ == "__main__":
    main()
