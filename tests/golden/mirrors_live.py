"""The Python mirrors next to the reference's own classes, on the same random scenarios (build container
only: needs /root/reference; the native layer is tests/fake_native.py, so this pins the HOST logic of the
mirrors - bookkeeping, eviction, persistence, histories, return shapes - not the kernels):

  * ReferenceBank (src/ref_bank.py): insert with de-duplication, fifo / lru / similarity eviction at capacity,
    interleaved query_similar (access counts, LRU order), KMeans clustering, query_by_cluster, statistics,
    the four JSON files written by one implementation and loaded by the other;
  * ConsistencyChecker (experiments/defenses/consistency_checker.py): 60 sequential make_decision calls per
    voting strategy (threshold history), calibrate_threshold, get_statistics;
  * compute_hubness (src/attacks/hubness_attack.py:464-498) and the README k-occurrence definition;
  * RetrievalEvaluator (src/utils/metrics.py:386-574) from ranked lists.

    python tests/golden/mirrors_live.py [seed ...]
"""
from __future__ import annotations

import json
import re
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parents[1]))

import make_golden as MG  # noqa: E402
import fake_native  # noqa: E402

KEYS = ["original_similarity", "text_variant_consistency", "text_variant_std", "retrieval_consistency",
        "retrieval_std", "generative_consistency", "generative_std", "cross_modal_variance"]


def _bank_state(bank):
    return [(r.metadata["i"], r.access_count, r.cluster_id) for r in bank.references]


def check_ref_bank(mods, rng):
    RB = mods["src.ref_bank"]
    from multimodal_detection_consistency_b200 import ref_bank as OB
    d, cap, n_ops = 24, 30, 90
    total = 0
    for strategy, clustering in (("fifo", True), ("lru", False), ("similarity", False)):
        cent = MG.unit(rng, 6, d)
        vecs = (cent[rng.integers(0, 6, n_ops)] * rng.uniform(0.5, 3.0, (n_ops, 1)) +
                0.35 * rng.standard_normal((n_ops, d))).astype(np.float64)
        for j in rng.integers(1, n_ops, 8):                       # near duplicates of earlier vectors: must be refused
            vecs[j] = vecs[rng.integers(0, j)] * 1.7 + 1e-3 * rng.standard_normal(d)
        with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tb:
            banks = []
            for mod, path in ((RB, ta), (OB, tb)):
                cfg = mod.ReferenceBankConfig(max_size=cap, similarity_threshold=0.9, update_strategy=strategy,
                                              persistence_enabled=True, save_path=path, auto_clustering=clustering,
                                              clustering_method="kmeans", num_clusters=4, clustering_interval=10,
                                              feature_dim=d)
                banks.append(mod.ReferenceBank(cfg))
            ref, ours = banks
            for i in range(n_ops):
                a = ref.add_reference(vecs[i], {"i": i})
                b = ours.add_reference(vecs[i], {"i": i})
                assert a == b, (strategy, i, a, b)
                if i % 3 == 2:                                    # interleaved lookups move access counts / LRU order
                    q = vecs[rng.integers(0, i + 1)] + 0.2 * rng.standard_normal(d)
                    k = int(rng.integers(1, 6))
                    ra = ref.query_similar(q, top_k=k, similarity_threshold=0.3)
                    rb = ours.query_similar(q, top_k=k, similarity_threshold=0.3)
                    assert [it.metadata["i"] for it, _ in ra] == [it.metadata["i"] for it, _ in rb], (strategy, i)
                    assert np.allclose([s for _, s in ra], [s for _, s in rb], rtol=0, atol=2e-6)
                    total += len(ra)
                assert _bank_state(ref) == _bank_state(ours), (strategy, i)
                assert list(ref.access_order) == list(ours.access_order), (strategy, i)
            assert len(ref.references) == cap and ref.stats["total_removed"] > 0
            sa, sb = ref.get_statistics(), ours.get_statistics()
            assert sorted(sa) == sorted(sb)
            for key in sa:
                if key != "last_clustering_time":
                    assert sa[key] == sb[key], (strategy, key, sa[key], sb[key])
            if clustering:
                assert ref.stats["clustering_count"] > 0
                assert {int(k): v for k, v in ref.clusters.items()} == {int(k): v for k, v in ours.clusters.items()}
                assert np.allclose(ref.get_cluster_centers(), ours.get_cluster_centers())
                for cid in ref.clusters:
                    assert [r.metadata["i"] for r in ref.query_by_cluster(cid, 5)] == \
                           [r.metadata["i"] for r in ours.query_by_cluster(int(cid), 5)]
            # persistence: same files, and each implementation loads what the other wrote
            ref._save_to_disk()        # the reference only persists inside add_reference: fold the last lookups in
            ours.flush()
            ja = json.loads((Path(ta) / "references.json").read_text())
            jb = json.loads((Path(tb) / "references.json").read_text())
            assert len(ja) == len(jb) == cap
            for x, y in zip(ja, jb):
                assert sorted(x) == sorted(y)
                assert x["metadata"] == y["metadata"] and x["access_count"] == y["access_count"] and \
                    x["cluster_id"] == y["cluster_id"] and np.allclose(x["vector"], y["vector"], rtol=0, atol=0)
            for name in ("clusters.json", "stats.json", "config.json"):
                assert sorted(json.loads((Path(ta) / name).read_text())) == sorted(json.loads((Path(tb) / name).read_text())), name
            cross_a = OB.ReferenceBank(OB.ReferenceBankConfig(max_size=cap, similarity_threshold=0.9, save_path=ta,
                                                              update_strategy=strategy, auto_clustering=False, feature_dim=d))
            cross_b = RB.ReferenceBank(RB.ReferenceBankConfig(max_size=cap, similarity_threshold=0.9, save_path=tb,
                                                              update_strategy=strategy, auto_clustering=False, feature_dim=d))
            assert _bank_state(cross_a) == _bank_state(ref) and _bank_state(cross_b) == _bank_state(ours)
            q = vecs[5] + 0.1 * rng.standard_normal(d)
            assert [it.metadata["i"] for it, _ in cross_a.query_similar(q, 4, 0.2)] == \
                   [it.metadata["i"] for it, _ in cross_b.query_similar(q, 4, 0.2)]
    return total


def _same(a, b, where, tol=2e-6):
    if isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a) == sorted(b), (where, sorted(a), sorted(b))
        for k in a:
            _same(a[k], b[k], f"{where}.{k}", tol)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), (where, a, b)
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{where}[{i}]", tol)
    elif isinstance(a, (bool, np.bool_, str, type(None))):
        assert a == b, (where, a, b)
    else:
        assert abs(float(a) - float(b)) <= tol, (where, a, b)


def check_consistency_checker(mods, rng):
    CC = mods["experiments.defenses.consistency_checker"]
    from multimodal_detection_consistency_b200 import defenses as OD
    n = 60
    S = np.zeros((n, 8))
    S[:, 0] = rng.uniform(-0.2, 1.0, n)
    S[:, 1] = rng.uniform(-0.2, 1.0, n)
    S[:, 2] = rng.uniform(0, 0.5, n)
    S[:, 3] = rng.uniform(0, 1.0, n) * (rng.uniform(size=n) > 0.2)
    S[:, 4] = rng.uniform(0, 0.5, n)
    S[:, 5] = rng.uniform(0, 1.0, n) * (rng.uniform(size=n) > 0.2)
    S[:, 6] = rng.uniform(0, 0.5, n)
    S[:, 7] = rng.uniform(0, 0.3, n) * (rng.uniform(size=n) > 0.3)
    dicts = [{k: float(S[i, j]) for j, k in enumerate(KEYS)} for i in range(n)]
    for voting in ("simple", "weighted", "adaptive"):
        for adaptive in (False, True):
            a = CC.ConsistencyChecker(threshold=0.45, adaptive_threshold=adaptive, voting_strategy=voting)
            b = OD.ConsistencyChecker(threshold=0.45, adaptive_threshold=adaptive, voting_strategy=voting)
            for i, dct in enumerate(dicts):                       # ONE checker each: the threshold history builds up
                ra, rb = a.make_decision(dict(dct)), b.make_decision(dict(dct))
                for key in ("overall_score", "threshold", "confidence"):
                    assert abs(float(ra[key]) - float(rb[key])) <= 2e-6, (voting, adaptive, i, key, ra[key], rb[key])
                if abs(float(ra["overall_score"]) - float(ra["threshold"])) > 1e-5:
                    assert bool(ra["is_adversarial"]) == bool(rb["is_adversarial"]), (voting, adaptive, i)
            assert np.allclose(a.threshold_history, b.threshold_history, rtol=0, atol=2e-6)
            assert OD.ConsistencyChecker().get_statistics() == CC.ConsistencyChecker().get_statistics()
            _same(a.get_statistics(), b.get_statistics(), "get_statistics")
            da = a.make_decision(dict(dicts[3]), return_details=True)["details"]
            db = b.make_decision(dict(dicts[3]), return_details=True)["details"]
            _same(da, db, "details")
            a.update_weights({"original_similarity": 0.4})
            b.update_weights({"original_similarity": 0.4})
            _same(a.make_decision(dict(dicts[4])), b.make_decision(dict(dicts[4])), "after update_weights")
    labels = [bool(x) for x in rng.uniform(size=n) < 0.4]
    a = CC.ConsistencyChecker(voting_strategy="weighted")
    b = OD.ConsistencyChecker(voting_strategy="weighted")
    ta, tb = a.calibrate_threshold([dict(x) for x in dicts], labels), b.calibrate_threshold([dict(x) for x in dicts], labels)
    assert abs(float(ta) - float(tb)) <= 1e-9, (ta, tb)
    return n * 6


def check_hubness(mods, rng):
    H = mods["src.attacks.hubness_attack"]
    from multimodal_detection_consistency_b200 import hubness as OH
    ni, nq, d = int(rng.integers(5, 60)), int(rng.integers(8, 80)), 48
    im = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((ni, d)).astype(np.float32)), dim=1)
    tx = torch.nn.functional.normalize(torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)), dim=1)
    im[0] = torch.nn.functional.normalize(tx[: max(1, nq // 3)].mean(0), dim=0)
    assert OH.compute_hubness(im.numpy(), tx.numpy(), 10) == float(H.HubnessAttack.compute_hubness(None, im, tx, 10))
    md = (MG.REF / "references" / "Adversarial_Hubness_Multi_Modal_Retrieval" / "README.md").read_text()
    block = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "def compute_hubness" in b][0]
    from sklearn.metrics.pairwise import cosine_similarity
    ns = {"np": np, "cosine_similarity": cosine_similarity}
    exec(block, ns)
    n, k = int(rng.integers(60, 300)), int(rng.integers(1, 10))
    f = (MG.unit(rng, 7, d)[rng.integers(0, 7, n)] + 0.1 * rng.standard_normal((n, d))).astype(np.float32)
    counts, scores = OH.hubness_scores(f, k)
    assert np.array_equal(scores, np.asarray(ns["compute_hubness"](f, k=k), np.float64)) and counts.sum() == n * k
    return n


def check_evaluator(mods, rng):
    M = mods["src.utils.metrics"]
    from multimodal_detection_consistency_b200 import metrics as OM
    nq, nc = int(rng.integers(10, 50)), int(rng.integers(60, 64))
    sims = rng.permutation(nq * nc).reshape(nq, nc).astype(np.float64) / (nq * nc)
    rel = (rng.uniform(size=(nq, nc)) < 0.06).astype(np.int64)
    rel[0] = 0
    ks = [1, 5, 10, 20, 50]
    want = M.RetrievalEvaluator.compute_retrieval_metrics(sims, rel, ks)
    order = np.argsort(-sims, axis=1, kind="stable")
    got = OM.RetrievalEvaluator.from_topk(order, rel, ks)
    for k in ks:
        assert abs(want.recall_at_k[k] - got.recall_at_k[k]) < 1e-6 and abs(want.precision_at_k[k] - got.precision_at_k[k]) < 1e-6
        assert abs(want.ndcg_at_k[k] - got.ndcg_at_k[k]) < 1e-6
    assert abs(want.mrr - got.mrr) < 1e-6 and abs(want.map_score - got.map_score) < 1e-6
    assert sorted(want.to_dict()) == sorted(got.to_dict())
    return nq


def main():
    seeds = [int(s) for s in sys.argv[1:]] or [31, 32]
    mods = MG.import_reference()
    with fake_native.installed():
        for seed in seeds:
            rng = np.random.default_rng(seed)
            np.random.seed(seed)
            done = {f.__name__[6:]: f(mods, rng) for f in (check_ref_bank, check_consistency_checker, check_hubness,
                                                           check_evaluator)}
            print(f"seed {seed}: " + ", ".join(f"{k} {v}" for k, v in done.items()))
    print("mirrors live check ok")


if __name__ == "__main__":
    main()
