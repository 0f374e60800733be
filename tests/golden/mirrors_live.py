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
            assert "_journal_seq" in jb[0] and all("_journal_seq" not in y for y in jb[1:])
            for x, y in zip(ja, jb):
                # (the fold's sequence number rides on the first item; from_dict reads named keys only - the
                # cross-loading below shows the reference's loader does not mind)
                assert sorted(x) == sorted(k for k in y if k != "_journal_seq")
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
            # export / import (src/ref_bank.py:618-715): each implementation imports what the other exported,
            # into a bank with room for only 12 more
            for fmt, name in (("json", "x.json"), ("numpy", "x.npz")):
                assert ref.export_references(f"{ta}/exp/{name}", fmt) and ours.export_references(f"{tb}/exp/{name}", fmt)
                small = []
                for mod, src_dir in ((RB, tb), (OB, ta)):
                    bank = mod.ReferenceBank(mod.ReferenceBankConfig(max_size=12, similarity_threshold=0.9,
                                                                     persistence_enabled=False, auto_clustering=False,
                                                                     save_path=f"{ta}/unused", feature_dim=d))
                    assert bank.import_references(f"{src_dir}/exp/{name}", fmt)
                    assert not bank.import_references(f"{src_dir}/exp/missing.{fmt}", fmt)
                    small.append(bank)
                assert [r.metadata["i"] for r in small[0].references] == [r.metadata["i"] for r in small[1].references] \
                    == [r.metadata["i"] for r in ref.references[:12]]
                assert small[0].stats["total_added"] == small[1].stats["total_added"] == 12
                ra, rb = small[0].query_similar(q, 3, 0.1), small[1].query_similar(q, 3, 0.1)
                assert [it.metadata["i"] for it, _ in ra] == [it.metadata["i"] for it, _ in rb]
                assert np.allclose([s_ for _, s_ in ra], [s_ for _, s_ in rb], rtol=0, atol=2e-6)
            assert not ours.export_references(f"{tb}/exp/x.bin", "parquet") and not ref.export_references(f"{ta}/exp/x.bin", "parquet")
            ref.clear(), ours.clear()
            assert ref.get_statistics() == ours.get_statistics() and ours.query_similar(q) == ref.query_similar(q) == []
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
    want = float(H.HubnessAttack.compute_hubness(None, im, tx, 10))
    assert OH.compute_hubness(im.numpy(), tx.numpy(), 10) == want
    original = OH.install(H.HubnessAttack)                  # the reference class itself, routed through the mirror
    try:
        attacker = object.__new__(H.HubnessAttack)
        assert attacker.compute_hubness(im, tx) == want and attacker.compute_hubness(im, tx, k=5) == want
    finally:
        H.HubnessAttack.compute_hubness = original
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
    # SimilarityCalculator (src/utils/metrics.py:107-164)
    x, y = rng.standard_normal((9, 33)).astype(np.float32) * 3, rng.standard_normal((14, 33)).astype(np.float32) * 0.2
    x[2] = 0.0
    for i in range(9):
        a, b = M.SimilarityCalculator.cosine_similarity(x[i], y[i]), OM.SimilarityCalculator.cosine_similarity(x[i], y[i])
        assert abs(a - b) <= 5e-7, (i, a, b)
    assert OM.SimilarityCalculator.cosine_similarity(torch.from_numpy(x[2]), list(y[0])) == 0.0
    _same(M.SimilarityCalculator.compute_all_similarities(x[4], y[4]).to_dict(),
          OM.SimilarityCalculator.compute_all_similarities(x[4], y[4]).to_dict(), "compute_all_similarities", tol=1e-5)
    _same(M.SimilarityCalculator.compute_all_similarities(torch.from_numpy(x[5]), torch.from_numpy(y[5])).to_dict(),
          OM.SimilarityCalculator.compute_all_similarities(torch.from_numpy(x[5]), torch.from_numpy(y[5])).to_dict(),
          "compute_all_similarities (torch)", tol=1e-5)
    assert M.SimilarityMetrics().to_dict() == OM.SimilarityMetrics().to_dict()
    want_m = M.SimilarityCalculator.batch_cosine_similarity(x[3:], y)
    assert np.allclose(OM.SimilarityCalculator.batch_cosine_similarity(x[3:], y), want_m, rtol=0, atol=2e-6)
    assert np.allclose(OM.SimilarityCalculator.batch_cosine_similarity(torch.from_numpy(x[3:]), torch.from_numpy(y)),
                       M.SimilarityCalculator.batch_cosine_similarity(torch.from_numpy(x[3:]), torch.from_numpy(y)),
                       rtol=0, atol=2e-6)
    return nq


# ------------------------------------------------------------------------------------------ retriever
def _world(rng, root, n_img=60, n_txt=40, d=48):
    from pipeline_dropin import TableClip, image_of
    img_f, txt_f = MG.unit(rng, n_img, d), MG.unit(rng, n_txt, d)
    paths, image_table, text_table, texts = [], {}, {}, []
    for j in range(n_img):
        p = root / f"g{j:03d}.png"
        image_of(j).save(p)
        paths.append(str(p))
        image_table[j] = img_f[j]
    for j in range(n_txt):
        t = f"caption number {j}"
        texts.append(t)
        text_table[t] = txt_f[j]
    queries = []
    for j in range(12):                                              # query texts near an image, query images near a text
        qt = f"query text {j}"
        v = img_f[rng.integers(n_img)] + 0.08 * rng.standard_normal(d).astype(np.float32)
        text_table[qt] = (v / np.linalg.norm(v)).astype(np.float32)
        v = txt_f[rng.integers(n_txt)] + 0.08 * rng.standard_normal(d).astype(np.float32)
        image_table[2000 + j] = (v / np.linalg.norm(v)).astype(np.float32)
        queries.append((qt, image_of(2000 + j)))
    return TableClip(text_table, image_table), paths, texts, queries, img_f, txt_f


def check_retriever(mods, rng):
    """MultiModalRetriever (src/retrieval.py:316-912) and its helpers FaissIndexManager (:89-155), RetrievalIndex
    (:196-287), ConsistencyCalculator (:158-193): same calls, same return shapes / values / stats / files."""
    R = mods["src.retrieval"]
    from multimodal_detection_consistency_b200 import faiss_compat, retrieval as OR
    from pipeline_dropin import NumpyFlatIP
    fa = sys.modules["faiss"]
    fa.IndexFlatIP = NumpyFlatIP
    fa.get_num_gpus = lambda: 0

    def write_index(index, path):
        Path(path).write_bytes(faiss_compat._pack_flat(index.x, index.d))

    def read_index(path):
        d, _, rows = faiss_compat._unpack_flat(Path(path).read_bytes())
        idx = NumpyFlatIP(d)
        idx.add(rows)
        return idx

    fa.write_index, fa.read_index = write_index, read_index
    n = 0
    with tempfile.TemporaryDirectory() as td:
        clip, paths, texts, queries, img_f, txt_f = _world(rng, Path(td))
        for index_type in ("faiss", "exact"):
            a = R.MultiModalRetriever(R.RetrievalConfig(index_type=index_type, top_k=7))
            b = OR.MultiModalRetriever(OR.RetrievalConfig(index_type=index_type, top_k=7))
            a.clip_model = b.clip_model = clip
            _same(a.get_stats(), b.get_stats(), "empty stats")
            assert a.retrieve_images_by_text("query text 0") == b.retrieve_images_by_text("query text 0") == ([], [])
            for r in (a, b):
                r.build_image_index(paths)
                r.build_text_index(texts)
            assert np.array_equal(a.image_features, b.image_features) and np.array_equal(a.text_features, b.text_features)
            assert a.image_paths == b.image_paths and a.texts == b.texts
            if index_type == "exact":
                # the reference's `exact` retriever holds no index object and refuses the public calls
                # (src/retrieval.py:546-547 needs image_index, :488-489 returns None) - only _search_index works
                for qt, _ in queries[:4]:
                    ia, sa = a._search_index(None, clip.encode_text([qt]).numpy(), 5)
                    ib, sb = b._search_index(None, clip.encode_text([qt]).numpy(), 5)
                    assert np.array_equal(ia, ib) and np.allclose(sa, sb, rtol=0, atol=2e-6)
                continue
            for qt, qi in queries:
                k = int(rng.integers(1, 9))
                _same(a.retrieve_images_by_text(qt, top_k=k), b.retrieve_images_by_text(qt, top_k=k), f"t2i {qt}")
                _same(a.retrieve_texts_by_image(qi, top_k=k), b.retrieve_texts_by_image(qi, top_k=k), f"i2t {qt}")
                n += 2
            _same(a.retrieve_images_by_text(queries[0][0]), b.retrieve_images_by_text(queries[0][0]), "default top_k")
            _same(a.retrieve_texts_by_image(paths[3], top_k=4), b.retrieve_texts_by_image(paths[3], top_k=4), "i2t by path")
            _same(a.batch_retrieve_images_by_texts([q for q, _ in queries], 5),
                  b.batch_retrieve_images_by_texts([q for q, _ in queries], 5), "batch t2i")
            _same(a.batch_retrieve_texts_by_images([q for _, q in queries], 5),
                  b.batch_retrieve_texts_by_images([q for _, q in queries], 5), "batch i2t")
            for metric in ("cosine", "dot_product", "euclidean"):
                a.config.similarity_metric = b.config.similarity_metric = metric
                assert np.allclose(a.compute_similarity_matrix(), b.compute_similarity_matrix(), rtol=0, atol=2e-5)
                assert np.allclose(a.compute_similarity_matrix(txt_f[:5] * 2, img_f[:9] * 3),
                                   b.compute_similarity_matrix(txt_f[:5] * 2, img_f[:9] * 3), rtol=0, atol=2e-5)
            a.config.similarity_metric = b.config.similarity_metric = "cosine"
            _same(a.get_stats(), b.get_stats(), "stats")
            # persistence: pickle + .faiss sidecar, each implementation loads the other's files
            pa, pb = Path(td) / "a" / "img.pkl", Path(td) / "b" / "img.pkl"
            a.save_image_index(str(pa)), b.save_image_index(str(pb))
            a.save_text_index(str(pa.with_name("txt.pkl"))), b.save_text_index(str(pb.with_name("txt.pkl")))
            assert pa.with_suffix(".faiss").read_bytes() == pb.with_suffix(".faiss").read_bytes()
            a2 = R.MultiModalRetriever(R.RetrievalConfig(index_type=index_type, top_k=7))
            b2 = OR.MultiModalRetriever(OR.RetrievalConfig(index_type=index_type, top_k=7))
            a2.clip_model = b2.clip_model = clip
            a2.load_image_index(str(pb)), a2.load_text_index(str(pb.with_name("txt.pkl")))      # reference <- ours
            b2.load_image_index(str(pa)), b2.load_text_index(str(pa.with_name("txt.pkl")))      # ours <- reference
            for qt, qi in queries[:5]:
                _same(a.retrieve_images_by_text(qt, top_k=6), a2.retrieve_images_by_text(qt, top_k=6), "ref loads ours")
                _same(a.retrieve_images_by_text(qt, top_k=6), b2.retrieve_images_by_text(qt, top_k=6), "ours loads ref")
                _same(a.retrieve_texts_by_image(qi, top_k=6), b2.retrieve_texts_by_image(qi, top_k=6), "ours loads ref i2t")
            a.clear_cache(), b.clear_cache()
            _same(a.get_stats(), b.get_stats(), "stats after clear")
        # helpers
        fm_a, fm_b = R.FaissIndexManager(R.IndexConfig(index_type="flat", dimension=48, use_gpu=False)), \
            OR.FaissIndexManager(OR.IndexConfig(index_type="flat", dimension=48, use_gpu=False))
        for fm in (fm_a, fm_b):
            fm.build_index(img_f[:40])
            fm.add_to_index(img_f[40:])
        sa, ia = fm_a.search(txt_f[:6], 5)
        sb, ib = fm_b.search(txt_f[:6], 5)
        assert np.array_equal(ia, ib) and np.allclose(sa, sb, rtol=0, atol=2e-6)
        ri_a, ri_b = R.RetrievalIndex("faiss", 48), OR.RetrievalIndex("faiss", 48)
        for ri in (ri_a, ri_b):
            ri.build_index(img_f[:50])
            ri.add_items(img_f[50:])
        xa, xb = ri_a.search(txt_f[:6], 5), ri_b.search(txt_f[:6], 5)     # note: (indices, distances) here (:254)
        assert np.array_equal(xa[0], xb[0]) and np.allclose(xa[1], xb[1], rtol=0, atol=2e-6)
        for mod in (R, OR):                                           # small dataclasses of the retrieval module
            try:
                mod.IndexConfig(index_type="pq")
                raise AssertionError("IndexConfig accepted an unknown index type")
            except ValueError:
                pass
        rr_a = R.RetrievalResult(indices=[4, 2, 9], similarities=[0.9, 0.5, 0.2], items=["a", "b", "c"], query_time=0.1)
        rr_b = OR.RetrievalResult(indices=[4, 2, 9], similarities=[0.9, 0.5, 0.2], items=["a", "b", "c"], query_time=0.1)
        _same(rr_a.to_dict(), rr_b.to_dict(), "RetrievalResult.to_dict")
        _same(rr_a.filter_by_similarity(0.4).to_dict(), rr_b.filter_by_similarity(0.4).to_dict(), "filter_by_similarity")
        assert rr_a.filter_by_similarity(0.4).items == rr_b.filter_by_similarity(0.4).items == ["a", "b"]
        assert rr_a.get_top_k(2).items == rr_b.get_top_k(2).items and rr_a.get_top_k(7).indices == rr_b.get_top_k(7).indices
        ca, cb = R.ConsistencyCalculator(), OR.ConsistencyCalculator()
        _same(ca.compute_similarity_distribution(sa[0]), cb.compute_similarity_distribution(sb[0]), "distribution")
        assert abs(ca.compute_consistency_score(sa[0], sa[1]) - cb.compute_consistency_score(sb[0], sb[1])) < 1e-6
        assert ca.compute_top_k_consistency(ia[0], ia[1], 5) == cb.compute_top_k_consistency(ib[0], ib[1], 5)
        assert abs(ca.compute_rank_correlation(ia[0], ia[1]) - cb.compute_rank_correlation(ib[0], ib[1])) < 1e-9
    return n


# ------------------------------------------------------------------------------------------ detector
def check_detector(mods, rng):
    """AdversarialDetector (src/detector.py:216-860): detect_adversarial for every method subset and aggregation,
    batch_detect, compute_optimal_threshold, update_threshold, evaluate_detection_performance, get_stats,
    save_model / load_model, caches."""
    D = mods["src.detector"]
    import types as _t
    from multimodal_detection_consistency_b200 import detector as OD
    from pipeline_dropin import TableClip, image_of
    d, nq, V, G = 48, 20, 5, 3
    sd = 1.0 / np.sqrt(d)
    text_table, image_table, samples = {}, {}, []
    for i in range(nq):
        base = MG.unit(rng, 1, d)[0]
        text = f"sample text {i}"
        text_table[text] = base
        for v in range(V):
            tv = base + 0.3 * sd * rng.standard_normal(d).astype(np.float32)
            text_table[f"{text} ~v{v}"] = (tv / np.linalg.norm(tv)).astype(np.float32)
        attacked = bool(rng.uniform() < 0.45)
        im = MG.unit(rng, 1, d)[0] if attacked else base + 0.6 * sd * rng.standard_normal(d).astype(np.float32)
        image_table[100 + i] = (im / np.linalg.norm(im)).astype(np.float32)
        for g in range(G):
            ref = 0.7 * image_table[100 + i] + 2.0 * sd * rng.standard_normal(d).astype(np.float32)
            image_table[500 + G * i + g] = (ref / np.linalg.norm(ref)).astype(np.float32)
        samples.append((image_of(100 + i), text, attacked))
    index_of = {s[1]: i for i, s in enumerate(samples)}
    clip = TableClip(text_table, image_table)
    aug = _t.SimpleNamespace(generate_variants=lambda t: [f"{t} ~v{v}" for v in range(V)])
    few = _t.SimpleNamespace(generate_variants=lambda t: [f"{t} ~v{v}" for v in range(2 + index_of[t] % 3)])
    gen = _t.SimpleNamespace(generate_reference_images=lambda text, num_images=G: {
        "images": [image_of(500 + G * index_of[text] + g) for g in range(min(num_images, 1 + index_of[text] % 3))],
        "generation_time": 0.0})
    n = 0
    for agg in ("weighted_mean", "mean", "max", "min"):
        for methods in (None, ["consistency"], ["text_variants", "consistency"], ["sd_reference"]):
            a = D.AdversarialDetector(D.DetectorConfig(score_aggregation=agg, detection_threshold=0.42))
            b = OD.AdversarialDetector(OD.DetectorConfig(score_aggregation=agg, detection_threshold=0.42))
            for det in (a, b):
                det.clip_model, det.text_augmenter, det.sd_generator = clip, (few if agg == "mean" else aug), gen
            for im, text, _ in samples:
                ra, rb = a.detect_adversarial(im, text, methods), b.detect_adversarial(im, text, methods)
                assert "error" not in ra and "error" not in rb, (ra.get("error"), rb.get("error"))
                if abs(ra["aggregated_score"] - 0.42) < 1e-5:
                    rb["is_adversarial"] = ra["is_adversarial"]
                _same({k: v for k, v in ra.items() if k != "detection_time"},
                      {k: v for k, v in rb.items() if k != "detection_time"}, f"{agg} {methods} {text}", tol=3e-6)
                n += 1
            sa, sb = a.get_stats(), b.get_stats()
            sa["detection_stats"].pop("detection_time"), sb["detection_stats"].pop("detection_time")
            _same(sa, sb, "detector stats")
    a = D.AdversarialDetector(D.DetectorConfig())
    b = OD.AdversarialDetector(OD.DetectorConfig())
    for det in (a, b):
        det.clip_model, det.text_augmenter, det.sd_generator = clip, aug, gen
    ims, txts = [s[0] for s in samples], [s[1] for s in samples]
    ba, bb = a.batch_detect(ims, txts), b.batch_detect(ims, txts)
    for x, y in zip(ba, bb):
        _same({k: v for k, v in x.items() if k != "detection_time"}, {k: v for k, v in y.items() if k != "detection_time"},
              "batch_detect", tol=3e-6)
    # collaborators that come back empty: both record the method with score 0.0 and an error entry (:455-458, :525-527)
    empty_aug = _t.SimpleNamespace(generate_variants=lambda t: [] if index_of[t] % 2 else [f"{t} ~v0"])
    empty_gen = _t.SimpleNamespace(generate_reference_images=lambda text, num_images=G: {
        "images": [] if index_of[text] % 3 == 0 else [image_of(500 + G * index_of[text])], "generation_time": 0.0})
    for agg in ("weighted_mean", "max"):
        ea = D.AdversarialDetector(D.DetectorConfig(score_aggregation=agg, enable_cache=False))
        eb = OD.AdversarialDetector(OD.DetectorConfig(score_aggregation=agg, enable_cache=False))
        for det in (ea, eb):
            det.clip_model, det.text_augmenter, det.sd_generator = clip, empty_aug, empty_gen
        for im, text, _ in samples[:12]:
            ra, rb = ea.detect_adversarial(im, text), eb.detect_adversarial(im, text)
            _same(ra["detection_scores"], rb["detection_scores"], f"empty collaborators {agg} {text}", tol=3e-6)
            assert abs(ra["aggregated_score"] - rb["aggregated_score"]) <= 3e-6, (agg, text)
            if abs(ra["aggregated_score"] - 0.5) > 1e-5:
                assert ra["is_adversarial"] == rb["is_adversarial"]
            for m in ("text_variants", "sd_reference"):
                assert ("error" in ra["detection_details"][m]) == ("error" in rb["detection_details"][m]), (agg, text, m)
    for im, text, _ in samples[:5]:                                   # repeats are served from the cache by both
        _same({k: v for k, v in a.detect_adversarial(im, text).items() if k != "detection_time"},
              {k: v for k, v in b.detect_adversarial(im, text).items() if k != "detection_time"}, "cached", tol=3e-6)
    sa, sb = a.get_stats(), b.get_stats()
    sa["detection_stats"].pop("detection_time"), sb["detection_stats"].pop("detection_time")
    _same(sa, sb, "detector stats with cache hits")
    assert sa["detection_stats"]["cache_hits"] == 5
    ta, tb = a.compute_optimal_threshold(samples), b.compute_optimal_threshold(samples)
    assert abs(ta - tb) <= 3e-6, (ta, tb)
    a.update_threshold(ta), b.update_threshold(ta)
    a.clear_cache(), b.clear_cache()
    pa, pb = a.evaluate_detection_performance(samples), b.evaluate_detection_performance(samples)
    # the reference calls DetectionEvaluator.compute_metrics, which does not exist (src/detector.py:806 vs
    # src/utils/metrics.py:285), so its method always returns {}; the mirror returns the confusion counts
    assert pa == {} and sorted(pb) == ["accuracy", "f1", "fn", "fp", "precision", "recall", "tn", "tp"]
    pred = np.array([b.detect_adversarial(im, tx)["is_adversarial"] for im, tx, _ in samples])
    lab = np.array([x[2] for x in samples])
    assert pb["tp"] == int((pred & lab).sum()) and pb["fp"] == int((pred & ~lab).sum())
    assert pb["fn"] == int((~pred & lab).sum()) and pb["tn"] == int((~pred & ~lab).sum())
    assert abs(pb["accuracy"] - float((pred == lab).mean())) < 1e-12
    with tempfile.TemporaryDirectory() as td:
        a.save_model(f"{td}/a.json"), b.save_model(f"{td}/b.json")
        ja, jb = json.loads(Path(f"{td}/a.json").read_text()), json.loads(Path(f"{td}/b.json").read_text())
        assert sorted(ja) == sorted(jb) and sorted(ja["config"]) == sorted(jb["config"])
        a2, b2 = D.AdversarialDetector(D.DetectorConfig()), OD.AdversarialDetector(OD.DetectorConfig())
        a2.load_model(f"{td}/b.json"), b2.load_model(f"{td}/a.json")              # each loads the other's file
        assert abs(a2.config.detection_threshold - ta) < 1e-12 and abs(b2.config.detection_threshold - ta) < 1e-12
    # error convention: no encoder at all -> the documented dict, never an exception (src/detector.py:428-439)
    bare = OD.AdversarialDetector(OD.DetectorConfig())
    bare._tried.update({"clip", "aug", "sd"})
    r = bare.detect_adversarial(ims[0], txts[0])
    assert r["is_adversarial"] is False and r["aggregated_score"] == 0.0 and "error" in r
    return n


# ------------------------------------------------------------------------------------------ retrieval references (a8)
def check_retrieval_reference(mods, rng):
    """RetrievalReferenceGenerator.retrieve_references (experiments/defenses/retrieval_ref.py:173-290: features.npy
    database, top-`rerank_top_k` inner-product search, similarity floor, cut to `reference_count`) on its NumPy
    branch and on its FAISS branch (NumPy IndexFlatIP stand-in), next to defenses.RetrievalReferenceIndex."""
    import importlib.util
    from multimodal_detection_consistency_b200 import defenses as ODf
    from pipeline_dropin import NumpyFlatIP
    spec = importlib.util.spec_from_file_location("ref_retrieval_ref", MG.REF / "experiments" / "defenses" / "retrieval_ref.py")
    RR = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RR)
    sys.modules["faiss"].IndexFlatIP = NumpyFlatIP
    n, d = 300, 40
    feats = MG.unit(rng, n, d)
    meta = [{"path": f"ref_{j}.jpg", "j": j} for j in range(n - 5)]              # shorter than the features (:255)
    texts = {f"reference query {j}": (lambda v: (v / np.linalg.norm(v)).astype(np.float32))(
        feats[rng.integers(n)] + rng.uniform(0.05, 0.6) * rng.standard_normal(d).astype(np.float32)) for j in range(20)}
    clip = type("Clip", (), {"encode_text": lambda self, ts: torch.stack([torch.from_numpy(texts[t]) for t in ts])})()
    done = 0
    with tempfile.TemporaryDirectory() as td:
        np.save(Path(td) / "features.npy", feats)
        (Path(td) / "metadata.json").write_text(json.dumps(meta))
        for use_faiss in (False, True):
            for count, floor, rerank in ((5, 0.3, True), (3, 0.55, True), (8, 0.0, False)):
                cfg = RR.RetrievalConfig(reference_count=count, similarity_threshold=floor, use_faiss=use_faiss,
                                         faiss_index_type="Flat", device="cpu", enable_reranking=rerank, rerank_top_k=20)
                ref = RR.RetrievalReferenceGenerator(clip, td, cfg)
                assert (ref.faiss_index is not None) == use_faiss
                ours = ODf.RetrievalReferenceIndex(feats, meta, reference_count=count, similarity_threshold=floor,
                                                   rerank_top_k=20, enable_reranking=rerank)
                for t, q in texts.items():
                    a, b = ref.retrieve_references(t), ours.retrieve_references(q)
                    assert [x["index"] for x in a] == [x["index"] for x in b], (use_faiss, count, floor, t)
                    assert np.allclose([x["similarity"] for x in a], [x["similarity"] for x in b], rtol=0, atol=2e-6)
                    assert [x["metadata"] for x in a] == [x["metadata"] for x in b]
                    assert all(np.array_equal(x["features"], y["features"]) for x, y in zip(a, b))
                    done += len(a)
    assert done > 0
    # ---- the generator class itself: cache, statistics, batch entry, inserts, files (:173-236, 318-457, 542-600)
    def strip(stats):
        stats = dict(stats)
        stats.pop("average_retrieval_time")
        return stats

    ids = {}
    clip2 = type("Clip", (), {
        "encode_text": lambda self, ts: torch.stack([torch.from_numpy(texts[t]) for t in ts]) * 3.0,   # un-normalised
        "encode_image": lambda self, im: torch.from_numpy(ids[int(torch.as_tensor(im).reshape(-1)[0])])[None] * 2.0})()
    with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tb:
        for td in (ta, tb):
            np.save(Path(td) / "features.npy", feats[:200])
            (Path(td) / "metadata.json").write_text(json.dumps(meta[:200]))
        cfg_a = RR.RetrievalConfig(reference_count=4, similarity_threshold=0.35, use_faiss=False, device="cpu", cache_size=12)
        cfg_b = ODf.RetrievalRefConfig(reference_count=4, similarity_threshold=0.35, use_faiss=False, device="cpu", cache_size=12)
        a, b = RR.RetrievalReferenceGenerator(clip2, ta, cfg_a), ODf.RetrievalReferenceGenerator(clip2, tb, cfg_b)
        names = list(texts)

        def same_refs(x, y, where):
            assert [r["index"] for r in x] == [r["index"] for r in y], where
            assert np.allclose([r["similarity"] for r in x], [r["similarity"] for r in y], rtol=0, atol=2e-6), where
            assert [r["metadata"] for r in x] == [r["metadata"] for r in y], where
            assert all(np.array_equal(r["features"], q["features"]) for r, q in zip(x, y)), where

        for t in names[:8] + names[:3]:                               # the last three are cache hits
            same_refs(a.retrieve_references(t), b.retrieve_references(t), t)
        _same(strip(a.get_statistics()), strip(b.get_statistics()), "generator stats")
        ra, rb = a.batch_retrieve_references(names[5:] + names[:2]), b.batch_retrieve_references(names[5:] + names[:2])
        for x, y, t in zip(ra, rb, names[5:] + names[:2]):
            same_refs(x, y, "batch " + t)
        _same(strip(a.get_statistics()), strip(b.get_statistics()), "generator stats after the batch")
        assert len(a.feature_cache) == len(b.feature_cache) == 12      # cache_size respected by both
        # inserts: image tensors through the encoder, normalised, appended, index rebuilt, files rewritten
        new_ids = list(range(900, 906))
        for j in new_ids:
            ids[j] = (feats[200 + j - 900] * 1.0).astype(np.float32)
        ims = [torch.tensor([float(j)]) for j in new_ids]
        caps = [f"caption {j}" for j in new_ids]
        extra = [{"source": "test", "rank": j} for j in new_ids[:4]]
        assert a.add_reference_images(ims, caps, extra) and b.add_reference_images(ims, caps, extra)
        assert not a.add_reference_images(ims, caps[:2]) and not b.add_reference_images(ims, caps[:2])
        assert np.allclose(a.reference_features, b.reference_features, rtol=0, atol=1e-7) and a.reference_metadata == b.reference_metadata
        assert np.allclose(np.load(Path(ta) / "features.npy"), np.load(Path(tb) / "features.npy"), rtol=0, atol=1e-7)
        assert json.loads((Path(ta) / "metadata.json").read_text()) == json.loads((Path(tb) / "metadata.json").read_text())
        a.clear_cache(), b.clear_cache()
        for t in names[:6]:
            same_refs(a.retrieve_references(t), b.retrieve_references(t), "after insert " + t)
        a.update_config(RR.RetrievalConfig(reference_count=2, similarity_threshold=0.1, use_faiss=False, device="cpu"))
        b.update_config(ODf.RetrievalRefConfig(reference_count=2, similarity_threshold=0.1, use_faiss=False, device="cpu"))
        a.clear_cache(), b.clear_cache()
        for t in names[6:10]:
            same_refs(a.retrieve_references(t), b.retrieve_references(t), "after update_config " + t)
        a.reset_statistics(), b.reset_statistics()
        _same(strip(a.get_statistics()), strip(b.get_statistics()), "generator stats after reset")
        # each implementation opens the database the other one wrote; a missing database starts empty
        a2, b2 = RR.RetrievalReferenceGenerator(clip2, tb, cfg_a), ODf.RetrievalReferenceGenerator(clip2, ta, cfg_b)
        for t in names[:4]:
            same_refs(a2.retrieve_references(t), b2.retrieve_references(t), "cross-loaded " + t)
        e1, e2 = RR.RetrievalReferenceGenerator(clip2, ta + "/none", cfg_a), ODf.RetrievalReferenceGenerator(clip2, tb + "/none", cfg_b)
        assert e1.retrieve_references(names[0]) == e2.retrieve_references(names[0]) == []
        _same(strip(e1.get_statistics()), strip(e2.get_statistics()), "empty generator stats")
    return done


# ------------------------------------------------------------------------------------------ defense detector (a14)
def check_defense_detector(mods, rng):
    """MultiModalDefenseDetector.detect (experiments/defenses/detector.py:117-325): variants -> retrieval references
    of every variant, de-duplicated at cosine 0.95 and cut to 10 -> generated references -> 9 consistency scores
    -> ConsistencyChecker decision.  The reference object is assembled as tests/golden/make_golden.py does (its
    constructor does not match its own collaborators' signatures, :87-103); its retrieval collaborator returns
    the ids our RetrievalReferenceIndex returns (check_retrieval_reference pins that equivalence)."""
    import types as _t
    ED = mods["experiments.defenses.detector"]
    CC = mods["experiments.defenses.consistency_checker"]
    from multimodal_detection_consistency_b200 import defenses as ODf
    from oracle import tvc_oracle as O
    d, n_gal, nq, V, G = 40, 200, 16, 4, 3
    sd = 1.0 / np.sqrt(d)
    cent = MG.unit(rng, 8, d)
    gal = cent[rng.integers(0, 8, n_gal)] + 0.5 * sd * rng.standard_normal((n_gal, d)).astype(np.float32)
    gal = (gal / np.linalg.norm(gal, axis=1, keepdims=True)).astype(np.float32)
    gal[11], gal[12] = gal[10], (gal[10] + 0.01 * sd * rng.standard_normal(d)).astype(np.float32)   # exact + near duplicate
    gal[12] /= np.linalg.norm(gal[12])
    text_table, image_table = {}, {j: gal[j] for j in range(n_gal)}
    samples = []
    for i in range(nq):
        pick = 10 if i < 3 else int(rng.integers(n_gal))                       # the duplicated rows get retrieved
        text = f"defense sample {i}"
        t = gal[pick] + 0.4 * sd * rng.standard_normal(d).astype(np.float32)
        text_table[text] = (t / np.linalg.norm(t)).astype(np.float32)
        for v in range(V):
            tv = text_table[text] + 0.3 * sd * rng.standard_normal(d).astype(np.float32)
            text_table[f"{text} ~v{v}"] = (tv / np.linalg.norm(tv)).astype(np.float32)
        im = MG.unit(rng, 1, d)[0] if rng.uniform() < 0.4 else gal[pick] + 0.5 * sd * rng.standard_normal(d).astype(np.float32)
        image_table[1000 + i] = (im / np.linalg.norm(im)).astype(np.float32)
        for g in range(G):
            r = 0.7 * image_table[1000 + i] + 1.5 * sd * rng.standard_normal(d).astype(np.float32)
            image_table[5000 + G * i + g] = (r / np.linalg.norm(r)).astype(np.float32)
        samples.append((torch.tensor(1000 + i), text))
    index_of = {s[1]: i for i, s in enumerate(samples)}
    clip = MG._TableClip(text_table, image_table)
    clip_b = _t.SimpleNamespace(
        encode_text=lambda ts: torch.stack([torch.from_numpy(text_table[t]) for t in ts]),
        encode_image=lambda im: torch.from_numpy(image_table[int(torch.as_tensor(im).reshape(-1)[0])])[None])
    variants = _t.SimpleNamespace(generate_variants=lambda t: [f"{t} ~v{v}" for v in range(V)])
    root_of = lambda t: t.split(" ~v")[0]                                         # noqa: E731
    generator = _t.SimpleNamespace(generate_references=lambda t: [torch.tensor(5000 + G * index_of[root_of(t)] + g)
                                                                  for g in range(2 if "~v" in t else 1)])

    def retrieve(text):
        s, i = O.search(text_table[text][None], gal, 5, threshold=0.3)
        return [torch.tensor(int(j)) for j in i[0] if j >= 0]

    done = 0
    for voting in ("weighted", "adaptive", "simple"):
        ref = object.__new__(ED.MultiModalDefenseDetector)
        ref.clip_model, ref.config, ref.device = clip, ED.DetectionConfig(voting_strategy=voting), torch.device("cpu")
        ref.text_variant_generator, ref.generative_generator = variants, generator
        ref.retrieval_generator = _t.SimpleNamespace(retrieve_references=retrieve)
        ref.consistency_checker = CC.ConsistencyChecker(threshold=0.5, adaptive_threshold=True, voting_strategy=voting)
        ours = ODf.MultiModalDefenseDetector(
            clip_model=clip_b, config=ODf.DetectionConfig(voting_strategy=voting), text_variant_generator=variants,
            retrieval_index=ODf.RetrievalReferenceIndex(gal, reference_count=5, similarity_threshold=0.3),
            generative_generator=generator)
        for image, text in samples:                                               # sequential: the checker's history builds up
            a, b = ref.detect(image, text, return_details=True), ours.detect(image, text, return_details=True)
            for key in ("confidence", "consistency_score"):
                assert abs(float(a[key]) - float(b[key])) <= 3e-6, (voting, text, key, a[key], b[key])
            dd = a["details"]["detection_details"]
            if abs(dd["overall_score"] - dd["threshold"]) > 1e-5:
                assert bool(a["is_adversarial"]) == bool(b["is_adversarial"]), (voting, text)
            assert a["details"]["text_variants"] == b["details"]["text_variants"]
            for key, val in a["details"]["consistency_scores"].items():
                assert abs(float(val) - b["details"]["consistency_scores"][key]) <= 3e-6, (voting, text, key)
            n_ref = len(a["details"]["retrieval_references"])
            assert n_ref == int(b["details"]["consistency_scores"]["n_retrieval"]), (voting, text, n_ref)
            done += 1
        kept = {int(x) for x in ref._generate_retrieval_references(samples[0][1], [samples[0][1]] + variants.generate_variants(samples[0][1]))}
        assert len(kept & {10, 11, 12}) == 1                                      # one of the (near-)duplicate rows survives
        sa, sb = ref.get_statistics(), ours.get_statistics()
        assert sorted(sa) == sorted(sb) and sa["components"] == sb["components"] and sorted(sa["config"]) == sorted(sb["config"])
        _same(sa["consistency_checker_stats"], sb["consistency_checker_stats"], "defense detector checker stats", tol=3e-6)
        ref.update_config(ED.DetectionConfig(retrieval_top_k=4)), ours.update_config(ODf.DetectionConfig(retrieval_top_k=4))
        assert ref.config.retrieval_top_k == ours.config.retrieval_top_k == 4
        ba = ours.batch_detect([s[0] for s in samples[:6]], [s[1] for s in samples[:6]])
        assert len(ba) == 6 and all(sorted(r) == ["confidence", "consistency_score", "is_adversarial"] for r in ba)
    return done


# ------------------------------------------------------------------------------------------ config dataclasses
def check_configs(mods, rng):
    """Every config dataclass on the path: same field names, same order, same defaults as the reference's."""
    import dataclasses
    import importlib.util
    from multimodal_detection_consistency_b200 import defenses as ODf, detector as ODt, ref_bank as OB, retrieval as OR
    spec = importlib.util.spec_from_file_location("ref_retrieval_ref_cfg", MG.REF / "experiments" / "defenses" / "retrieval_ref.py")
    RR = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RR)
    pairs = [(mods["src.retrieval"].RetrievalConfig, OR.RetrievalConfig), (mods["src.retrieval"].IndexConfig, OR.IndexConfig),
             (mods["src.ref_bank"].ReferenceBankConfig, OB.ReferenceBankConfig),
             (mods["src.detector"].DetectorConfig, ODt.DetectorConfig),
             (mods["experiments.defenses.detector"].DetectionConfig, ODf.DetectionConfig),
             (RR.RetrievalConfig, ODf.RetrievalRefConfig)]
    n = 0
    for ref_cls, our_cls in pairs:
        fa, fb = dataclasses.fields(ref_cls), dataclasses.fields(our_cls)
        assert [f.name for f in fa] == [f.name for f in fb], (ref_cls.__name__, [f.name for f in fa], [f.name for f in fb])
        a, b = ref_cls(), our_cls()
        for f in fa:
            assert getattr(a, f.name) == getattr(b, f.name), (ref_cls.__name__, f.name, getattr(a, f.name), getattr(b, f.name))
            n += 1
    # public methods: everything callable the reference classes expose exists on the mirror.  Left out on
    # purpose: the two data-loading helpers of the retrieval-reference generator (image files -> torchvision
    # transform -> encoder: upstream of the path)
    from multimodal_detection_consistency_b200 import metrics as OM
    allowed = {"RetrievalReferenceGenerator": {"build_reference_database", "load_reference_images"}}
    classes = [(mods["src.retrieval"], OR, ["MultiModalRetriever", "FaissIndexManager", "RetrievalIndex", "ConsistencyCalculator",
                                             "RetrievalResult"]),
               (mods["src.ref_bank"], OB, ["ReferenceBank", "ReferenceItem"]),
               (mods["src.detector"], ODt, ["AdversarialDetector"]),
               (mods["experiments.defenses.consistency_checker"], ODf, ["ConsistencyChecker"]),
               (mods["experiments.defenses.detector"], ODf, ["MultiModalDefenseDetector"]),
               (RR, ODf, ["RetrievalReferenceGenerator"]),
               (mods["src.utils.metrics"], OM, ["RetrievalEvaluator", "SimilarityCalculator", "RetrievalMetrics", "SimilarityMetrics"])]
    for ref_mod, our_mod, names in classes:
        for name in names:
            ra, rb = getattr(ref_mod, name), getattr(our_mod, name)
            pub = lambda c: {m for m in dir(c) if not m.startswith("_") and callable(getattr(c, m))}  # noqa: E731
            missing = pub(ra) - pub(rb) - allowed.get(name, set())
            assert not missing, (name, sorted(missing))
            n += len(pub(ra))
            # ... with the reference's parameter names, order and defaults (the mirror may append parameters)
            import inspect
            for meth in ["__init__"] + sorted(pub(ra) & pub(rb)):
                try:
                    pa = list(inspect.signature(getattr(ra, meth)).parameters.values())
                    pb = list(inspect.signature(getattr(rb, meth)).parameters.values())
                except (TypeError, ValueError):
                    continue
                assert [q.name for q in pb[:len(pa)]] == [q.name for q in pa], (name, meth, pa, pb)
                for x, y in zip(pa, pb):
                    if x.default is not inspect.Parameter.empty and x.default is not None:
                        assert x.default == y.default, (name, meth, x.name, x.default, y.default)
    for ref_mod, our_mod in ((mods["src.retrieval"], OR), (mods["src.ref_bank"], OB), (mods["src.detector"], ODt)):
        import types as _ty
        funcs = {k for k, v in vars(ref_mod).items() if isinstance(v, _ty.FunctionType) and v.__module__ == ref_mod.__name__
                 and not k.startswith("_")}
        assert funcs <= set(vars(our_mod)), (ref_mod.__name__, sorted(funcs - set(vars(our_mod))))
    ia = mods["src.ref_bank"].ReferenceItem(vector=np.arange(3.0), metadata={"a": 1}, timestamp=5.0)
    ib = OB.ReferenceItem(vector=np.arange(3.0), metadata={"a": 1}, timestamp=5.0)
    assert ia.to_dict() == ib.to_dict() and OB.ReferenceItem.from_dict(ia.to_dict()).to_dict() == ia.to_dict()
    for mod in (mods["src.ref_bank"], OB):                           # validation errors of ReferenceBankConfig (:37-44)
        for bad in (dict(clustering_method="spectral"), dict(update_strategy="mru")):
            try:
                mod.ReferenceBankConfig(**bad)
                raise AssertionError(f"{mod.__name__} accepted {bad}")
            except ValueError:
                pass
    return n


def main():
    seeds = [int(s) for s in sys.argv[1:]] or [31, 32]
    mods = MG.import_reference()
    with fake_native.installed():
        for seed in seeds:
            rng = np.random.default_rng(seed)
            np.random.seed(seed)
            done = {f.__name__[6:]: f(mods, rng) for f in (check_ref_bank, check_consistency_checker, check_hubness,
                                                           check_evaluator, check_retriever, check_detector,
                                                           check_retrieval_reference, check_defense_detector, check_configs)}
            print(f"seed {seed}: " + ", ".join(f"{k} {v}" for k, v in done.items()))
    print("mirrors live check ok")


if __name__ == "__main__":
    main()
