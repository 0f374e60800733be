"""The reference-shaped host classes on the GPU, checked against the oracle and against the outputs
of the reference's own classes (tests/golden/*.npz)."""
import types
from pathlib import Path

import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


class TableClip:
    """Encoder stand-in: text / image ids -> rows of seeded embedding tables."""

    def __init__(self, text_table, image_table):
        self.t, self.i = text_table, image_table

    def encode_text(self, texts, normalize=True):
        return np.stack([self.t[s] for s in texts])

    def encode_image(self, images, normalize=True):
        if not isinstance(images, (list, tuple)):
            images = [images]
        return np.stack([self.i[int(s)] for s in images])


def _unit(rng, n, d):
    return O.l2_normalize(rng.standard_normal((n, d), dtype=np.float32))


def test_retriever_drop_in(tmp_path):
    from multimodal_detection_consistency_b200 import MultiModalRetriever, RetrievalConfig
    rng = np.random.default_rng(0)
    n, d = 3000, 128
    gal = _unit(rng, n, d)
    texts = {f"caption {i}": O.l2_normalize(gal[i % n][None] + 0.05 * rng.standard_normal((1, d), dtype=np.float32))[0]
             for i in range(40)}
    clip = TableClip(texts, {})
    r = MultiModalRetriever(RetrievalConfig(top_k=10), clip_model=clip)
    assert r.retrieve_images_by_text("caption 0") == ([], [])        # no index yet: reference error convention
    paths = [f"img_{i}.jpg" for i in range(n)]
    r.build_image_index_from_features(gal, paths)
    q = np.stack(list(texts.values()))
    ref_s, ref_i = O.search(q, gal, 10)
    for j, t in enumerate(texts):
        p, s = r.retrieve_images_by_text(t)
        assert p == [paths[i] for i in ref_i[j]]
        assert np.abs(np.array(s) - ref_s[j]).max() <= 1e-5
        assert r.retrieve(t, k=5)[0] == p[:5] and r.search(t, top_k=5)[0] == p[:5]   # caller-side spellings
    batch = r.batch_retrieve_images_by_texts(list(texts), top_k=7)
    assert [b[0] for b in batch] == [[paths[i] for i in ref_i[j][:7]] for j in range(len(texts))]
    assert r.get_stats()["image_count"] == n and r.get_stats()["retrieval_cache_size"] > 0
    # _search_index keeps the (indices, scores) order and FAISS -1 padding
    small = MultiModalRetriever(RetrievalConfig(), clip_model=clip)
    small.build_image_index_from_features(gal[:4], paths[:4])
    idx, sc = small._search_index(small.image_index, q[:1], 6)
    assert idx.tolist()[4:] == [-1, -1] and np.isneginf(sc[4:]).all()
    # persistence round trip
    r.save_image_index(str(tmp_path / "idx.pkl"))
    r2 = MultiModalRetriever(RetrievalConfig(), clip_model=clip)
    r2.load_image_index(str(tmp_path / "idx.pkl"))
    assert r2.retrieve_images_by_text("caption 3") == r.retrieve_images_by_text("caption 3")
    # similarity matrix metrics
    r.text_features = q
    for metric, tol in [("dot_product", 2e-3), ("cosine", 2e-3), ("euclidean", 2e-3)]:
        r.config.similarity_metric = metric
        assert np.abs(r.compute_similarity_matrix() - O.similarity_matrix(q, gal, metric)).max() <= tol


def test_faiss_compat_surface(tmp_path):
    from multimodal_detection_consistency_b200 import faiss_compat as faiss
    rng = np.random.default_rng(1)
    g, q = _unit(rng, 1500, 64), _unit(rng, 20, 64)
    index = faiss.IndexFlatIP(64)
    assert index.is_trained and index.ntotal == 0
    index.add(g[:1000])
    index.add(g[1000:])
    assert index.ntotal == 1500
    D, I = index.search(q, 10)
    ref_s, ref_i = O.search(q, g, 10)
    assert np.array_equal(I, ref_i) and np.abs(D - ref_s).max() <= 1e-5
    faiss.write_index(index, str(tmp_path / "x.faiss"))
    D2, I2 = faiss.read_index(str(tmp_path / "x.faiss")).search(q, 10)
    assert np.array_equal(I2, I)
    assert faiss.index_cpu_to_gpu(faiss.StandardGpuResources(), 0, index) is index and faiss.get_num_gpus() >= 1


def test_reference_bank_matches_reference_outputs(tmp_path):
    """Golden: the reference's ReferenceBank.query_similar on its shipped snapshot and on a random bank."""
    from multimodal_detection_consistency_b200 import ReferenceBank, ReferenceBankConfig, ReferenceItem
    z = np.load(GOLD / "ref_bank.npz")
    for vecs, queries, idx_w, sim_w, cfg_thr, thr_args in [
            (z["snap_vectors"], z["snap_queries"], z["snap_idx"][None], z["snap_sim"][None], 0.8, [None]),
            (z["bank"], z["queries"], z["idx"], z["sim"], 0.6, [None, 0.0, 0.3, 0.75])]:
        bank = ReferenceBank(ReferenceBankConfig(max_size=1000, similarity_threshold=cfg_thr,
                                                 persistence_enabled=False, save_path=str(tmp_path / "b"),
                                                 auto_clustering=False, feature_dim=vecs.shape[1]))
        items = [ReferenceItem(vector=v.copy(), metadata={"i": i}, timestamp=0.0) for i, v in enumerate(vecs)]
        bank.references.extend(items)
        bank._dev_append(items)
        k = idx_w.shape[2]
        for ti, ta in enumerate(thr_args):
            for qi, qv in enumerate(queries):
                got = bank.query_similar(qv, top_k=k, similarity_threshold=ta)
                n = int((idx_w[ti, qi] >= 0).sum())
                assert [it.metadata["i"] for it, _ in got] == idx_w[ti, qi, :n].tolist()
                assert np.abs(np.array([s for _, s in got]) - sim_w[ti, qi, :n]).max(initial=0) <= 1e-5
            batch = bank.query_similar_batch(queries, top_k=k, similarity_threshold=ta)
            assert [[it.metadata["i"] for it, _ in row] for row in batch] == \
                   [idx_w[ti, qi][idx_w[ti, qi] >= 0].tolist() for qi in range(len(queries))]
    assert sum(it.access_count for it in items) > 0


def test_reference_bank_insert_dedup_eviction_persistence(tmp_path):
    from multimodal_detection_consistency_b200 import ReferenceBank, ReferenceBankConfig
    rng = np.random.default_rng(2)
    cfg = ReferenceBankConfig(max_size=50, similarity_threshold=0.9, persistence_enabled=True,
                              save_path=str(tmp_path / "bank"), auto_clustering=False, feature_dim=32)
    bank = ReferenceBank(cfg)
    vecs = rng.standard_normal((80, 32)).astype(np.float32)
    assert bank.add_reference(vecs[0], {"i": 0})
    assert not bank.add_reference(vecs[0] * 1.5 + 1e-3, {"i": "dup"})      # cosine > 0.9 -> rejected
    for i in range(1, 80):
        assert bank.add_reference(vecs[i], {"i": i})
    assert len(bank.references) == 50 and bank.stats["total_removed"] == 30   # fifo eviction
    assert [r.metadata["i"] for r in bank.references] == list(range(30, 80))
    live = np.stack([r.vector for r in bank.references])
    for qv in vecs[40:45]:
        got = bank.query_similar(qv + 0.2 * rng.standard_normal(32).astype(np.float32), top_k=5,
                                 similarity_threshold=0.5)
        idx, sim = O.ref_bank_query(live, qv, top_k=5, similarity_threshold=0.5)
        assert len(got) >= 1 and got[0][0].metadata["i"] == 30 + int(idx[0])
    sims = bank._compute_similarities(vecs[33])
    assert np.abs(sims - O.ref_bank_similarities(live, vecs[33])).max() <= 2e-3
    # write-through journal: 110 operations so far, nothing folded yet; a reload replays it
    root = tmp_path / "bank"
    assert (root / "references.journal.jsonl").exists() and not (root / "references.json").exists()
    bank2 = ReferenceBank(cfg)
    assert len(bank2.references) == 50 and bank2.stats["total_added"] == 80 and bank2.stats["total_removed"] == 30
    assert [r.metadata["i"] for r in bank2.references] == list(range(30, 80))
    # fold into the reference's 4-file JSON layout (src/ref_bank.py:505-576) and reload from that
    bank.flush()
    import json
    assert not (root / "references.journal.jsonl").exists()
    assert len(json.loads((root / "references.json").read_text())) == 50
    assert sorted(p.name for p in root.iterdir()) == ["clusters.json", "config.json", "references.json", "stats.json"]
    bank2 = ReferenceBank(cfg)
    assert len(bank2.references) == 50
    a, b = bank.query_similar(vecs[60], 3, 0.5), bank2.query_similar(vecs[60], 3, 0.5)
    assert [x[0].metadata for x in a] == [x[0].metadata for x in b]
    st = bank.get_statistics()
    assert st["current_size"] == 50 and st["update_strategy"] == "fifo"
    # similarity eviction removes one member of the closest pair
    cfg3 = ReferenceBankConfig(max_size=4, similarity_threshold=0.99999, update_strategy="similarity",
                               persistence_enabled=False, save_path=str(tmp_path / "b3"), auto_clustering=False)
    b3 = ReferenceBank(cfg3)
    base = rng.standard_normal((4, 32)).astype(np.float32)
    base[2] = base[1] + 0.05 * rng.standard_normal(32).astype(np.float32)
    for i in range(4):
        assert b3.add_reference(base[i], {"i": i})
    assert b3.add_reference(rng.standard_normal(32).astype(np.float32), {"i": 4})
    kept = [r.metadata["i"] for r in b3.references]
    assert len(kept) == 4 and (1 in kept) != (2 in kept)


def test_reference_bank_journal_never_drifts_from_memory(tmp_path):
    """clear(), import_references() and perform_clustering() change the bank outside the add / evict
    journal; the disk state must still reload to exactly the memory state, and a crash between a fold's
    replace of references.json and the unlink of the journal must not replay folded operations."""
    import json
    import shutil
    from multimodal_detection_consistency_b200 import ReferenceBank, ReferenceBankConfig
    rng = np.random.default_rng(9)
    root = tmp_path / "bank"
    cfg = ReferenceBankConfig(max_size=12, similarity_threshold=0.95, persistence_enabled=True, save_path=str(root),
                              auto_clustering=False, feature_dim=16, num_clusters=3)
    vecs = rng.standard_normal((60, 16)).astype(np.float32)
    ids = lambda b: [r.metadata["i"] for r in b.references]   # noqa: E731
    bank = ReferenceBank(cfg)
    for i in range(5):
        assert bank.add_reference(vecs[i], {"i": i})
    # clear -> add: the five journalled adds must not come back
    bank.clear()
    assert bank.add_reference(vecs[5], {"i": 5})
    assert ids(ReferenceBank(cfg)) == [5]
    # import -> add -> evict: the imported rows are on disk before later `remove index` lines refer to them
    donor = ReferenceBank(ReferenceBankConfig(max_size=50, persistence_enabled=False, save_path=str(tmp_path / "d"),
                                              auto_clustering=False, feature_dim=16))
    for i in range(10, 18):
        assert donor.add_reference(vecs[i], {"i": i})
    assert donor.export_references(str(tmp_path / "donor.json"))
    assert bank.import_references(str(tmp_path / "donor.json"))
    assert ids(bank) == [5] + list(range(10, 18))
    for i in range(20, 28):                                   # 9 + 8 > 12: fifo evictions of 5, 10, 11, ...
        assert bank.add_reference(vecs[i], {"i": i})
    assert len(bank.references) == 12
    again = ReferenceBank(cfg)
    assert ids(again) == ids(bank)
    assert again.stats["total_added"] == bank.stats["total_added"]
    assert again.stats["total_removed"] == bank.stats["total_removed"]
    # public perform_clustering folds the new cluster ids
    assert bank.perform_clustering(force=True)
    again = ReferenceBank(cfg)
    assert [r.cluster_id for r in again.references] == [r.cluster_id for r in bank.references]
    assert {int(k): v for k, v in again.clusters.items()} == {int(k): v for k, v in bank.clusters.items()}
    # crash between the fold's os.replace(references.json) and the unlink of the journal
    assert bank.add_reference(vecs[30], {"i": 30}) and bank.add_reference(vecs[31], {"i": 31})
    journal = root / "references.journal.jsonl"
    assert journal.exists()
    shutil.copy(journal, tmp_path / "journal.keep")
    bank.flush()
    assert not journal.exists()
    shutil.copy(tmp_path / "journal.keep", journal)           # the unlink "did not happen"
    again = ReferenceBank(cfg)
    assert ids(again) == ids(bank) and len(again.references) == 12
    first = json.loads((root / "references.json").read_text())[0]
    assert first["_journal_seq"] >= 2 and not list(root.glob("*.tmp"))
    # and the operations after such a recovery keep counting from the folded number
    assert again.add_reference(vecs[32], {"i": 32})
    assert ids(ReferenceBank(cfg)) == ids(again)


def test_reference_bank_kmeans_assignment_matches_sklearn(tmp_path):
    """SURVEY.md §8f rank 3 / src/ref_bank.py:296-300: KMeans with the assignment step as an inner-product top-1
    search (kernel a) gives scikit-learn's labels and centres."""
    from sklearn.cluster import KMeans
    from multimodal_detection_consistency_b200 import ReferenceBank, ReferenceBankConfig
    rng = np.random.default_rng(21)
    for n, d, c, dtype in ((300, 48, 7, np.float64), (900, 96, 20, np.float32), (150, 24, 40, np.float64)):
        cent = rng.standard_normal((c, d)) * 2.0
        vecs = (cent[rng.integers(0, c, n)] + 0.6 * rng.standard_normal((n, d))).astype(dtype)
        bank = ReferenceBank(ReferenceBankConfig(max_size=n, similarity_threshold=0.9999, persistence_enabled=False,
                                                 save_path=str(tmp_path / "km"), auto_clustering=False, feature_dim=d,
                                                 num_clusters=c))
        for i in range(n):
            assert bank.add_reference(vecs[i], {"i": i})
        fit = bank._kmeans_device_assign(np.array([r.vector for r in bank.references]), c)
        assert fit is not None, "the device-assignment path must serve a non-degenerate input"
        assert bank.perform_clustering()
        km = KMeans(n_clusters=c, random_state=42, n_init=10)
        want = km.fit_predict(np.array([r.vector for r in bank.references]))
        got = np.array([r.cluster_id for r in bank.references])
        assert np.array_equal(got, want), (n, d, c, int((got != want).sum()))
        assert np.allclose(bank.get_cluster_centers(), km.cluster_centers_, rtol=1e-5, atol=1e-6)


def test_adversarial_detector_matches_reference_outputs():
    """Golden: the reference's AdversarialDetector.detect_adversarial on the same table encoders."""
    from multimodal_detection_consistency_b200 import AdversarialDetector, DetectorConfig
    z = np.load(GOLD / "detectors.npz")
    img, txt, var, gen, g_cnt = z["img"], z["txt"], z["var"], z["gen"], z["g_cnt"]
    nq, V = var.shape[0], var.shape[1]
    for mode_i, mode in enumerate([str(m) for m in z["agg_modes"]]):
        want = z["det_scores"][mode_i]
        results = []
        dets = []
        for i in range(nq):
            tt = {"orig": txt[i], **{f"v{v}": var[i, v] for v in range(V)}}
            it = {0: img[i], **{1 + g: gen[i, g] for g in range(gen.shape[1])}}
            ng = int(g_cnt[i])
            det = AdversarialDetector(
                DetectorConfig(score_aggregation=mode, enable_cache=False), clip_model=TableClip(tt, it),
                text_augmenter=types.SimpleNamespace(generate_variants=lambda t: [f"v{v}" for v in range(V)]),
                sd_generator=types.SimpleNamespace(
                    generate_reference_images=lambda text, num_images, ng=ng: {"images": list(range(1, 1 + ng))}))
            r = det.detect(0, "orig")
            assert "error" not in r, r
            results.append(r)
            dets.append(det)
        got = np.array([[r["detection_scores"]["text_variants"], r["detection_scores"]["sd_reference"],
                         r["detection_scores"]["consistency"], r["aggregated_score"]] for r in results])
        assert np.abs(got - want[:, :4]).max() <= 1e-5
        margin = np.abs(want[:, 3] - 0.5) > 1e-5
        assert np.array_equal(np.array([r["is_adversarial"] for r in results])[margin], want[margin, 4].astype(bool))
        d0 = results[0]["detection_details"]
        assert set(d0["text_variants"]) >= {"original_similarity", "variant_similarities", "mean_variant_similarity",
                                            "std_variant_similarity", "consistency_score", "variability_score",
                                            "num_variants"}
        assert abs(d0["text_variants"]["std_variant_similarity"] - want[0, 5]) <= 1e-5
    # batched embedding entry: one launch for all samples
    det = AdversarialDetector(DetectorConfig())
    out = det.detect_embeddings(img, txt, var, gen, g_cnt)
    assert np.abs(out["aggregated_score"] - z["det_scores"][0][:, 3]).max() <= 1e-5
    # never-raise convention
    bad = AdversarialDetector(DetectorConfig()).detect_adversarial(0, "x")
    assert bad["is_adversarial"] is False and bad["aggregated_score"] == 0.0 and "error" in bad


def test_consistency_checker_matches_reference_outputs():
    from multimodal_detection_consistency_b200 import ConsistencyChecker
    z = np.load(GOLD / "consistency_checker.npz")
    keys = [str(k) for k in z["keys"]]
    S = z["scores"]
    for voting in ["simple", "weighted", "adaptive"]:
        for adaptive in [0, 1]:
            want = z[f"{voting}_{adaptive}"]
            for i in range(0, 120):
                chk = ConsistencyChecker(threshold=0.5, adaptive_threshold=bool(adaptive), voting_strategy=voting)
                r = chk.make_decision({k: float(S[i, j]) for j, k in enumerate(keys)})
                assert abs(r["overall_score"] - want[i, 0]) <= 1e-5
                assert abs(r["threshold"] - want[i, 1]) <= 1e-6
                assert abs(r["confidence"] - want[i, 2]) <= 1e-5
                if abs(want[i, 0] - want[i, 1]) > 1e-5:
                    assert r["is_adversarial"] == bool(want[i, 3])
    chk = ConsistencyChecker(threshold=0.5, adaptive_threshold=True, voting_strategy="weighted")
    for i in range(z["history"].shape[0]):   # stateful threshold history, 40 sequential decisions
        r = chk.make_decision({k: float(S[i, j]) for j, k in enumerate(keys)})
        w = z["history"][i]
        assert abs(r["threshold"] - w[1]) <= 1e-6 and abs(r["confidence"] - w[2]) <= 1e-5
        if abs(w[0] - w[1]) > 1e-5:
            assert r["is_adversarial"] == bool(w[3])
    assert chk.get_statistics()["total_detections"] == 40


def test_defense_detector_batched_matches_oracle():
    from multimodal_detection_consistency_b200 import DetectionConfig, MultiModalDefenseDetector, RetrievalReferenceIndex
    d, nq, V = 256, 120, 5
    g = O.synth_gallery(2500, d, seed=5, clusters=32)
    img, txt, var = O.synth_queries(g, nq, V, seed=6)
    rng = np.random.default_rng(7)
    gen = O.l2_normalize((img[:, None, :] + 0.6 * rng.standard_normal((nq, 3, d), dtype=np.float32)).reshape(-1, d)).reshape(nq, 3, d)
    idx = RetrievalReferenceIndex(g)
    det = MultiModalDefenseDetector(clip_model=None, config=DetectionConfig(), retrieval_index=idx)
    res, scores = det.detect_embeddings(img, txt, var, gen, return_scores=True)
    rows = np.concatenate([txt[:, None, :], var], 1).reshape(-1, d)
    rs, ri = O.search(rows, g, 20, threshold=0.3)
    cand = ri.reshape(nq, V + 1, 20)[:, :, :5].reshape(nq, -1)
    ref, rflags, _ = O.consistency_emb(img, txt, var, ret_rows=g, ret_idx=cand, gen=gen)
    assert np.abs(scores - ref).max() <= 1e-4
    # one checker decides the batch in order, so the reference's threshold history applies (:234-239)
    hist = []
    for i in range(nq):
        s = ref[i]
        overall, thr, conf, adv = O.consistency_from_scores(
            s[O.S_ORIGINAL], s[O.S_TV_MEAN], s[O.S_TV_STD], s[O.S_RET_MEAN], s[O.S_RET_STD], s[O.S_GEN_MEAN],
            s[O.S_GEN_STD], s[O.S_CROSS_MODAL_VAR], threshold_history=hist)
        hist.append(thr)
        assert abs(res[i]["consistency_score"] - overall) <= 1e-4 and abs(res[i]["confidence"] - conf) <= 1e-3
        if abs(overall - thr) > 1e-4:
            assert res[i]["is_adversarial"] == adv
    assert det.consistency_checker.threshold_history == pytest.approx(hist, abs=1e-5)
    del rflags
    assert set(res[0]) == {"is_adversarial", "confidence", "consistency_score"}


def test_hubness_api_matches_reference_outputs():
    import torch
    from multimodal_detection_consistency_b200 import compute_hubness, hubness_scores, k_occurrence
    z = np.load(GOLD / "hubness.npz")
    for (ni, nq, d) in [(10, 5, 128), (50, 20, 256), (100, 50, 512)]:
        im, tx = z[f"bench_{ni}_{nq}_{d}_img"], z[f"bench_{ni}_{nq}_{d}_txt"]
        assert compute_hubness(im, tx, 10) == float(z[f"bench_{ni}_{nq}_{d}_score"])
        assert compute_hubness(torch.from_numpy(im).cuda(), torch.from_numpy(tx).cuda()) == float(z[f"bench_{ni}_{nq}_{d}_score"])
        c = k_occurrence(tx, im, 3)
        _, ri = O.search(tx, im, 3, metric="cosine")
        assert np.array_equal(c, O.k_occurrence(ri, ni)) and 0 <= c.min() and c.max() <= nq
    f = z["spec_features"]
    counts, hub = hubness_scores(f, 10)
    assert counts.sum() == f.shape[0] * 10
    assert np.abs(hub - z["spec_hubness"]).sum() <= 0.01 * z["spec_hubness"].sum() + 1e-12


def test_pipeline_worker_threads_are_micro_batched(tvc_ctx):
    """The reference's pipeline calls retrieve_images_by_text / detect_adversarial one sample at a
    time from 4 worker threads (src/pipeline.py:288,450-476,519-526,555-560); concurrent calls must
    return exactly the sequential results while sharing encoder calls and kernel launches."""
    from concurrent.futures import ThreadPoolExecutor
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(0)
    d, n = 64, 500
    gallery = O.l2_normalize(rng.standard_normal((n, d)).astype(np.float32))

    class Clip:                                  # table look-up encoder, counts its calls
        def __init__(self):
            self.calls = 0

        def _emb(self, keys):
            self.calls += 1
            out = np.stack([np.random.default_rng(abs(hash(str(k))) % (2 ** 31)).standard_normal(d) for k in keys])
            return O.l2_normalize(out.astype(np.float32))

        def encode_text(self, texts, normalize=True):
            return self._emb(texts)

        def encode_image(self, images, normalize=True):
            return self._emb(images)

    class Aug:
        def generate_variants(self, text):
            return [f"{text} v{i}" for i in range(5)]

    texts = [f"a photo of thing {i}" for i in range(48)]
    images = [f"img{i}" for i in range(48)]
    seq_clip, par_clip = Clip(), Clip()
    results = {}
    for tag, clip, workers in (("seq", seq_clip, 1), ("par", par_clip, 8)):
        r = tvc.MultiModalRetriever(tvc.RetrievalConfig(enable_cache=False), clip_model=clip)
        r.build_image_index_from_features(gallery, [f"p{i}.jpg" for i in range(n)])
        det = tvc.AdversarialDetector(tvc.DetectorConfig(detection_methods=["text_variants", "consistency"],
                                                         enable_cache=False), clip_model=clip, text_augmenter=Aug())
        r._t2i_batcher.max_delay_s = det._batcher.max_delay_s = 0.02
        with ThreadPoolExecutor(max_workers=workers) as ex:
            ret = list(ex.map(lambda t: r.retrieve_images_by_text(t, top_k=5), texts))
            dets = list(ex.map(lambda it: det.detect_adversarial(it[0], it[1]), zip(images, texts)))
        results[tag] = (ret, dets, r._t2i_batcher.stats(), det._batcher.stats())
    (ret_s, det_s, _, _), (ret_p, det_p, st_r, st_d) = results["seq"], results["par"]
    assert ret_p == ret_s and all(len(p) == 5 for p, _ in ret_p)
    for a, b in zip(det_p, det_s):
        assert a["is_adversarial"] == b["is_adversarial"]
        assert abs(a["aggregated_score"] - b["aggregated_score"]) < 1e-6
        assert a["detection_details"]["text_variants"]["num_variants"] == 5
    assert st_r["rounds"] < len(texts) and st_d["rounds"] < len(texts)      # calls were coalesced
    assert par_clip.calls < seq_clip.calls


def test_c_probe_runs_the_hot_path_from_c(tmp_path):
    """tests/c/abi_probe.c on the GPU box: a plain-C process (dlopen only, no Python, no torch) builds a gallery,
    searches it with host buffers, checks the result against a scalar loop and histograms it."""
    import shutil
    import subprocess
    from multimodal_detection_consistency_b200 import _native as N
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "abi_probe"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(root / "include"),
                    str(root / "tests" / "c" / "abi_probe.c"), "-o", str(exe), "-ldl", "-lm"], check=True)
    out = subprocess.run([str(exe), str(N.LIB_PATH)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "search + k-occurrence from C" in out.stdout
