"""Generates tests/golden/*.npz by running the REFERENCE's own classes (imported from
/root/reference, read-only) on seeded inputs.  Run in the build container only — the GPU box has
no /root/reference; the committed .npz files are what travels.

    python tests/golden/make_golden.py

What runs is the reference's code, unmodified: ReferenceBank.query_similar, ConsistencyChecker.
make_decision, SimilarityCalculator, MultiModalRetriever._search_index (sklearn fallback branch),
ConsistencyCalculator, AdversarialDetector.detect_adversarial,
MultiModalDefenseDetector._compute_consistency_scores / _deduplicate_references,
HubnessAttack.compute_hubness, RetrievalEvaluator.compute_retrieval_metrics, and the `compute_hubness` pseudo-code block of
references/Adversarial_Hubness_Multi_Modal_Retrieval/README.md (exec'd from the markdown).
Modules the reference imports but does not ship or that are not installed here (src.models, faiss,
matplotlib, seaborn, plotly, nltk, and the syntactically broken experiments/defenses/
text_variants.py) are replaced by inert stubs; encoders are replaced by table look-ups into the
seeded synthetic embeddings, which is exactly the boundary north_star draws ("encoders remain
upstream producers of L2-normalised embeddings").
"""
from __future__ import annotations

import importlib
import importlib.machinery
import json
import re
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()

    def __getitem__(self, k):
        return _Any()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())


def _stub(name):
    m = types.ModuleType(name)
    m.__file__ = "<stub>"
    m.__path__ = []
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)

    def ga(n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Any()

    m.__getattr__ = ga
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, str(REF))
    for n in ["matplotlib", "matplotlib.pyplot", "seaborn", "plotly", "plotly.graph_objects", "plotly.express",
              "plotly.subplots", "nltk", "nltk.corpus", "nltk.tokenize", "nltk.tag", "faiss", "src.models",
              "src.models.clip_model"]:
        _stub(n)
    pkg = types.ModuleType("experiments.defenses")
    pkg.__path__ = [str(REF / "experiments" / "defenses")]
    sys.modules["experiments.defenses"] = pkg
    for n in ["text_variants", "retrieval_ref", "generative_ref"]:
        _stub("experiments.defenses." + n)
    mods = {}
    for name in ["src.retrieval", "src.detector", "src.attacks.hubness_attack", "src.ref_bank", "src.utils.metrics",
                 "experiments.defenses.consistency_checker", "experiments.defenses.detector"]:
        mods[name] = importlib.import_module(name)
    return mods


def unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


# ------------------------------------------------------------------------------------------
def gold_ref_bank(mods, out):
    RB = mods["src.ref_bank"]
    # (1) the snapshot the reference ships: cache/ref_bank/references.json (20 x 512, un-normalised)
    snap = json.loads((REF / "cache" / "ref_bank" / "references.json").read_text())
    snap = snap["references"] if isinstance(snap, dict) else snap
    vecs = np.array([r["vector"] for r in snap], dtype=np.float64)
    rng = np.random.default_rng(42)
    queries = vecs[rng.integers(0, len(vecs), 12)] + 0.35 * rng.standard_normal((12, vecs.shape[1])) * vecs.std()
    res = {}
    for tag, bank_vecs, qs, thr_cfg in [("snap", vecs, queries, 0.8)]:
        with tempfile.TemporaryDirectory() as td:
            cfg = RB.ReferenceBankConfig(max_size=1000, similarity_threshold=thr_cfg, persistence_enabled=False,
                                         save_path=td, auto_clustering=False, feature_dim=bank_vecs.shape[1])
            bank = RB.ReferenceBank(cfg)
            for i, v in enumerate(bank_vecs):
                bank.references.append(RB.ReferenceItem(vector=v.copy(), metadata={"i": i}, timestamp=0.0))
            idx_out = np.full((len(qs), 10), -1, np.int64)
            sim_out = np.full((len(qs), 10), -np.inf, np.float64)
            for qi, qv in enumerate(qs):
                for thr_arg, slot in [(None, 0)]:
                    r = bank.query_similar(qv, top_k=10, similarity_threshold=thr_arg)
                    for j, (item, s) in enumerate(r):
                        idx_out[qi, j] = item.metadata["i"]
                        sim_out[qi, j] = s
            res[tag] = (idx_out, sim_out)
    # (2) random fp32 bank, explicit thresholds incl. the falsy-zero quirk (src/ref_bank.py:191)
    bank_vecs = (rng.standard_normal((400, 128)) * 2.5).astype(np.float32)
    qs = (bank_vecs[rng.integers(0, 400, 16)] + 1.5 * rng.standard_normal((16, 128))).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        cfg = RB.ReferenceBankConfig(max_size=1000, similarity_threshold=0.6, persistence_enabled=False, save_path=td,
                                     auto_clustering=False, feature_dim=128)
        bank = RB.ReferenceBank(cfg)
        for i, v in enumerate(bank_vecs):
            bank.references.append(RB.ReferenceItem(vector=v.copy(), metadata={"i": i}, timestamp=0.0))
        thr_args = [None, 0.0, 0.3, 0.75]
        idx2 = np.full((len(thr_args), len(qs), 7), -1, np.int64)
        sim2 = np.full((len(thr_args), len(qs), 7), -np.inf, np.float64)
        for ti, ta in enumerate(thr_args):
            for qi, qv in enumerate(qs):
                for j, (item, s) in enumerate(bank.query_similar(qv, top_k=7, similarity_threshold=ta)):
                    idx2[ti, qi, j] = item.metadata["i"]
                    sim2[ti, qi, j] = s
        access = np.array([r.access_count for r in bank.references], np.int64)
    np.savez_compressed(out / "ref_bank.npz", snap_vectors=vecs, snap_queries=queries, snap_idx=res["snap"][0],
                        snap_sim=res["snap"][1], snap_cfg_threshold=0.8, bank=bank_vecs, queries=qs,
                        thr_args=np.array([np.nan, 0.0, 0.3, 0.75]), cfg_threshold=0.6, idx=idx2, sim=sim2,
                        access_count=access)


def gold_consistency_checker(mods, out):
    CC = mods["experiments.defenses.consistency_checker"]
    rng = np.random.default_rng(7)
    n = 400
    keys = ["original_similarity", "text_variant_consistency", "text_variant_std", "retrieval_consistency",
            "retrieval_std", "generative_consistency", "generative_std", "cross_modal_variance"]
    S = np.zeros((n, len(keys)))
    S[:, 0] = rng.uniform(-0.1, 0.9, n)
    S[:, 1] = rng.uniform(-0.1, 0.9, n)
    S[:, 2] = rng.uniform(0, 0.45, n)
    S[:, 3] = rng.uniform(0, 0.9, n) * (rng.uniform(size=n) > 0.15)
    S[:, 4] = rng.uniform(0, 0.45, n)
    S[:, 5] = rng.uniform(0, 0.9, n) * (rng.uniform(size=n) > 0.15)
    S[:, 6] = rng.uniform(0, 0.45, n)
    S[:, 7] = rng.uniform(0, 0.25, n) * (rng.uniform(size=n) > 0.3)
    res = {}
    for voting in ["simple", "weighted", "adaptive"]:
        for adaptive in [False, True]:
            o = np.zeros((n, 4))
            for i in range(n):
                chk = CC.ConsistencyChecker(threshold=0.5, adaptive_threshold=adaptive, voting_strategy=voting)
                r = chk.make_decision({k: float(S[i, j]) for j, k in enumerate(keys)})
                o[i] = [float(r["overall_score"]), float(r["threshold"]), float(r["confidence"]),
                        float(bool(r["is_adversarial"]))]
            res[f"{voting}_{int(adaptive)}"] = o
    # stateful threshold history (consistency_checker.py:234-239): one checker, 40 decisions
    chk = CC.ConsistencyChecker(threshold=0.5, adaptive_threshold=True, voting_strategy="weighted")
    hist = np.zeros((40, 4))
    for i in range(40):
        r = chk.make_decision({k: float(S[i, j]) for j, k in enumerate(keys)})
        hist[i] = [float(r["overall_score"]), float(r["threshold"]), float(r["confidence"]),
                   float(bool(r["is_adversarial"]))]
    np.savez_compressed(out / "consistency_checker.npz", keys=np.array(keys), scores=S, history=hist, **res)


def gold_similarity(mods, out):
    M = mods["src.utils.metrics"]
    R = mods["src.retrieval"]
    rng = np.random.default_rng(3)
    x = rng.standard_normal((30, 96)).astype(np.float32)
    y = rng.standard_normal((30, 96)).astype(np.float32)
    x[5] = 0.0
    pair = np.array([M.SimilarityCalculator.cosine_similarity(x[i], y[i]) for i in range(30)], np.float64)
    batch_np = M.SimilarityCalculator.batch_cosine_similarity(x[6:], y[6:])
    batch_t = M.SimilarityCalculator.batch_cosine_similarity(torch.from_numpy(x[6:]), torch.from_numpy(y[6:]))
    # MultiModalRetriever._search_index, sklearn fallback branch (src/retrieval.py:658-674)
    g = unit(rng, 2000, 128)
    q = unit(rng, 64, 128)
    ret = object.__new__(R.MultiModalRetriever)
    ret.config = R.RetrievalConfig(index_type="exact")
    ret.image_features = g
    ret.text_features = None
    idx = np.zeros((64, 10), np.int64)
    sc = np.zeros((64, 10), np.float32)
    for i in range(64):
        ii, ss = ret._search_index(None, q[i:i + 1], 10)
        idx[i], sc[i] = ii, ss
    ret.text_features = q
    mat_cos = ret.__class__.compute_similarity_matrix(ret)
    ret.config.similarity_metric = "dot_product"
    mat_dot = ret.__class__.compute_similarity_matrix(ret)
    ret.config.similarity_metric = "euclidean"
    mat_euc = ret.__class__.compute_similarity_matrix(ret)
    cc = R.ConsistencyCalculator()
    dist = cc.compute_similarity_distribution(sc[0])
    overlap = np.array([cc.compute_top_k_consistency(idx[i], idx[i + 1], 10) for i in range(63)])
    corr = np.array([cc.compute_consistency_score(sc[i], sc[i + 1]) for i in range(63)])
    np.savez_compressed(out / "similarity.npz", x=x, y=y, pair=pair, batch_np=batch_np, batch_t=batch_t, gallery=g,
                        queries=q, idx=idx, scores=sc, mat_cos=mat_cos[:8], mat_dot=mat_dot[:8], mat_euc=mat_euc[:8],
                        dist=np.array([dist[k] for k in ["mean", "std", "min", "max", "median"]]), overlap=overlap,
                        corr=corr)


class _TableClip:
    """Encoder stand-in: embeddings come from seeded tables keyed by id (src.models is not shipped)."""

    def __init__(self, text_table, image_table):
        self.t, self.i = text_table, image_table

    def encode_text(self, texts, normalize=True):
        return torch.stack([torch.from_numpy(self.t[s]) for s in texts])

    def encode_image(self, image, normalize=True):
        if isinstance(image, torch.Tensor) and image.ndim >= 1 and image.numel() == 1:
            return torch.from_numpy(self.i[int(image.reshape(-1)[0])])[None]
        return torch.from_numpy(self.i[int(image)])[None]

    def get_text_image_similarity(self, text, image):
        a = torch.from_numpy(self.t[text])
        b = torch.from_numpy(self.i[int(image)])
        return torch.nn.functional.cosine_similarity(a[None], b[None])[0]


def gold_detectors(mods, out, seed=11, nq=96):
    D = mods["src.detector"]
    ED = mods["experiments.defenses.detector"]
    rng = np.random.default_rng(seed)
    d, V, G, R = 64, 5, 3, 10
    gal = unit(rng, 300, d)
    gal[7] = gal[3]                      # exact duplicate rows
    gal[9] = unit(rng, 1, d)[0] * 0.02 + gal[4]
    gal[9] /= np.linalg.norm(gal[9])     # near duplicate (cos > 0.95)
    pick = rng.integers(0, 300, nq)
    sd = 1.0 / np.sqrt(d)
    attacked = rng.uniform(size=nq) < 0.35
    txt = gal[pick] + 0.5 * sd * rng.standard_normal((nq, d)).astype(np.float32)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    img = gal[pick] + 0.5 * sd * rng.standard_normal((nq, d)).astype(np.float32)
    img[attacked] = unit(rng, int(attacked.sum()), d) + 0.3 * gal[rng.integers(0, 300, int(attacked.sum()))]
    img /= np.linalg.norm(img, axis=1, keepdims=True)
    var = txt[:, None, :] + 0.25 * sd * rng.standard_normal((nq, V, d)).astype(np.float32)
    var /= np.linalg.norm(var, axis=2, keepdims=True)
    gen = img[:, None, :] * 0.6 + 0.8 * sd * rng.standard_normal((nq, G, d)).astype(np.float32) * 3
    gen /= np.linalg.norm(gen, axis=2, keepdims=True)
    g_cnt = rng.integers(0, G + 1, nq).astype(np.int32)
    txt, img, var, gen = (a.astype(np.float32) for a in (txt, img, var, gen))
    # retrieval candidates: top-10 of every variant row (variant-major), exact fp32
    sims = var.reshape(-1, d) @ gal.T
    cand = np.argsort(-sims, axis=1, kind="stable")[:, :10].reshape(nq, V * 10).astype(np.int64)
    cand[:, 3] = 3
    cand[:, 4] = 7       # duplicate row of 3
    cand[:, 5] = 4
    cand[:, 6] = 9       # near duplicate of 4

    # ---- src/detector.py AdversarialDetector.detect_adversarial --------------------------
    text_table = {}
    image_table = {}
    det_scores = np.zeros((4, nq, 6))
    agg_modes = ["weighted_mean", "mean", "max", "min"]
    for ai, mode in enumerate(agg_modes):
        for i in range(nq):
            text_table.clear()
            image_table.clear()
            text_table["orig"] = txt[i]
            for v in range(V):
                text_table[f"v{v}"] = var[i, v]
            image_table[0] = img[i]
            for g in range(G):
                image_table[1 + g] = gen[i, g]
            clip = _TableClip(text_table, image_table)
            det = D.AdversarialDetector(D.DetectorConfig(score_aggregation=mode, enable_cache=False))
            det._get_clip_model = lambda clip=clip: clip
            det._get_text_augmenter = lambda: types.SimpleNamespace(generate_variants=lambda t: [f"v{v}" for v in range(V)])
            ng = int(g_cnt[i])
            det._get_sd_generator = lambda ng=ng: types.SimpleNamespace(
                generate_reference_images=lambda text, num_images: {"images": list(range(1, 1 + ng))})
            det._image_to_features = lambda image, image_table=image_table: image_table[int(image)]
            r = det.detect_adversarial(0, "orig")
            assert "error" not in r, r
            ds = r["detection_scores"]
            det_scores[ai, i] = [ds["text_variants"], ds["sd_reference"], ds["consistency"], r["aggregated_score"],
                                 float(bool(r["is_adversarial"])),
                                 r["detection_details"]["text_variants"]["std_variant_similarity"]]

    # ---- experiments/defenses/detector.py _compute_consistency_scores --------------------
    keys = ["original_similarity", "text_variant_consistency", "text_variant_std", "retrieval_consistency",
            "retrieval_std", "generative_consistency", "generative_std", "cross_modal_variance"]
    cs = np.zeros((nq, len(keys)))
    n_ret = np.zeros(nq, np.int64)
    kept_idx = np.full((nq, R), -1, np.int64)
    for i in range(nq):
        text_table.clear()
        image_table.clear()
        text_table["orig"] = txt[i]
        for v in range(V):
            text_table[f"v{v}"] = var[i, v]
        image_table[100000] = img[i]
        for g in range(G):
            image_table[200000 + g] = gen[i, g]
        for c in cand[i]:
            image_table[int(c)] = gal[int(c)]
        clip = _TableClip(text_table, image_table)
        det = object.__new__(ED.MultiModalDefenseDetector)
        det.clip_model = clip
        det.config = ED.DetectionConfig()
        refs = [torch.tensor(int(c)) for c in cand[i]]
        uniq = det._deduplicate_references(refs)[: det.config.retrieval_top_k]
        n_ret[i] = len(uniq)
        kept_idx[i, : len(uniq)] = [int(u) for u in uniq]
        gens = [torch.tensor(200000 + g) for g in range(int(g_cnt[i]))]
        sc = det._compute_consistency_scores(torch.tensor(100000), "orig", ["orig"] + [f"v{v}" for v in range(V)], uniq, gens)
        cs[i] = [float(sc[k]) for k in keys]
    np.savez_compressed(out / "detectors.npz", gallery=gal, img=img, txt=txt, var=var, gen=gen, g_cnt=g_cnt, cand=cand,
                        agg_modes=np.array(agg_modes), det_scores=det_scores, cs_keys=np.array(keys), cs=cs,
                        n_ret=n_ret, kept_idx=kept_idx)


def gold_hubness(mods, out):
    H = mods["src.attacks.hubness_attack"]
    res = {}
    # the shapes/seed of benchmarks/hubness_attack_benchmark.py:317-329
    for (ni, nq, d) in [(10, 5, 128), (50, 20, 256), (100, 50, 512)]:
        torch.manual_seed(42)
        im = torch.nn.functional.normalize(torch.randn(ni, d), p=2, dim=1)
        tx = torch.nn.functional.normalize(torch.randn(nq, d), p=2, dim=1)
        # make image 0 a hub for part of the queries
        im[0] = torch.nn.functional.normalize(tx[: max(1, nq // 2)].mean(0), dim=0)
        score = H.HubnessAttack.compute_hubness(None, im, tx, 10)
        res[f"bench_{ni}_{nq}_{d}_img"] = im.numpy()
        res[f"bench_{ni}_{nq}_{d}_txt"] = tx.numpy()
        res[f"bench_{ni}_{nq}_{d}_score"] = np.float64(score)
    # the k-occurrence pseudo-code, exec'd from the markdown
    md = (REF / "references" / "Adversarial_Hubness_Multi_Modal_Retrieval" / "README.md").read_text()
    block = [b for b in re.findall(r"```python\n(.*?)```", md, flags=re.S) if "def compute_hubness" in b][0]
    ns = {"np": np}
    from sklearn.metrics.pairwise import cosine_similarity
    ns["cosine_similarity"] = cosine_similarity
    exec(block, ns)
    rng = np.random.default_rng(42)
    cent = unit(rng, 12, 64)
    f = cent[rng.integers(0, 12, 600)] + 0.08 * rng.standard_normal((600, 64)).astype(np.float32)
    f = f.astype(np.float32)
    hub = ns["compute_hubness"](f, k=10)
    res["spec_features"] = f
    res["spec_hubness"] = np.asarray(hub, np.float64)
    np.savez_compressed(out / "hubness.npz", **res)


def gold_retrieval_metrics(mods, out):
    """RetrievalEvaluator.compute_retrieval_metrics (src/utils/metrics.py:386-574) on similarity
    matrices whose values are distinct (the reference's argsort has no defined tie order)."""
    M = mods["src.utils.metrics"]
    rng = np.random.default_rng(11)
    res = {}
    for tag, nq, nc, p_rel in [("small", 40, 50, 0.08), ("wide", 60, 400, 0.01)]:
        sims = rng.permutation(nq * nc).reshape(nq, nc).astype(np.float64) / (nq * nc)
        rel = (rng.uniform(size=(nq, nc)) < p_rel).astype(np.int64)
        rel[0] = 0                                        # a query without relevant items
        ks = [1, 5, 10, 20, 50]
        m = M.RetrievalEvaluator.compute_retrieval_metrics(sims, rel, ks)
        per_q = np.zeros((nq, 2), np.float64)
        order = np.argsort(-sims, axis=1)
        for i in range(nq):
            sr = rel[i][order[i]]
            per_q[i, 0] = M.RetrievalEvaluator._compute_reciprocal_rank(sr)
            per_q[i, 1] = M.RetrievalEvaluator._compute_average_precision(sr)
        res.update({f"{tag}_sims": sims.astype(np.float32), f"{tag}_rel": rel, f"{tag}_ks": np.array(ks),
                    f"{tag}_recall": np.array([m.recall_at_k[k] for k in ks]),
                    f"{tag}_precision": np.array([m.precision_at_k[k] for k in ks]),
                    f"{tag}_ndcg": np.array([m.ndcg_at_k[k] for k in ks]),
                    f"{tag}_map": np.float64(m.map_score), f"{tag}_mrr": np.float64(m.mrr), f"{tag}_per_query": per_q})
    np.savez_compressed(out / "retrieval_metrics.npz", **res)


def main():
    # `--out DIR` writes somewhere else (tests/test_golden_reproducible.py regenerates into a temp
    # directory and compares with the committed fixtures)
    argv = sys.argv[1:]
    out = OUT
    if "--out" in argv:
        i = argv.index("--out")
        out = Path(argv[i + 1])
        out.mkdir(parents=True, exist_ok=True)
        del argv[i:i + 2]
    mods = import_reference()
    if argv and argv[0] == "retrieval_metrics":
        gold_retrieval_metrics(mods, out)
        return
    gold_ref_bank(mods, out)
    gold_consistency_checker(mods, out)
    gold_similarity(mods, out)
    gold_detectors(mods, out)
    gold_hubness(mods, out)
    gold_retrieval_metrics(mods, out)
    for f in sorted(out.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
