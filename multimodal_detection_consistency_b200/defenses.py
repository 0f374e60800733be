"""Drop-ins for the experiments-side TVC stack: `ConsistencyChecker`
(experiments/defenses/consistency_checker.py:31), `DetectionConfig` / `MultiModalDefenseDetector`
(experiments/defenses/detector.py:19,46) and the retrieval-reference generator front
(experiments/defenses/retrieval_ref.py:34).

The per-sample arithmetic of `_compute_consistency_scores` (:228-293: <= 19 encoder calls + scalar
cosines + np.mean / np.std / np.var) and of `make_decision` (voting, stateless adaptive threshold,
confidence) is kernel (b); retrieval of references (:184-204 -> retrieval_ref.py:246-290) is kernel
(a).  `detect_batch` / `detect_embeddings` score whole batches per launch.  The one stateful piece —
the threshold history blend (consistency_checker.py:234-239) — stays on the host, applied to the
kernel's stateless threshold, as the reference applies it after its own stateless adjustments.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _native as N
from ._native import Gallery

logger = logging.getLogger(__name__)

_VOTING = {"simple": 0, "weighted": 1, "adaptive": 2}
_CC_KEYS = ("original_similarity", "text_variant_consistency", "retrieval_consistency", "generative_consistency")


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


class ConsistencyChecker:
    """Same constructor, attributes and result dict as the reference (consistency_checker.py:31-117)."""

    def __init__(self, threshold: float = 0.5, adaptive_threshold: bool = True, voting_strategy: str = "weighted",
                 weights: Optional[Dict[str, float]] = None):
        if voting_strategy not in _VOTING:
            raise ValueError(f"unknown voting strategy: {voting_strategy}")
        self.base_threshold = threshold
        self.adaptive_threshold = adaptive_threshold
        self.voting_strategy = voting_strategy
        self.weights = weights or {k: 0.25 for k in _CC_KEYS}
        self.detection_history: List[Dict[str, Any]] = []
        self.threshold_history: List[float] = []

    def params(self, v: int = 5, r: int = 10, g: int = 3, **extra) -> N.DetectorParams:
        return N.default_params(n_variants=v, n_retrieval=r, n_generative=g, voting=_VOTING[self.voting_strategy],
                                cc_weights=[float(self.weights.get(k, 0.0)) for k in _CC_KEYS],
                                cc_base_threshold=float(self.base_threshold), cc_adaptive=int(self.adaptive_threshold),
                                **extra)

    # -- stateful host part -------------------------------------------------------------------
    def _blend(self, stateless_thr: float) -> float:
        """consistency_checker.py:234-242 applied to the kernel's (already adjusted) threshold."""
        thr = float(stateless_thr)
        if self.adaptive_threshold and len(self.threshold_history) > 10:
            thr = 0.7 * thr + 0.3 * float(np.mean(self.threshold_history[-10:]))
            thr = float(np.clip(thr, 0.1, 0.9))
        return thr

    def decide_from_kernel(self, score_rows: np.ndarray) -> List[Dict[str, Any]]:
        """Turn kernel (b) rows [Q, 24] into the reference's decision dicts, in order, maintaining the
        threshold / detection histories exactly as sequential `make_decision` calls would."""
        ix = N.SCORE_INDEX
        out = []
        for s in np.asarray(score_rows):
            overall = float(s[ix["overall_score"]])
            thr = self._blend(float(s[ix["threshold"]]))
            if thr == float(s[ix["threshold"]]):
                conf = float(s[ix["confidence"]])
            else:  # history moved the threshold: redo the distance term (consistency_checker.py:249-271)
                cmv = float(s[ix["cross_modal_variance"]])
                valid = [float(s[ix[k]]) for k in _CC_KEYS if float(s[ix[k]]) > 0]
                cons = 1.0 - float(np.std(valid)) if len(valid) > 1 else 0.5
                conf = float(np.clip(np.mean([abs(overall - thr) / thr, cons, 1.0 - min(cmv, 1.0)]), 0.0, 1.0))
            res = {"is_adversarial": bool(overall < thr), "confidence": conf, "overall_score": overall,
                   "threshold": thr}
            self.detection_history.append(dict(res))
            self.threshold_history.append(thr)
            out.append(res)
        return out

    def make_decision(self, consistency_scores: Dict[str, float], return_details: bool = False) -> Dict[str, Any]:
        """consistency_checker.py:74-117 for one score dict.  The dict already holds the means / stds,
        so it is fed to kernel (b) as two-point similarity lists with exactly those moments."""
        g = consistency_scores.get

        def pair(mean, std):
            return np.array([[mean - std, mean + std]], np.float32)

        s0 = np.array([g("original_similarity", 0.0)], np.float32)
        sv = pair(g("text_variant_consistency", 0.0), g("text_variant_std", 0.0))
        sr = pair(g("retrieval_consistency", 0.0), g("retrieval_std", 0.0))
        sg = pair(g("generative_consistency", 0.0), g("generative_std", 0.0))
        r_cnt = np.array([0 if g("retrieval_consistency", 0.0) == 0 else 2], np.int32)
        g_cnt = np.array([0 if g("generative_consistency", 0.0) == 0 else 2], np.int32)
        scores, _ = N.Context.get().consistency_sims(self.params(2, 2, 2), s0, sv, sr, r_cnt, sg, g_cnt)
        row = scores[0].astype(np.float64)
        ix = N.SCORE_INDEX
        # cross_modal_variance is an INPUT of the reference's make_decision: honour the caller's value
        cmv = float(g("cross_modal_variance", 0.0))
        thr = float(self.base_threshold)
        if self.adaptive_threshold:
            if cmv > 0.1:
                thr += 0.1
            if np.mean([g("text_variant_std", 0), g("retrieval_std", 0), g("generative_std", 0)]) > 0.2:
                thr += 0.05
            thr = float(np.clip(thr, 0.1, 0.9))
        row[ix["threshold"]] = thr
        row[ix["cross_modal_variance"]] = cmv
        overall = float(row[ix["overall_score"]])
        valid = [float(g(k, 0.0)) for k in _CC_KEYS if float(g(k, 0.0)) > 0]
        cons = 1.0 - float(np.std(valid)) if len(valid) > 1 else 0.5
        row[ix["confidence"]] = float(np.clip(np.mean([abs(overall - thr) / thr, cons, 1.0 - min(cmv, 1.0)]), 0, 1))
        res = self.decide_from_kernel(row[None])[0]
        if return_details:
            res["details"] = self._get_detailed_analysis(consistency_scores, res["overall_score"], res["threshold"])
        return res

    def _get_detailed_analysis(self, scores: Dict[str, float], overall_score: float, threshold: float) -> Dict[str, Any]:
        """consistency_checker.py:274-317: per-score breakdown plus the risk / confidence labels (the label
        strings are output values consumers may match on, so they are the reference's)."""
        top = max(scores.values()) if scores else 0.0
        breakdown = {name: {"value": val, "normalized": val / top if top > 0 else 0,
                            "contribution": self.weights.get(name, 0) * val}
                     for name, val in scores.items() if val > 0}
        risks = [label for label, hit in (("高跨模态方差", scores.get("cross_modal_variance", 0) > 0.1),
                                          ("文本变体不一致", scores.get("text_variant_std", 0) > 0.2),
                                          ("低原始相似度", scores.get("original_similarity", 1) < 0.3)) if hit]
        trust = [label for label, hit in (("多模块验证", sum(1 for v in scores.values() if v > 0) >= 3),
                                          ("高检索一致性", scores.get("retrieval_consistency", 0) > 0.7),
                                          ("高生成一致性", scores.get("generative_consistency", 0) > 0.7)) if hit]
        return {"individual_scores": scores, "overall_score": overall_score, "threshold": threshold,
                "score_analysis": breakdown, "risk_factors": risks, "confidence_factors": trust}

    def update_weights(self, new_weights: Dict[str, float]):
        """consistency_checker.py:361-364."""
        self.weights.update(new_weights)

    def calibrate_threshold(self, validation_scores: List[Dict[str, float]], validation_labels: List[bool]) -> float:
        """consistency_checker.py:366-409: sweep 81 thresholds in [0.1, 0.9] for the best F1 (one kernel
        launch scores every validation sample; the sweep itself is a tiny host loop)."""
        g = [s.get for s in validation_scores]

        def col(k):
            return np.array([f(k, 0.0) for f in g], np.float32)

        def pair(m, s):
            return np.stack([col(m) - col(s), col(m) + col(s)], 1)

        scores, _ = N.Context.get().consistency_sims(
            self.params(2, 2, 2), col("original_similarity"), pair("text_variant_consistency", "text_variant_std"),
            pair("retrieval_consistency", "retrieval_std"),
            np.where(col("retrieval_consistency") == 0, 0, 2).astype(np.int32),
            pair("generative_consistency", "generative_std"),
            np.where(col("generative_consistency") == 0, 0, 2).astype(np.int32))
        overall = scores[:, N.SCORE_INDEX["overall_score"]].astype(np.float64)
        labels = np.asarray(validation_labels, dtype=bool)
        best_f1, best_thr = -1.0, self.base_threshold
        for thr in np.linspace(0.1, 0.9, 81):
            pred = overall < thr
            tp = int((pred & labels).sum())
            fp = int((pred & ~labels).sum())
            fn = int((~pred & labels).sum())
            prec = tp / (tp + fp) if tp + fp else 0.0
            rec = tp / (tp + fn) if tp + fn else 0.0
            f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
            if f1 > best_f1:
                best_f1, best_thr = f1, float(thr)
        self.base_threshold = best_thr
        return best_thr

    def get_statistics(self) -> Dict[str, Any]:
        """consistency_checker.py:319-353 (same keys; the reference's message when nothing was decided yet)."""
        if not self.detection_history:
            return {"message": "暂无检测历史"}

        def moments(key):
            v = np.array([d[key] for d in self.detection_history], dtype=np.float64)
            return {"mean": float(v.mean()), "std": float(v.std()), "min": float(v.min()), "max": float(v.max())}

        adv = sum(1 for d in self.detection_history if d["is_adversarial"])
        return {"total_detections": len(self.detection_history), "adversarial_detections": adv,
                "adversarial_rate": adv / len(self.detection_history), "score_statistics": moments("overall_score"),
                "threshold_statistics": moments("threshold"), "confidence_statistics": moments("confidence")}

    def reset_history(self):
        self.detection_history.clear()
        self.threshold_history.clear()


@dataclass
class DetectionConfig:
    """experiments/defenses/detector.py:19-43."""
    use_text_variants: bool = True
    text_variant_count: int = 5
    use_retrieval_ref: bool = True
    retrieval_top_k: int = 10
    retrieval_weight: float = 0.3
    use_generative_ref: bool = True
    generation_count: int = 3
    generation_weight: float = 0.4
    consistency_threshold: float = 0.5
    adaptive_threshold: bool = True
    voting_strategy: str = "weighted"
    device: str = "cuda"
    debug_mode: bool = False


class RetrievalReferenceIndex:
    """The retrieval side of experiments/defenses/retrieval_ref.py (features.npy database, top-
    `rerank_top_k` inner-product search, similarity floor 0.3, cut to `reference_count`, :173-216,246-290)
    for whole batches on the GPU."""

    def __init__(self, features, metadata: Optional[Sequence[Dict[str, Any]]] = None, reference_count: int = 5,
                 similarity_threshold: float = 0.3, rerank_top_k: int = 20, enable_reranking: bool = True):
        self.reference_features = _np(features)
        self.reference_metadata = list(metadata) if metadata is not None else []
        self.reference_count = reference_count
        self.similarity_threshold = similarity_threshold
        self.search_k = rerank_top_k if enable_reranking else reference_count
        self.gallery = Gallery(self.reference_features)

    def retrieve_batch(self, query_features):
        """[Q, d] or [Q, V, d] -> (sims, idx) [..., reference_count]; entries under the floor are -1."""
        k = min(self.search_k, max(1, len(self.gallery)))
        sims, idx = self.gallery.search(query_features, k, threshold=self.similarity_threshold)
        return sims[..., : self.reference_count], idx[..., : self.reference_count]

    def retrieve_references(self, query_features) -> List[Dict[str, Any]]:
        """One query row -> the reference's list of dicts (:250-262)."""
        sims, idx = self.retrieve_batch(_np(query_features).reshape(1, -1))
        return self.rows_to_dicts(sims[0], idx[0])

    def rows_to_dicts(self, sims_row, idx_row) -> List[Dict[str, Any]]:
        out = []
        for s, i in zip(sims_row, idx_row):
            if i >= 0:
                out.append({"index": int(i), "similarity": float(s),
                            "metadata": self.reference_metadata[i] if i < len(self.reference_metadata) else {},
                            "features": self.reference_features[i]})
        return out


@dataclass
class RetrievalRefConfig:
    """`RetrievalConfig` of experiments/defenses/retrieval_ref.py:20-31 (field for field; named apart from
    retrieval.RetrievalConfig of src/retrieval.py, which lives in this package too).  `use_faiss`,
    `faiss_index_type`, `nlist`, `nprobe` are accepted and reported; the search itself is always the exact
    inner-product search on the GPU (IVF / HNSW would be approximations of it)."""
    reference_count: int = 5
    similarity_threshold: float = 0.3
    use_faiss: bool = True
    faiss_index_type: str = "IVF"
    nlist: int = 100
    nprobe: int = 10
    device: str = "cuda"
    cache_size: int = 1000
    enable_reranking: bool = True
    rerank_top_k: int = 20


class RetrievalReferenceGenerator:
    """experiments/defenses/retrieval_ref.py:34-600 with the database resident in HBM: same constructor,
    `features.npy` + `metadata.json` database (:85-124, 442-457), result dicts (:250-262), cache, statistics
    keys (:542-570) and never-raise behaviour (:233-236).  `retrieve_references` is one exact search;
    `batch_retrieve_references` - a Python loop in the reference (:318-333) - encodes the uncached texts in
    ONE encoder call and searches them in ONE launch, then books them in order as the loop would."""

    def __init__(self, clip_model, reference_db_path: str, config: Optional[RetrievalRefConfig] = None):
        from pathlib import Path
        self.clip_model = clip_model
        self.config = config or RetrievalRefConfig()
        self.reference_db_path = Path(reference_db_path)
        self.reference_features: np.ndarray = np.empty((0, 512), np.float32)
        self.reference_metadata: List[Dict[str, Any]] = []
        self.feature_cache: Dict[int, List[Dict[str, Any]]] = {}
        self._index: Optional[RetrievalReferenceIndex] = None
        self.reset_statistics()
        self._load_reference_database()

    # -- database ---------------------------------------------------------------------------------
    def _load_reference_database(self):
        import json
        try:
            fp, mp = self.reference_db_path / "features.npy", self.reference_db_path / "metadata.json"
            if not fp.exists() or not mp.exists():
                logger.warning("reference database not found at %s: starting empty", self.reference_db_path)
                self._create_empty_database()
                return
            self.reference_features = np.load(fp)
            self.reference_metadata = json.loads(mp.read_text(encoding="utf-8"))
            self._rebuild_index()
        except Exception as e:  # noqa: BLE001
            logger.error("loading the reference database failed: %s", e)
            self._create_empty_database()

    def _create_empty_database(self):
        self.reference_features = np.empty((0, 512), np.float32)
        self.reference_metadata = []
        self._index = None

    def _rebuild_index(self):
        if self._index is not None:
            self._index.gallery.close()
        self._index = None
        if self.reference_features.shape[0] > 0:
            c = self.config
            self._index = RetrievalReferenceIndex(self.reference_features, self.reference_metadata, c.reference_count,
                                                  c.similarity_threshold, c.rerank_top_k, c.enable_reranking)
            self._index.reference_features = self.reference_features      # results hand out the stored rows

    def _save_reference_database(self):
        import json
        try:
            self.reference_db_path.mkdir(parents=True, exist_ok=True)
            np.save(self.reference_db_path / "features.npy", self.reference_features)
            (self.reference_db_path / "metadata.json").write_text(
                json.dumps(self.reference_metadata, ensure_ascii=False, indent=2), encoding="utf-8")
        except Exception as e:  # noqa: BLE001
            logger.error("saving the reference database failed: %s", e)

    def add_reference_features(self, features, metadata_list: Sequence[Dict[str, Any]]) -> bool:
        """Embedding-level insert (what add_reference_images does after its encoder calls, :419-438)."""
        try:
            new = np.asarray(_np(features)).reshape(len(metadata_list), -1)
            self.reference_features = new if self.reference_features.shape[0] == 0 else \
                np.vstack([self.reference_features, new.astype(self.reference_features.dtype)])
            self.reference_metadata.extend(metadata_list)
            self._rebuild_index()
            self._save_reference_database()
            return True
        except Exception as e:  # noqa: BLE001
            logger.error("adding references failed: %s", e)
            return False

    def add_reference_images(self, images, texts: List[str],
                             metadata_list: Optional[List[Dict[str, Any]]] = None) -> bool:
        """:366-440.  Tensors go to `clip_model.encode_image(t.unsqueeze(0))` as in the reference; paths and PIL
        images are handed to the encoder as they are (its own preprocessing applies; the reference's torchvision
        transform is upstream of the path)."""
        try:
            if len(images) != len(texts):
                raise ValueError("images and texts differ in number")
            feats, metas = [], []
            for i, (image, text) in enumerate(zip(images, texts)):
                if isinstance(image, str):
                    from PIL import Image
                    f = self.clip_model.encode_image([Image.open(image).convert("RGB")])
                elif hasattr(image, "unsqueeze"):
                    f = self.clip_model.encode_image(image.unsqueeze(0))
                else:
                    f = self.clip_model.encode_image([image])
                f = _np(f).reshape(1, -1)
                feats.append(f / np.linalg.norm(f, axis=-1, keepdims=True))
                meta = {"text": text, "image_path": str(image) if isinstance(image, str) else None,
                        "index": len(self.reference_metadata) + i}
                if metadata_list and i < len(metadata_list):
                    meta.update(metadata_list[i])
                metas.append(meta)
            return self.add_reference_features(np.vstack(feats), metas)
        except Exception as e:  # noqa: BLE001
            logger.error("adding reference images failed: %s", e)
            return False

    # -- retrieval --------------------------------------------------------------------------------
    def _encode_texts(self, texts: List[str]) -> np.ndarray:
        """:238-244 for a list: encode, L2-normalise, fp32."""
        f = _np(self.clip_model.encode_text(list(texts))).reshape(len(texts), -1)
        return (f / np.linalg.norm(f, axis=-1, keepdims=True)).astype(np.float32)

    def _encode_text(self, text: str) -> np.ndarray:
        return self._encode_texts([text])

    def _book(self, text: str, refs: List[Dict[str, Any]], seconds: float):
        """:218-231: running averages (over the count BEFORE this query), cache while there is room, counters."""
        st = self.retrieval_stats
        n = st["total_queries"]
        st["average_retrieval_time"] = (st["average_retrieval_time"] * n + seconds) / (n + 1)
        if refs:
            st["average_similarity"] = (st["average_similarity"] * n + float(np.mean([r["similarity"] for r in refs]))) / (n + 1)
        if len(self.feature_cache) < self.config.cache_size:
            self.feature_cache[hash(text)] = refs
        st["total_queries"] += 1
        if refs:
            st["successful_retrievals"] += 1

    def retrieve_references(self, text: str) -> List[Dict[str, Any]]:
        """:173-236."""
        import time
        t0 = time.time()
        key = hash(text)
        if key in self.feature_cache:
            self.retrieval_stats["cache_hits"] += 1
            return self.feature_cache[key]
        try:
            q = self._encode_text(text)
            if self._index is None:
                logger.warning("the reference database is empty")
                return []
            refs = self._index.retrieve_references(q)
            self._book(text, refs, time.time() - t0)
            return refs
        except Exception as e:  # noqa: BLE001
            logger.error("retrieving references failed: %s", e)
            return []

    def batch_retrieve_references(self, texts: List[str]) -> List[List[Dict[str, Any]]]:
        """:318-333 with one encoder call and one search launch for the texts the cache does not hold."""
        import time
        try:
            t0 = time.time()
            fresh: Dict[str, List[Dict[str, Any]]] = {}
            todo = [t for t in dict.fromkeys(texts) if hash(t) not in self.feature_cache]
            if todo and self._index is not None:
                q = self._encode_texts(todo)
                sims, idx = self._index.retrieve_batch(q)
                for t, s_row, i_row in zip(todo, sims, idx):
                    fresh[t] = self._index.rows_to_dicts(s_row, i_row)
            per_text = (time.time() - t0) / max(1, len(todo))
            out = []
            for t in texts:                                   # booked in order, exactly as the reference's loop would
                if hash(t) in self.feature_cache:
                    self.retrieval_stats["cache_hits"] += 1
                    out.append(self.feature_cache[hash(t)])
                elif t in fresh:
                    self._book(t, fresh[t], per_text)
                    out.append(fresh[t])
                else:
                    out.append([])
            return out
        except Exception as e:  # noqa: BLE001
            logger.error("batched reference retrieval failed: %s", e)
            return [self.retrieve_references(t) for t in texts]

    # -- bookkeeping ------------------------------------------------------------------------------
    def get_statistics(self) -> Dict[str, Any]:
        """:542-570 (same keys)."""
        st = dict(self.retrieval_stats)
        st["database_info"] = {"total_references": int(self.reference_features.shape[0]),
                               "feature_dimension": int(self.reference_features.shape[1]),
                               "use_faiss": self.config.use_faiss,
                               "faiss_index_type": self.config.faiss_index_type if self.config.use_faiss else None}
        n = st["total_queries"]
        st["success_rate"] = st["successful_retrievals"] / n if n > 0 else 0.0
        st["cache_hit_rate"] = st["cache_hits"] / n if n > 0 else 0.0
        st["config"] = {"reference_count": self.config.reference_count,
                        "similarity_threshold": self.config.similarity_threshold,
                        "cache_size": self.config.cache_size, "enable_reranking": self.config.enable_reranking}
        return st

    def reset_statistics(self):
        self.retrieval_stats = {"total_queries": 0, "successful_retrievals": 0, "cache_hits": 0,
                                "average_retrieval_time": 0.0, "average_similarity": 0.0}

    def clear_cache(self):
        self.feature_cache.clear()

    def update_config(self, new_config: RetrievalRefConfig):
        self.config = new_config
        self._rebuild_index()


class MultiModalDefenseDetector:
    """experiments/defenses/detector.py:46-170 with batched scoring.  `clip_model` must provide
    encode_image / encode_text; `text_variant_generator.generate_variants(text)`,
    `retrieval_index` (RetrievalReferenceIndex or anything with `.gallery` and `.retrieve_batch`) and
    `generative_generator.generate_references(text) -> list[image]` are optional, as in the reference."""

    def __init__(self, clip_model=None, qwen_model=None, sd_model=None, config: Optional[DetectionConfig] = None,
                 text_variant_generator=None, retrieval_index: Optional[RetrievalReferenceIndex] = None,
                 generative_generator=None):
        self.clip_model = clip_model
        self.config = config or DetectionConfig()
        self.text_variant_generator = text_variant_generator if self.config.use_text_variants else None
        self.retrieval_generator = retrieval_index if self.config.use_retrieval_ref else None
        self.generative_generator = generative_generator if self.config.use_generative_ref else None
        self.consistency_checker = ConsistencyChecker(threshold=self.config.consistency_threshold,
                                                      adaptive_threshold=self.config.adaptive_threshold,
                                                      voting_strategy=self.config.voting_strategy)

    def update_config(self, new_config: DetectionConfig):
        """experiments/defenses/detector.py:353-357."""
        self.config = new_config

    def get_statistics(self) -> Dict[str, Any]:
        """experiments/defenses/detector.py:359-375 (same keys)."""
        return {"config": self.config.__dict__,
                "components": {"text_variant_generator": self.text_variant_generator is not None,
                               "retrieval_generator": self.retrieval_generator is not None,
                               "generative_generator": self.generative_generator is not None,
                               "consistency_checker": self.consistency_checker is not None},
                "consistency_checker_stats": self.consistency_checker.get_statistics()}

    # -- embedding entry: everything after the encoders, one launch ------------------------------
    def detect_embeddings(self, image_emb, text_emb, variant_emb=None, generative_emb=None, generative_counts=None,
                          return_scores: bool = False):
        """image_emb/text_emb [Q,d], variant_emb [Q,V,d], generative_emb [Q,G,d].  Retrieval references
        come from searching every variant row (original text first when no variants) in the retrieval
        index.  Returns the list of reference-shaped result dicts (and the [Q,24] score matrix)."""
        img, txt = _np(image_emb), _np(text_emb)
        var = _np(variant_emb) if variant_emb is not None else None
        gen = _np(generative_emb) if generative_emb is not None else None
        q = img.shape[0]
        v = 0 if var is None else var.shape[1]
        g = 0 if gen is None else gen.shape[1]
        ret_gal, ret_idx = None, None
        if self.retrieval_generator is not None:
            rows = np.concatenate([txt[:, None, :], var], axis=1) if var is not None else txt[:, None, :]
            _, idx = self.retrieval_generator.retrieve_batch(rows)          # [Q, 1+V, reference_count]
            ret_gal, ret_idx = self.retrieval_generator.gallery, idx.reshape(q, -1)
        params = self.consistency_checker.params(v, self.config.retrieval_top_k if ret_idx is not None else 0, g)
        scores, _ = N.Context.get().consistency_emb(params, img, txt, var, ret_gallery=ret_gal, ret_idx=ret_idx,
                                                    gen=gen, g_cnt=generative_counts)
        decisions = self.consistency_checker.decide_from_kernel(scores)
        results = [{"is_adversarial": d["is_adversarial"], "confidence": d["confidence"],
                    "consistency_score": d["overall_score"]} for d in decisions]
        return (results, scores) if return_scores else results

    # -- reference-shaped entries -----------------------------------------------------------------
    def _encode(self, image, text):
        variants = [text]
        if self.text_variant_generator is not None:
            try:
                variants = [text] + list(self.text_variant_generator.generate_variants(text))
            except Exception as e:  # noqa: BLE001
                logger.warning("text variant generation failed: %s", e)
        temb = _np(self.clip_model.encode_text(variants))
        iemb = _np(self.clip_model.encode_image(image)).reshape(1, -1)
        gens = []
        if self.generative_generator is not None:
            try:
                for t in variants[: min(len(variants), 3)]:
                    gens.extend(self.generative_generator.generate_references(t))
                gens = gens[: self.config.generation_count]
            except Exception as e:  # noqa: BLE001
                logger.warning("generative reference generation failed: %s", e)
                gens = []
        gemb = np.concatenate([_np(self.clip_model.encode_image(gi)).reshape(1, -1) for gi in gens]) if gens else None
        return iemb[0], temb[0], temb[1:], gemb, variants

    def detect(self, image, text: str, return_details: bool = False) -> Dict[str, Any]:
        """experiments/defenses/detector.py:117-170."""
        return self.batch_detect([image], [text], return_details)[0]

    def batch_detect(self, images, texts: List[str], return_details: bool = False) -> List[Dict[str, Any]]:
        enc = [self._encode(im, tx) for im, tx in zip(images, texts)]
        buckets: Dict[Any, List[int]] = {}
        for i, e in enumerate(enc):
            buckets.setdefault((e[2].shape[0], 0 if e[3] is None else e[3].shape[0]), []).append(i)
        out: List[Optional[Dict[str, Any]]] = [None] * len(enc)
        for (v, g), members in sorted(buckets.items(), key=lambda kv: kv[1][0]):
            img = np.stack([enc[i][0] for i in members])
            txt = np.stack([enc[i][1] for i in members])
            var = np.stack([enc[i][2] for i in members]) if v else None
            gen = np.stack([enc[i][3] for i in members]) if g else None
            res, scores = self.detect_embeddings(img, txt, var, gen, return_scores=True)
            for row, i in enumerate(members):
                r = res[row]
                if return_details:
                    r["details"] = {"text_variants": enc[i][4],
                                    "consistency_scores": {n: float(scores[row, j]) for j, n in enumerate(N.SCORE_NAMES)}}
                out[i] = r
        return out  # type: ignore[return-value]
