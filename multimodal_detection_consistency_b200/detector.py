"""Drop-in for the reference's src/detector.py: `DetectorConfig` (:171-212) and `AdversarialDetector`
(:217) with the same result dictionary (:402-410), caches, statistics and never-raise behaviour
(:428-439), scored by kernel (b) (tvc_consistency_emb) instead of 1 + V + G scalar cosine calls and
host NumPy (:461-485, 528-542, 573-579, 653-682, 399).

Encoders stay upstream: pass `clip_model` (encode_text / encode_image -> [n, d]), `text_augmenter`
(`generate_variants(text) -> list[str]`) and `sd_generator`
(`generate_reference_images(text, num_images=) -> {'images': [...]}`), or use the embedding entry
`detect_embeddings`, which scores a whole batch in one launch.  `batch_detect` encodes once and
launches once per (V, G) group instead of looping (:711-734).  `detect` is the caller-side alias
(experiments/run_experiments.py:3253, README.md:442).
"""
from __future__ import annotations

import json
import logging
import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .batching import MicroBatcher

logger = logging.getLogger(__name__)

_METHOD_BITS = {"text_variants": 1, "sd_reference": 2, "consistency": 4}
_AGG = {"weighted_mean": 0, "mean": 1, "max": 2, "min": 3}


@dataclass
class DetectorConfig:
    """src/detector.py:171-212 (field for field)."""
    clip_model: str = "ViT-B/32"
    device: str = "cuda"
    detection_methods: List[str] = None
    use_text_variants: bool = True
    num_text_variants: int = 5
    text_similarity_threshold: float = 0.85
    use_sd_reference: bool = True
    num_reference_images: int = 3
    reference_similarity_threshold: float = 0.75
    consistency_threshold: float = 0.8
    consistency_weight: float = 0.5
    detection_threshold: float = 0.5
    adaptive_threshold: bool = True
    threshold_percentile: float = 95.0
    score_aggregation: str = "weighted_mean"
    enable_cache: bool = True
    cache_size: int = 1000
    batch_size: int = 32

    def __post_init__(self):
        if self.detection_methods is None:
            self.detection_methods = ["text_variants", "sd_reference", "consistency"]


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


class AdversarialDetector:
    def __init__(self, config: Optional[DetectorConfig] = None, clip_model=None, text_augmenter=None,
                 sd_generator=None):
        self.config = config or DetectorConfig()
        self.clip_model = clip_model
        self.text_augmenter = text_augmenter
        self.sd_generator = sd_generator
        self.detection_cache: Dict[str, Any] = {}
        self.threshold_cache: Dict[str, Any] = {}
        self.detection_stats = {"total_detections": 0, "cache_hits": 0, "detection_time": 0.0,
                                "method_usage": {m: 0 for m in self.config.detection_methods}}
        self._ctx = None
        self._tried = set()         # components whose lazy construction was already attempted
        # single-sample calls made concurrently are coalesced (batching.MicroBatcher)
        self.micro_batch = True
        self._batcher = MicroBatcher(lambda key, items: self._detect_group(items, list(key)), max_batch=256)

    # Component accessors, as in the reference (src/detector.py:252-343): what was injected wins; otherwise the
    # component is built lazily, on first use, from the HOST APPLICATION's packages - `src.models` (not shipped
    # with the reference, .gitignore:51), `src.text_augment`, `src.sd_ref` - with the reference's arguments.
    # That is what lets `AdversarialDetector(config)` be constructed by src/pipeline.py:324-327 unchanged.
    # A package that is not importable leaves the component None (tried once, not on every call).
    def _get_clip_model(self):
        if self.clip_model is None and "clip" not in self._tried:
            self._tried.add("clip")
            self.clip_model = self._initialize_clip_model(self.config.clip_model, self.config.device)
        return self.clip_model

    def _initialize_clip_model(self, model_name: str, device: str):
        """src/detector.py:257-279."""
        try:
            from src.models import CLIPConfig, CLIPModel  # type: ignore
            return CLIPModel(CLIPConfig(model_name=model_name, device=device))
        except Exception as e:  # noqa: BLE001
            logger.warning("no CLIP encoder available (%s); pass clip_model= or use detect_embeddings", e)
            return None

    def _get_text_augmenter(self):
        if self.text_augmenter is None and self.config.use_text_variants and "aug" not in self._tried:
            self._tried.add("aug")
            self.text_augmenter = self._initialize_text_augmenter(self.config.num_text_variants,
                                                                  self.config.text_similarity_threshold)
        return self.text_augmenter

    def _initialize_text_augmenter(self, num_variants: int, similarity_threshold: float):
        """src/detector.py:288-315 (same TextAugmentConfig arguments)."""
        try:
            from src.text_augment import TextAugmentConfig, TextAugmenter  # type: ignore
            return TextAugmenter(TextAugmentConfig(max_variants=num_variants,
                                                   min_similarity_threshold=similarity_threshold,
                                                   paraphrase_model="Qwen/Qwen2-7B-Instruct",
                                                   device=self.config.device))
        except Exception as e:  # noqa: BLE001
            logger.warning("text augmenter not available: %s", e)
            return None

    def _get_sd_generator(self):
        if self.sd_generator is None and self.config.use_sd_reference and "sd" not in self._tried:
            self._tried.add("sd")
            self.sd_generator = self._initialize_sd_generator(self.config.num_reference_images, self.config.device)
        return self.sd_generator

    def _initialize_sd_generator(self, num_images_per_prompt: int, device: str):
        """src/detector.py:326-343."""
        try:
            from src.sd_ref import SDReferenceConfig, SDReferenceGenerator  # type: ignore
            return SDReferenceGenerator(SDReferenceConfig(num_images_per_prompt=num_images_per_prompt, device=device))
        except Exception as e:  # noqa: BLE001
            logger.warning("SD reference generator not available: %s", e)
            return None

    def _context(self) -> N.Context:
        if self._ctx is None:
            self._ctx = N.Context.get()
        return self._ctx

    def _params(self, v: int, g: int, methods: Sequence[str]) -> N.DetectorParams:
        bits = 0
        for m in methods:
            bits |= _METHOD_BITS.get(m, 0)
        return N.default_params(n_variants=v, n_retrieval=0, n_generative=g, methods=bits,
                                aggregation=_AGG.get(self.config.score_aggregation, 1),
                                detection_threshold=self.config.detection_threshold)

    # ------------------------------------------------------------------ batched embedding entry
    def detect_embeddings(self, image_emb, text_emb, variant_emb=None, reference_emb=None, reference_counts=None,
                          methods: Optional[Sequence[str]] = None):
        """One launch for Q samples.  image_emb/text_emb [Q,d]; variant_emb [Q,V,d]; reference_emb [Q,G,d]
        (+ reference_counts [Q]).  Returns dict of arrays: is_adversarial [Q] bool, aggregated_score [Q],
        scores [Q, 24] (columns: multimodal_detection_consistency_b200.SCORE_NAMES), variant_similarities
        [Q,V], reference_similarities [Q,G].  numpy in -> numpy out; torch cuda in -> torch cuda out."""
        methods = list(methods or self.config.detection_methods)
        v = 0 if variant_emb is None else int(variant_emb.shape[1])
        g = 0 if reference_emb is None else int(reference_emb.shape[1])
        active = [m for m in methods if m == "consistency" or (m == "text_variants" and variant_emb is not None)
                  or (m == "sd_reference" and reference_emb is not None)]
        params = self._params(v, g, active)
        scores, flags, (sv, _, sg) = self._context().consistency_emb(
            params, image_emb, text_emb, variant_emb if v else None, gen=reference_emb if g else None,
            g_cnt=reference_counts, return_sims=True)
        return {"is_adversarial": (flags & N.FLAG_DET_ADV) != 0, "aggregated_score": scores[:, N.SCORE_INDEX["aggregated_score"]],
                "scores": scores, "variant_similarities": sv, "reference_similarities": sg, "methods_used": active}

    # ------------------------------------------------------------------ reference-shaped entry
    def _encode_samples(self, samples: Sequence[Tuple[Any, str]], methods):
        """Variants / references per sample (upstream generators), then ONE text-encoder call and ONE
        image-encoder call for the whole group (the reference encodes every string and image in its own
        call, src/detector.py:461-470,531)."""
        clip = self._get_clip_model()
        if clip is None:
            raise ValueError("no clip_model: pass one to AdversarialDetector or use detect_embeddings")
        aug = self._get_text_augmenter() if "text_variants" in methods else None
        sd = self._get_sd_generator() if "sd_reference" in methods else None
        texts, images, meta = [], [], []
        for image, text in samples:
            variants = list(aug.generate_variants(text) or []) if aug is not None else []
            refs, gen_time = [], 0.0
            if sd is not None:
                r = sd.generate_reference_images(text, num_images=self.config.num_reference_images)
                refs, gen_time = list(r.get("images", [])), r.get("generation_time", 0.0)
            meta.append((len(texts), 1 + len(variants), len(images), 1 + len(refs), variants, refs, gen_time))
            texts += [text] + variants
            images += [image] + refs
        temb_all = _np(clip.encode_text(texts))
        iemb_all = _np(clip.encode_image(images))
        return temb_all, iemb_all, meta

    def detect_adversarial(self, image, text: str, methods: Optional[List[str]] = None) -> Dict[str, Any]:
        """src/detector.py:345-439."""
        methods = methods or self.config.detection_methods
        try:
            key = self._get_cache_key(image, text, methods)
            if self.config.enable_cache and key in self.detection_cache:
                self.detection_stats["cache_hits"] += 1
                return self.detection_cache[key]
            t0 = time.time()
            if self.micro_batch:
                # concurrent callers (the pipeline's worker threads, src/pipeline.py:555-560) share one
                # encoder call and one kernel launch
                result = dict(self._batcher.submit((image, text), key=tuple(methods)))
            else:
                result = self._detect_group([(image, text)], methods)[0]
            result["detection_time"] = time.time() - t0
            if self.config.enable_cache:
                if len(self.detection_cache) >= self.config.cache_size:
                    del self.detection_cache[next(iter(self.detection_cache))]
                self.detection_cache[key] = result
            self.detection_stats["total_detections"] += 1
            self.detection_stats["detection_time"] += result["detection_time"]
            return result
        except Exception as e:  # noqa: BLE001
            logger.error("adversarial detection failed: %s", e)
            return {"is_adversarial": False, "aggregated_score": 0.0, "detection_scores": {}, "detection_details": {},
                    "detection_time": 0.0, "methods_used": methods, "threshold": self.config.detection_threshold,
                    "error": str(e)}

    detect = detect_adversarial

    def _detect_group(self, samples: Sequence[Tuple[Any, str]], methods: Sequence[str]) -> List[Dict[str, Any]]:
        """Encode every sample, bucket by (V, G), one kernel launch per bucket."""
        temb_all, iemb_all, meta = self._encode_samples(samples, methods)
        buckets: Dict[Tuple[int, int], List[int]] = {}
        for i, m in enumerate(meta):
            buckets.setdefault((m[1] - 1, m[3] - 1), []).append(i)
        results: List[Optional[Dict[str, Any]]] = [None] * len(samples)
        aug_on = "text_variants" in methods and self._get_text_augmenter() is not None
        sd_on = "sd_reference" in methods and self._get_sd_generator() is not None
        active = [m for m in methods if (m == "text_variants" and aug_on) or (m == "sd_reference" and sd_on)
                  or m == "consistency"]
        methods_list = list(methods)
        for (v, g), members in buckets.items():
            # the bucket's operands are gathered from the two encoder outputs by row index (one C-level gather each
            # instead of stacking per-sample views)
            t0s = np.fromiter((meta[i][0] for i in members), np.int64, len(members))
            i0s = np.fromiter((meta[i][2] for i in members), np.int64, len(members))
            img, txt = iemb_all[i0s], temb_all[t0s]
            var = temb_all[t0s[:, None] + np.arange(1, v + 1)] if v else None
            gen = iemb_all[i0s[:, None] + np.arange(1, g + 1)] if g else None
            params = self._params(v, g, active)
            scores, flags, (sv, _, sg) = self._context().consistency_emb(params, img, txt, var, gen=gen,
                                                                        return_sims=True)
            # plain Python floats once per bucket (ndarray.tolist), not one NumPy scalar conversion per field
            s_rows, f_rows = np.asarray(scores).tolist(), np.asarray(flags).tolist()
            sv_rows = np.asarray(sv).tolist() if v else None
            sg_rows = np.asarray(sg).tolist() if g else None
            for row, i in enumerate(members):
                results[i] = self._result_dict(s_rows[row], int(f_rows[row]), sv_rows[row] if v else [],
                                               sg_rows[row] if g else [], active, aug_on, sd_on, meta[i][6],
                                               list(methods_list))
        for m in active:
            if m in self.detection_stats["method_usage"]:
                self.detection_stats["method_usage"][m] += len(samples)
        return results  # type: ignore[return-value]

    def _result_dict(self, s, flag, sv, sg, active, aug_on, sd_on, gen_time, methods) -> Dict[str, Any]:
        """One sample's reference-shaped result (src/detector.py:396-427) from its row of kernel (b) outputs; `s`,
        `sv`, `sg` are lists of Python floats."""
        ix = N.SCORE_INDEX
        s0 = s[ix["original_similarity"]]
        scores: Dict[str, float] = {}
        details: Dict[str, Any] = {}
        if "text_variants" in active:
            if len(sv):
                mean_v, std_v = s[ix["text_variant_consistency"]], s[ix["text_variant_std"]]
                details["text_variants"] = {
                    "original_similarity": s0, "variant_similarities": list(sv),
                    "mean_variant_similarity": mean_v, "std_variant_similarity": std_v,
                    "consistency_score": 1.0 - abs(s0 - mean_v), "variability_score": 1.0 - std_v,
                    "num_variants": len(sv)}
            else:
                details["text_variants"] = {"error": "no text variants generated"}
            scores["text_variants"] = s[ix["det_text_variants"]]
        if "sd_reference" in active:
            if len(sg):
                details["sd_reference"] = {
                    "reference_similarities": list(sg),
                    "mean_similarity": s[ix["generative_consistency"]],
                    "max_similarity": s[ix["generative_max"]], "std_similarity": s[ix["generative_std"]],
                    "num_references": len(sg), "generation_time": gen_time}
            else:
                details["sd_reference"] = {"error": "no reference images generated"}
            scores["sd_reference"] = s[ix["det_sd_reference"]]
        if "consistency" in active:
            details["consistency"] = {"image_text_similarity": s0, "consistency_score": s0}
            scores["consistency"] = s[ix["det_consistency"]]
        return {"is_adversarial": bool(flag & N.FLAG_DET_ADV), "aggregated_score": s[ix["aggregated_score"]],
                "detection_scores": scores, "detection_details": details, "detection_time": 0.0,
                "methods_used": methods, "threshold": self.config.detection_threshold}

    def batch_detect(self, images: List[Any], texts: List[str], methods: Optional[List[str]] = None):
        """src/detector.py:711-734, batched: cached samples are answered from the cache, the rest are
        encoded and scored together."""
        methods = methods or self.config.detection_methods
        try:
            out: List[Optional[Dict[str, Any]]] = [None] * len(texts)
            todo = []
            for i, (im, tx) in enumerate(zip(images, texts)):
                key = self._get_cache_key(im, tx, methods)
                if self.config.enable_cache and key in self.detection_cache:
                    self.detection_stats["cache_hits"] += 1
                    out[i] = self.detection_cache[key]
                else:
                    todo.append((i, key))
            if todo:
                t0 = time.time()
                res = self._detect_group([(images[i], texts[i]) for i, _ in todo], methods)
                dt = (time.time() - t0) / len(todo)
                for (i, key), r in zip(todo, res):
                    r["detection_time"] = dt
                    out[i] = r
                    if self.config.enable_cache and len(self.detection_cache) < self.config.cache_size:
                        self.detection_cache[key] = r
                self.detection_stats["total_detections"] += len(todo)
                self.detection_stats["detection_time"] += dt * len(todo)
            return out
        except Exception as e:  # noqa: BLE001
            logger.error("batch detection failed: %s", e)
            return [self.detect_adversarial(im, tx, methods) for im, tx in zip(images, texts)]

    # ------------------------------------------------------------------ the rest of the surface
    def _get_cache_key(self, image, text: str, methods: Sequence[str]) -> str:
        """src/detector.py:684-709."""
        if hasattr(image, "detach"):
            ih = hash(image.detach().cpu().numpy().tobytes())
        else:
            try:
                ih = hash(np.asarray(image).tobytes())
            except Exception:  # noqa: BLE001
                ih = hash(repr(image))
        return f"{hash(text)}_{ih}_{hash(tuple(sorted(methods)))}"

    def compute_optimal_threshold(self, validation_data: List[Tuple[Any, Any, bool]],
                                  methods: Optional[List[str]] = None) -> float:
        """src/detector.py:736-770: the ROC point maximising TPR - FPR."""
        try:
            res = self.batch_detect([v[0] for v in validation_data], [v[1] for v in validation_data], methods)
            scores = np.array([r["aggregated_score"] for r in res], dtype=np.float64)
            labels = np.array([int(v[2]) for v in validation_data])
            from sklearn.metrics import roc_curve
            fpr, tpr, thr = roc_curve(labels, scores)
            return float(thr[int(np.argmax(tpr - fpr))])
        except Exception as e:  # noqa: BLE001
            logger.error("optimal threshold failed: %s", e)
            return self.config.detection_threshold

    def update_threshold(self, new_threshold: float):
        self.config.detection_threshold = new_threshold

    def evaluate_detection_performance(self, test_data: List[Tuple[Any, Any, bool]],
                                       methods: Optional[List[str]] = None) -> Dict[str, Any]:
        """src/detector.py:782-814 (accuracy / precision / recall / f1 computed here; the reference calls a
        `DetectionEvaluator.compute_metrics` that does not exist)."""
        try:
            res = self.batch_detect([t[0] for t in test_data], [t[1] for t in test_data], methods)
            pred = np.array([bool(r["is_adversarial"]) for r in res])
            lab = np.array([bool(t[2]) for t in test_data])
            tp, fp = int((pred & lab).sum()), int((pred & ~lab).sum())
            fn, tn = int((~pred & lab).sum()), int((~pred & ~lab).sum())
            prec = tp / (tp + fp) if tp + fp else 0.0
            rec = tp / (tp + fn) if tp + fn else 0.0
            return {"accuracy": (tp + tn) / max(1, len(lab)), "precision": prec, "recall": rec,
                    "f1": 2 * prec * rec / (prec + rec) if prec + rec else 0.0, "tp": tp, "fp": fp, "fn": fn, "tn": tn}
        except Exception as e:  # noqa: BLE001
            logger.error("performance evaluation failed: %s", e)
            return {}

    def clear_cache(self):
        """src/detector.py:817-823, plus: hand the device staging workspaces back."""
        self.detection_cache.clear()
        self.threshold_cache.clear()
        N.Context.release_all_workspaces()

    def get_stats(self) -> Dict[str, Any]:
        """src/detector.py:825-842 (same keys)."""
        return {"detection_stats": dict(self.detection_stats), "cache_size": len(self.detection_cache),
                "config": {"detection_methods": self.config.detection_methods,
                           "detection_threshold": self.config.detection_threshold,
                           "score_aggregation": self.config.score_aggregation,
                           "use_text_variants": self.config.use_text_variants,
                           "use_sd_reference": self.config.use_sd_reference}}

    def save_model(self, save_path: str):
        try:
            with open(save_path, "w") as f:
                json.dump({"config": self.config.__dict__, "detection_stats": self.detection_stats,
                           "threshold_cache": self.threshold_cache}, f, indent=2)
        except Exception as e:  # noqa: BLE001
            logger.error("save_model failed: %s", e)

    def load_model(self, load_path: str):
        try:
            with open(load_path) as f:
                data = json.load(f)
            for k, v in data["config"].items():
                if hasattr(self.config, k):
                    setattr(self.config, k, v)
            self.detection_stats.update(data.get("detection_stats", {}))
            self.threshold_cache.update(data.get("threshold_cache", {}))
        except Exception as e:  # noqa: BLE001
            logger.error("load_model failed: %s", e)


def create_adversarial_detector(config: Optional[DetectorConfig] = None, **components) -> AdversarialDetector:
    return AdversarialDetector(config or DetectorConfig(), **components)
