"""Drop-in for the reference's src/retrieval.py on the B200 hot path.

Same names, constructor/config dataclasses, return shapes and error behaviour as the reference
(`MultiModalRetriever` src/retrieval.py:316; `RetrievalConfig` :290; `FaissIndexManager` :89;
`RetrievalIndex` :196; `RetrievalResult` :40; `ConsistencyCalculator` :158), so
`src/pipeline.py:306-331,441-476` runs unchanged.  What differs is underneath: the FAISS / sklearn /
argsort call sites (:136,253,257-259,652-656,669-671,706-708) are one tcgen05 GEMM + top-k kernel
over an HBM-resident gallery, and the `batch_*` methods issue ONE launch for the whole batch instead
of a Python loop (:724-762).  The CLIP encoder stays an upstream producer: pass any object with
`encode_text(list, normalize=) / encode_image(list, normalize=)` returning [n, d] tensors/arrays.

Caller-side spellings that the reference's drivers use but its class does not define
 are provided as aliases: `search`, `search_by_text`, `retrieve`.
"""
from __future__ import annotations

import logging
import pickle
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import faiss_compat
from ._native import Context, Gallery
from .batching import MicroBatcher

logger = logging.getLogger(__name__)


def _to_numpy(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


# ----------------------------------------------------------------------------------------------
@dataclass
class IndexConfig:
    """src/retrieval.py:25-38, field for field.  Every index type is served by the exact search (the
    approximate families would only approximate it)."""
    index_type: str = "ivf"  # ivf, hnsw, flat
    dimension: int = 512
    n_clusters: int = 100
    n_links: int = 32
    ef_construction: int = 200
    ef_search: int = 100
    use_gpu: bool = True

    def __post_init__(self):
        if self.index_type not in ("ivf", "hnsw", "flat"):
            raise ValueError(f"Unsupported index type: {self.index_type}")


@dataclass
class RetrievalResult:
    """src/retrieval.py:40-86."""
    indices: List[int]
    similarities: List[float]
    items: List[Any] = None
    query_time: float = 0.0

    def to_dict(self) -> Dict[str, Any]:
        return {"indices": self.indices, "similarities": self.similarities, "query_time": self.query_time,
                "count": len(self.indices)}

    def filter_by_similarity(self, threshold: float) -> "RetrievalResult":
        keep = [i for i, s in enumerate(self.similarities) if s >= threshold]
        return RetrievalResult([self.indices[i] for i in keep], [self.similarities[i] for i in keep],
                               [self.items[i] for i in keep] if self.items else None, self.query_time)

    def get_top_k(self, k: int) -> "RetrievalResult":
        k = min(k, len(self.indices))
        return RetrievalResult(self.indices[:k], self.similarities[:k], self.items[:k] if self.items else None,
                               self.query_time)


class FaissIndexManager:
    """src/retrieval.py:89-155; `search` returns (similarities, indices)."""

    def __init__(self, config: IndexConfig):
        self.config = config
        self.index = None
        self.is_trained = False

    def create_index(self):
        if self.config.index_type not in ("flat", "ivf", "hnsw"):
            raise ValueError(f"Unsupported index type: {self.config.index_type}")
        self.index = faiss_compat.IndexFlatIP(self.config.dimension)
        return self.index

    def build_index(self, features: np.ndarray):
        if self.index is None:
            self.create_index()
        self.is_trained = True
        self.index.add(np.asarray(features, dtype=np.float32))

    def search(self, query_features: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        if self.index is None:
            raise ValueError("Index not built")
        return self.index.search(np.asarray(query_features, dtype=np.float32), k)

    def save_index(self, path: str):
        if self.index is None:
            raise ValueError("Index not built")
        faiss_compat.write_index(self.index, path)

    def load_index(self, path: str):
        self.index = faiss_compat.read_index(path)
        self.is_trained = True

    def add_to_index(self, features: np.ndarray):
        if self.index is None:
            raise ValueError("Index not built")
        self.index.add(np.asarray(features, dtype=np.float32))


class ConsistencyCalculator:
    """src/retrieval.py:158-193 (small host-side statistics on top-k result lists)."""

    def compute_similarity_distribution(self, similarities: np.ndarray) -> Dict[str, float]:
        s = np.asarray(similarities)
        return {"mean": float(np.mean(s)), "std": float(np.std(s)), "min": float(np.min(s)), "max": float(np.max(s)),
                "median": float(np.median(s))}

    def compute_consistency_score(self, similarities1: np.ndarray, similarities2: np.ndarray) -> float:
        c = np.corrcoef(similarities1, similarities2)[0, 1]
        return 0.0 if np.isnan(c) else float(c)

    def compute_top_k_consistency(self, indices1: np.ndarray, indices2: np.ndarray, k: int) -> float:
        a = set(np.asarray(indices1)[:k].tolist())
        b = set(np.asarray(indices2)[:k].tolist())
        return len(a & b) / k

    def compute_rank_correlation(self, indices1: np.ndarray, indices2: np.ndarray) -> float:
        try:
            from scipy.stats import spearmanr
            c, _ = spearmanr(indices1, indices2)
            return 0.0 if np.isnan(c) else float(c)
        except Exception:
            return 0.0


class RetrievalIndex:
    """src/retrieval.py:196-287; `search` returns (indices, scores) — note the order (:254)."""

    def __init__(self, index_type: str = "faiss", dimension: int = 512):
        self.index_type = index_type
        self.dimension = dimension
        self.index = None
        self.features = None

    def build_index(self, features: np.ndarray):
        self.features = np.asarray(features, dtype=np.float32)
        self.index = Gallery(self.features, normalize=self.index_type != "faiss")

    def search(self, query_features: np.ndarray, top_k: int = 10) -> Tuple[np.ndarray, np.ndarray]:
        if self.index is None:
            raise ValueError("index not built")
        q = np.asarray(query_features, dtype=np.float32)
        sims, idx = self.index.search(q, top_k, normalize_queries=self.index_type != "faiss")
        if self.index_type == "faiss":
            return idx, sims
        keep = idx[0] >= 0  # the numpy branch returns 1-D arrays for the first query (:257-260)
        return idx[0][keep], sims[0][keep]

    def add_items(self, features: np.ndarray):
        features = np.asarray(features, dtype=np.float32)
        if self.index is None:
            return self.build_index(features)
        self.index.append(features)
        self.features = np.vstack([self.features, features])


# ----------------------------------------------------------------------------------------------
@dataclass
class RetrievalConfig:
    """src/retrieval.py:290-313 (field for field)."""
    clip_model: str = "ViT-B/32"
    device: str = "cuda"
    batch_size: int = 256
    top_k: int = 10
    similarity_metric: str = "cosine"
    index_type: str = "faiss"
    faiss_index_type: str = "IndexFlatIP"
    n_clusters: int = 100
    enable_cache: bool = True
    cache_dir: Optional[str] = None
    normalize_features: bool = True
    use_gpu_index: bool = True


class MultiModalRetriever:
    """Text<->image retrieval over encoder embeddings with the gallery resident in HBM."""

    def __init__(self, config: Optional[RetrievalConfig] = None, clip_model=None):
        self.config = config or RetrievalConfig()
        self.device = self.config.device
        self.clip_model = clip_model if clip_model is not None else self._initialize_clip_model()
        self.image_features = None
        self.text_features = None
        self.image_paths: List[str] = []
        self.texts: List[str] = []
        self.image_index: Optional[Gallery] = None
        self.text_index: Optional[Gallery] = None
        self.feature_cache: Dict[str, Any] = {}
        self.retrieval_cache: Dict[str, Any] = {}
        # single-query calls made concurrently (the pipeline's worker threads, src/pipeline.py:555-560)
        # are coalesced into one encoder call + one search launch (batching.MicroBatcher)
        self.micro_batch = True
        # (the raising forms: a failing round is re-run item by item by the batcher, so one bad query only
        # empties its own caller's result - the reference fails per query, src/retrieval.py:574-576)
        self._t2i_batcher = MicroBatcher(lambda k, texts: self._batch_t2i(texts, k))
        self._i2t_batcher = MicroBatcher(lambda k, images: self._batch_i2t(images, k))

    def _initialize_clip_model(self):
        """The reference builds `src.models.CLIPModel` here (src/retrieval.py:347-369); that package is
        not shipped (.gitignore:51).  Use it when the host application provides it."""
        try:
            from src.models import CLIPConfig, CLIPModel  # type: ignore
            return CLIPModel(CLIPConfig(model_name=self.config.clip_model, device=self.config.device,
                                        batch_size=self.config.batch_size, normalize=self.config.normalize_features))
        except Exception as e:  # noqa: BLE001
            logger.warning("no CLIP encoder available (%s); pass clip_model= or use the *_features APIs", e)
            return None

    # -- index construction ------------------------------------------------------------------
    def _metric_flags(self):
        cosine = self.config.similarity_metric == "cosine" and self.config.index_type != "faiss"
        return cosine

    def _build_faiss_index(self, features: np.ndarray) -> Optional[Gallery]:
        """src/retrieval.py:477-525: returns None on any failure."""
        try:
            f = np.ascontiguousarray(features, dtype=np.float32)
            return Gallery(f, normalize=self._metric_flags())
        except Exception as e:  # noqa: BLE001
            logger.error("index build failed: %s", e)
            return None

    def build_image_index_from_features(self, image_features, image_paths: Sequence[str],
                                        save_path: Optional[str] = None) -> np.ndarray:
        self.image_features = np.ascontiguousarray(_to_numpy(image_features), dtype=np.float32)
        self.image_paths = list(image_paths)
        self.image_index = self._build_faiss_index(self.image_features)
        if save_path:
            self.save_image_index(save_path)
        return self.image_features

    def build_text_index_from_features(self, text_features, texts: Sequence[str],
                                       save_path: Optional[str] = None) -> np.ndarray:
        self.text_features = np.ascontiguousarray(_to_numpy(text_features), dtype=np.float32)
        self.texts = list(texts)
        self.text_index = self._build_faiss_index(self.text_features)
        if save_path:
            self.save_text_index(save_path)
        return self.text_features

    def build_image_index(self, image_paths: List[str], save_path: Optional[str] = None) -> np.ndarray:
        """src/retrieval.py:371-432 (raises on failure, like the reference)."""
        from PIL import Image
        images, valid = [], []
        for p in image_paths:
            try:
                images.append(Image.open(p).convert("RGB"))
                valid.append(p)
            except Exception as e:  # noqa: BLE001
                logger.warning("cannot load image %s: %s", p, e)
        if not images:
            raise ValueError(f"no image could be loaded out of {len(image_paths)}")
        feats = self.clip_model.encode_image(images, normalize=self.config.normalize_features)
        return self.build_image_index_from_features(feats, valid, save_path)

    def build_text_index(self, texts: List[str], save_path: Optional[str] = None) -> np.ndarray:
        """src/retrieval.py:434-475."""
        feats = self.clip_model.encode_text(texts, normalize=self.config.normalize_features)
        return self.build_text_index_from_features(feats, texts, save_path)

    # -- search ------------------------------------------------------------------------------
    def _search_index(self, index: Optional[Gallery], query_features: np.ndarray,
                      top_k: int) -> Tuple[np.ndarray, np.ndarray]:
        """src/retrieval.py:636-680: one query row -> (indices [<=k], scores [<=k]); empty arrays on error."""
        try:
            if index is None:
                feats = self.image_features if self.image_features is not None else self.text_features
                if feats is None:
                    raise ValueError("no feature matrix available")
                index = Gallery(np.ascontiguousarray(feats, dtype=np.float32), normalize=True)
                sims, idx = index.search(np.asarray(query_features, np.float32), top_k, normalize_queries=True)
            else:
                sims, idx = index.search(np.asarray(query_features, np.float32), top_k,
                                         normalize_queries=self._metric_flags())
            if self.config.index_type == "faiss":
                return idx[0], sims[0]          # FAISS keeps the -1 padding (:652-656)
            keep = idx[0] >= 0
            return idx[0][keep], sims[0][keep]
        except Exception as e:  # noqa: BLE001
            logger.error("index search failed: %s", e)
            return np.array([]), np.array([])

    def search_features(self, query_features, top_k: Optional[int] = None, index: str = "image"):
        """Batched entry the reference lacks: [Q, d] or [Q, V, d] rows -> (sims, idx) with the same
        leading shape.  numpy in/out or torch-cuda in/out (no host round trip)."""
        gal = self.image_index if index == "image" else self.text_index
        if gal is None:
            raise ValueError(f"{index} index not built")
        return gal.search(query_features, top_k or self.config.top_k, normalize_queries=self._metric_flags())

    def retrieve_images_by_text(self, query_text: str, top_k: Optional[int] = None) -> Tuple[List[str], List[float]]:
        """src/retrieval.py:527-576: ([], []) on any error."""
        try:
            top_k = top_k or self.config.top_k
            key = f"text2img_{query_text}_{top_k}"
            if self.config.enable_cache and key in self.retrieval_cache:
                return self.retrieval_cache[key]
            if self.image_features is None or self.image_index is None:
                raise ValueError("image index not built")
            if self.micro_batch:
                return self._t2i_batcher.submit(query_text, key=top_k)
            q = _to_numpy(self.clip_model.encode_text([query_text], normalize=self.config.normalize_features))
            idx, scores = self._search_index(self.image_index, q, top_k)
            paths = [self.image_paths[i] for i in idx if i >= 0]
            result = (paths, [float(s) for s, i in zip(scores, idx) if i >= 0])
            if self.config.enable_cache:
                self.retrieval_cache[key] = result
            return result
        except Exception as e:  # noqa: BLE001
            logger.error("text->image retrieval failed: %s", e)
            return [], []

    def retrieve_texts_by_image(self, query_image, top_k: Optional[int] = None) -> Tuple[List[str], List[float]]:
        """src/retrieval.py:578-634."""
        try:
            top_k = top_k or self.config.top_k
            if self.text_features is None or self.text_index is None:
                raise ValueError("text index not built")
            if isinstance(query_image, str):
                from PIL import Image
                image = Image.open(query_image).convert("RGB")
                key = f"img2text_{query_image}_{top_k}"
            else:
                image, key = query_image, f"img2text_pil_{id(query_image)}_{top_k}"
            if self.config.enable_cache and key in self.retrieval_cache:
                return self.retrieval_cache[key]
            if self.micro_batch:
                return self._i2t_batcher.submit(query_image, key=top_k)   # cached there, under the same key
            q = _to_numpy(self.clip_model.encode_image([image], normalize=self.config.normalize_features))
            idx, scores = self._search_index(self.text_index, q, top_k)
            result = ([self.texts[i] for i in idx if i >= 0], [float(s) for s, i in zip(scores, idx) if i >= 0])
            if self.config.enable_cache:
                self.retrieval_cache[key] = result
            return result
        except Exception as e:  # noqa: BLE001
            logger.error("image->text retrieval failed: %s", e)
            return [], []

    def _batch_t2i(self, query_texts: List[str], top_k: int):
        """One encoder call and ONE search launch for the whole batch; raises on failure."""
        if self.image_index is None:
            raise ValueError("image index not built")
        todo = [t for t in dict.fromkeys(query_texts)
                if not (self.config.enable_cache and f"text2img_{t}_{top_k}" in self.retrieval_cache)]
        fresh = {}
        if todo:
            q = _to_numpy(self.clip_model.encode_text(todo, normalize=self.config.normalize_features))
            sims, idx = self.search_features(np.ascontiguousarray(q, np.float32), top_k)
            for t, s_row, i_row in zip(todo, sims, idx):
                keep = i_row >= 0
                fresh[t] = ([self.image_paths[i] for i in i_row[keep]], [float(s) for s in s_row[keep]])
                if self.config.enable_cache:
                    self.retrieval_cache[f"text2img_{t}_{top_k}"] = fresh[t]
        return [fresh[t] if t in fresh else self.retrieval_cache[f"text2img_{t}_{top_k}"] for t in query_texts]

    def _isolated(self, fn, items, top_k, what):
        """The batch as one call; if that fails, item by item, so only the offending items come back empty
        (the reference's batch_* are loops over the never-raising single call, src/retrieval.py:724-762)."""
        try:
            return fn(items, top_k)
        except Exception as e:  # noqa: BLE001
            logger.error("batched %s retrieval failed: %s", what, e)
            if len(items) <= 1:
                return [([], []) for _ in items]
        out = []
        for it in items:
            try:
                out.append(fn([it], top_k)[0])
            except Exception as e:  # noqa: BLE001
                logger.error("%s retrieval failed: %s", what, e)
                out.append(([], []))
        return out

    def batch_retrieve_images_by_texts(self, query_texts: List[str], top_k: Optional[int] = None):
        """src/retrieval.py:724-742, but one encoder call and ONE search launch for the whole batch."""
        return self._isolated(self._batch_t2i, list(query_texts), top_k or self.config.top_k, "text->image")

    def batch_retrieve_texts_by_images(self, query_images: List[Any], top_k: Optional[int] = None):
        """src/retrieval.py:744-762 with one launch."""
        return self._isolated(self._batch_i2t, list(query_images), top_k or self.config.top_k, "image->text")

    def _batch_i2t(self, query_images: List[Any], top_k: int):
        if self.text_index is None:
            raise ValueError("text index not built")
        # same cache keys as the single call (:596-600), so either entry point serves the other
        keys = [f"img2text_{im}_{top_k}" if isinstance(im, str) else f"img2text_pil_{id(im)}_{top_k}"
                for im in query_images]
        cache = self.retrieval_cache if self.config.enable_cache else {}
        fresh: Dict[str, Any] = {}
        todo = [(key, im) for key, im in dict(zip(keys, query_images)).items() if key not in cache]
        if todo:
            images = []
            for _, im in todo:
                if isinstance(im, str):
                    from PIL import Image
                    im = Image.open(im).convert("RGB")
                images.append(im)
            q = _to_numpy(self.clip_model.encode_image(images, normalize=self.config.normalize_features))
            sims, idx = self.search_features(np.ascontiguousarray(q, np.float32), top_k, index="text")
            for (key, _), s_row, i_row in zip(todo, sims, idx):
                fresh[key] = ([self.texts[i] for i in i_row[i_row >= 0]], [float(s) for s in s_row[i_row >= 0]])
                if self.config.enable_cache:
                    self.retrieval_cache[key] = fresh[key]
        return [fresh[key] if key in fresh else cache[key] for key in keys]

    # caller-side spellings (experiments/run_experiments.py:3143, README.md:368,810)
    def retrieve(self, text: str, k: Optional[int] = None, top_k: Optional[int] = None):
        return self.retrieve_images_by_text(text, k or top_k)

    search = retrieve
    search_by_text = retrieve

    def compute_similarity_matrix(self, text_features: Optional[np.ndarray] = None,
                                  image_features: Optional[np.ndarray] = None) -> np.ndarray:
        """src/retrieval.py:682-722 (raises like the reference).  cosine / dot_product run on the tensor
        cores; the 1/(1+||t-i||) 'euclidean' variant is derived from the same GEMM."""
        tf = self.text_features if text_features is None else np.asarray(text_features, np.float32)
        imf = self.image_features if image_features is None else np.asarray(image_features, np.float32)
        if tf is None or imf is None:
            raise ValueError("text or image features missing")
        metric = self.config.similarity_metric
        if metric not in ("cosine", "dot_product", "euclidean"):
            raise ValueError(f"unsupported similarity metric: {metric}")
        own = image_features is None and self.image_index is not None and not self._metric_flags() \
            and metric != "cosine"
        gal = self.image_index if own else Gallery(np.ascontiguousarray(imf, np.float32), normalize=metric == "cosine")
        dots = gal.similarity_matrix(np.ascontiguousarray(tf, np.float32), normalize_queries=metric == "cosine")
        if metric != "euclidean":
            return dots
        t2 = (tf.astype(np.float64) ** 2).sum(1)[:, None]
        i2 = (imf.astype(np.float64) ** 2).sum(1)[None, :]
        return (1.0 / (1.0 + np.sqrt(np.maximum(t2 + i2 - 2.0 * dots, 0.0)))).astype(np.float32)

    # -- persistence (src/retrieval.py:764-882): pickle {features, paths|texts, config} + .faiss sidecar --
    def _save(self, save_path, payload, features):
        save_path = Path(save_path)
        save_path.parent.mkdir(parents=True, exist_ok=True)
        with open(save_path, "wb") as f:
            pickle.dump(payload, f)
        if features is not None:
            # the sidecar the reference writes with faiss.write_index (:781-783), in FAISS's flat layout
            with open(save_path.with_suffix(".faiss"), "wb") as f:
                f.write(faiss_compat._pack_flat(np.asarray(features, np.float32), int(np.asarray(features).shape[1])))

    def save_image_index(self, save_path: str):
        self._save(save_path, {"image_features": self.image_features, "image_paths": self.image_paths,
                               "config": self.config}, self.image_features if self.image_index is not None else None)

    def save_text_index(self, save_path: str):
        self._save(save_path, {"text_features": self.text_features, "texts": self.texts, "config": self.config},
                   self.text_features if self.text_index is not None else None)

    @staticmethod
    def _load_payload(load_path) -> Dict[str, Any]:
        """Files written by the reference pickle its own `src.retrieval.RetrievalConfig`; when that module
        is not importable the class is mapped onto ours instead of failing the load."""
        class _Unpickler(pickle.Unpickler):
            def find_class(self, module, name):
                try:
                    return super().find_class(module, name)
                except (ImportError, AttributeError):
                    if name == "RetrievalConfig":
                        return RetrievalConfig
                    if name == "IndexConfig":
                        return IndexConfig
                    raise
        with open(Path(load_path), "rb") as f:
            return _Unpickler(f).load()

    def _load(self, load_path, feat_key):
        data = self._load_payload(load_path)
        feats = data.get(feat_key)
        side = Path(load_path).with_suffix(".faiss")
        if feats is None and side.exists():          # features only in the sidecar
            with open(side, "rb") as f:
                _, _, feats = faiss_compat._unpack_flat(f.read())
        return data, feats

    def load_image_index(self, load_path: str):
        data, feats = self._load(load_path, "image_features")
        self.build_image_index_from_features(feats, data["image_paths"])

    def load_text_index(self, load_path: str):
        data, feats = self._load(load_path, "text_features")
        self.build_text_index_from_features(feats, data["texts"])

    def clear_cache(self):
        """src/retrieval.py:884-890, plus: hand the device staging workspaces back (the galleries stay)."""
        self.feature_cache.clear()
        self.retrieval_cache.clear()
        Context.release_all_workspaces()

    def get_stats(self) -> Dict[str, Any]:
        """src/retrieval.py:892-912 (same keys)."""
        return {
            "image_count": len(self.image_paths) if self.image_paths else 0,
            "text_count": len(self.texts) if self.texts else 0,
            "image_features_shape": self.image_features.shape if self.image_features is not None else None,
            "text_features_shape": self.text_features.shape if self.text_features is not None else None,
            "feature_cache_size": len(self.feature_cache),
            "retrieval_cache_size": len(self.retrieval_cache),
            "config": {"clip_model": self.config.clip_model, "top_k": self.config.top_k,
                       "similarity_metric": self.config.similarity_metric, "index_type": self.config.index_type},
        }


def create_retriever(config: Optional[RetrievalConfig] = None, clip_model=None) -> MultiModalRetriever:
    """src/retrieval.py:915-927."""
    return MultiModalRetriever(config or RetrievalConfig(), clip_model=clip_model)
