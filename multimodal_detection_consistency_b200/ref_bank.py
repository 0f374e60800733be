"""Drop-in for the reference's src/ref_bank.py with the lookups on the B200.

`ReferenceBankConfig` (src/ref_bank.py:24-44), `ReferenceItem` (:47-83) and `ReferenceBank` (:86) keep
their fields, methods, return shapes, locking and error behaviour.  The arithmetic changes place:
the reference rebuilds `np.array([ref.vector ...])` and runs dot/(|r||q|+1e-8) + where + argsort on
the host for every query (:462-484,197-203); here the vectors live L2-normalised in an HBM gallery
and `query_similar` is one tvc_search launch with the `>= threshold` filter in its epilogue.
`query_similar_batch` (new) serves many queries per launch.  Insert-time de-duplication is an exact
max-similarity query over the whole bank instead of the reference's unseeded random 100-sample
(:341-363) — identical whenever the bank holds <= 100 vectors, deterministic and stricter beyond.
Clustering and the eviction policies are host bookkeeping as in the reference (out of the GPU path).
"""
from __future__ import annotations

import json
import os
import logging
import pickle
import time
from collections import OrderedDict, deque
from dataclasses import asdict, dataclass
from pathlib import Path
from threading import Lock
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from ._native import Gallery

logger = logging.getLogger(__name__)


@dataclass
class ReferenceBankConfig:
    max_size: int = 10000
    similarity_threshold: float = 0.9
    clustering_method: str = "kmeans"
    num_clusters: int = 100
    update_strategy: str = "fifo"
    persistence_enabled: bool = True
    save_path: str = "./cache/ref_bank"
    auto_clustering: bool = True
    clustering_interval: int = 1000
    feature_dim: int = 512

    def __post_init__(self):
        if self.clustering_method not in ("kmeans", "dbscan", "none"):
            raise ValueError(f"unsupported clustering method: {self.clustering_method}")
        if self.update_strategy not in ("fifo", "lru", "random", "similarity"):
            raise ValueError(f"unsupported update strategy: {self.update_strategy}")


@dataclass
class ReferenceItem:
    vector: np.ndarray
    metadata: Dict[str, Any]
    timestamp: float
    access_count: int = 0
    cluster_id: Optional[int] = None
    similarity_scores: Optional[Dict[str, float]] = None

    def __post_init__(self):
        if self.similarity_scores is None:
            self.similarity_scores = {}

    def to_dict(self) -> Dict[str, Any]:
        return {"vector": np.asarray(self.vector).tolist(), "metadata": self.metadata, "timestamp": self.timestamp,
                "access_count": self.access_count, "cluster_id": self.cluster_id,
                "similarity_scores": self.similarity_scores}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "ReferenceItem":
        return cls(vector=np.array(data["vector"]), metadata=data["metadata"], timestamp=data["timestamp"],
                   access_count=data.get("access_count", 0), cluster_id=data.get("cluster_id"),
                   similarity_scores=data.get("similarity_scores", {}))


class ReferenceBank:
    def __init__(self, config: ReferenceBankConfig):
        self.config = config
        self.references: List[ReferenceItem] = []
        self.clusters: Dict[int, List[int]] = {}
        self.cluster_centers: Optional[np.ndarray] = None
        # LRU order of `references` positions.  The reference keeps a deque and pays O(B) `in` + `remove` per hit
        # (src/ref_bank.py:212-214); here the order lives in an OrderedDict (O(1) move-to-end / pop-oldest) and
        # `access_order` shows it as the deque observers expect
        self._access: "OrderedDict[int, None]" = OrderedDict()
        self._lock = Lock()
        self.stats = self._fresh_stats()
        # device mirror: gallery row r holds the (normalised) vector of self._row_items[r]
        self._pos_cache: Optional[Dict[int, int]] = None
        self._gallery: Optional[Gallery] = None
        self._row_items: List[ReferenceItem] = []
        self._row_of: Dict[int, int] = {}
        # write-through persistence: an append-only journal next to the reference's four JSON files, folded
        # into them every `journal_compact_every` operations (the reference rewrites all of references.json
        # on every insert, src/ref_bank.py:163,505-516 - O(B) per insert)
        self.journal_compact_every = 256
        self._journal_ops = 0
        self._seq = 0                   # operations journalled since the bank was created (persisted)
        # KMeans with the assignment step on the GPU; False = scikit-learn end to end
        self.gpu_kmeans = True
        if self.config.persistence_enabled:
            Path(self.config.save_path).mkdir(parents=True, exist_ok=True)
        self._load_from_disk()

    @property
    def access_order(self) -> deque:
        return deque(self._access)

    @access_order.setter
    def access_order(self, order):
        self._access = OrderedDict((int(i), None) for i in order)

    @staticmethod
    def _fresh_stats():
        return {"total_added": 0, "total_removed": 0, "total_queries": 0, "clustering_count": 0,
                "last_clustering_time": None}

    # ------------------------------------------------------------------ device mirror
    def _dev_append(self, items: List[ReferenceItem]):
        if not items:
            return
        rows = np.stack([np.asarray(it.vector, dtype=np.float32).ravel() for it in items])
        if self._gallery is None:
            self._gallery = Gallery(rows, normalize=True, capacity=min(self.config.max_size, 1 << 20))
        else:
            self._gallery.append(rows)
        for it in items:
            self._row_of[id(it)] = len(self._row_items)
            self._row_items.append(it)

    def _dev_remove(self, item: ReferenceItem):
        r = self._row_of.pop(id(item))
        last = len(self._row_items) - 1
        if r != last:
            self._gallery.move_row(last, r)
            moved = self._row_items[last]
            self._row_items[r] = moved
            self._row_of[id(moved)] = r
        self._row_items.pop()
        self._gallery.truncate(last)

    def _dev_rebuild(self):
        if self._gallery is not None:
            self._gallery.close()
        self._gallery, self._row_items, self._row_of = None, [], {}
        self._dev_append(list(self.references))

    def _search(self, queries: np.ndarray, top_k: int, threshold: float):
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32).reshape(-1, queries.shape[-1]))
        return self._gallery.search(q, top_k, threshold=threshold, normalize_queries=True)

    # ------------------------------------------------------------------ public API
    def add_reference(self, vector: np.ndarray, metadata: Dict[str, Any]) -> bool:
        """src/ref_bank.py:123-170."""
        try:
            with self._lock:
                if self._is_too_similar(vector):
                    return False
                item = ReferenceItem(vector=np.array(vector, copy=True), metadata=dict(metadata), timestamp=time.time())
                if len(self.references) >= self.config.max_size:
                    self._remove_reference()
                self.references.append(item)
                if self._pos_cache is not None:
                    self._pos_cache[id(item)] = len(self.references) - 1
                self._dev_append([item])
                self.stats["total_added"] += 1
                clustered = False
                if self.config.auto_clustering and len(self.references) % self.config.clustering_interval == 0:
                    clustered = self._perform_clustering()
                if self.config.persistence_enabled:
                    if clustered:
                        self._save_to_disk()              # cluster ids changed on many items: fold everything
                    else:
                        self._journal({"op": "add", "item": item.to_dict()})
                return True
        except Exception as e:  # noqa: BLE001
            logger.error("add_reference failed: %s", e)
            return False

    def query_similar(self, query_vector: np.ndarray, top_k: int = 10,
                      similarity_threshold: Optional[float] = None) -> List[Tuple[ReferenceItem, float]]:
        """src/ref_bank.py:172-224; [] on error.  `threshold = arg or config` keeps the reference's
        falsy-zero behaviour (:191)."""
        try:
            with self._lock:
                if not self.references:
                    return []
                threshold = similarity_threshold or self.config.similarity_threshold
                q = np.asarray(query_vector, dtype=np.float32).reshape(1, -1)
                sims, rows = self._search(q, top_k, float(threshold))
                return self._collect(sims[0], rows[0])
        except Exception as e:  # noqa: BLE001
            logger.error("query_similar failed: %s", e)
            return []

    def query_similar_batch(self, query_vectors: np.ndarray, top_k: int = 10,
                            similarity_threshold: Optional[float] = None) -> List[List[Tuple[ReferenceItem, float]]]:
        """Many queries, one launch (the batched path the reference lacks)."""
        try:
            with self._lock:
                qv = np.asarray(query_vectors, dtype=np.float32)
                if not self.references:
                    return [[] for _ in range(len(qv))]
                threshold = similarity_threshold or self.config.similarity_threshold
                sims, rows = self._search(qv, top_k, float(threshold))
                return [self._collect(s, r) for s, r in zip(sims, rows)]
        except Exception as e:  # noqa: BLE001
            logger.error("query_similar_batch failed: %s", e)
            return [[] for _ in range(len(query_vectors))]

    def _collect(self, sims_row, rows_row):
        """Access statistics of one query's hits (src/ref_bank.py:205-219): count, LRU order - for every
        update strategy, as the reference does - and `total_queries`, which the reference only bumps when
        something passed the threshold (it returns early at :198-199)."""
        out = []
        pos = None
        for s, r in zip(sims_row, rows_row):
            if r < 0:
                break
            item = self._row_items[int(r)]
            item.access_count += 1
            if pos is None:
                pos = self._positions()
            idx = pos[id(item)]
            self._access.pop(idx, None)
            self._access[idx] = None
            out.append((item, float(s)))
        if out:
            self.stats["total_queries"] += 1
        return out

    def _positions(self) -> Dict[int, int]:
        """id(item) -> position in `references` (cached until the list changes length or order)."""
        if self._pos_cache is None or len(self._pos_cache) != len(self.references):
            self._pos_cache = {id(it): i for i, it in enumerate(self.references)}
        return self._pos_cache

    def query_by_cluster(self, cluster_id: int, top_k: int = 10) -> List[ReferenceItem]:
        try:
            with self._lock:
                if cluster_id not in self.clusters:
                    return []
                return [self.references[i] for i in self.clusters[cluster_id][:top_k]]
        except Exception as e:  # noqa: BLE001
            logger.error("query_by_cluster failed: %s", e)
            return []

    def get_cluster_centers(self) -> Optional[np.ndarray]:
        with self._lock:
            return None if self.cluster_centers is None else self.cluster_centers.copy()

    def perform_clustering(self, force: bool = False) -> bool:
        try:
            with self._lock:
                done = self._perform_clustering(force)
                if done and self.config.persistence_enabled:
                    self._save_to_disk()  # cluster ids changed on many items: fold everything
                return done
        except Exception as e:  # noqa: BLE001
            logger.error("clustering failed: %s", e)
            return False

    def _perform_clustering(self, force: bool = False) -> bool:
        """src/ref_bank.py:276-339 (host-side sklearn, as in the reference)."""
        if len(self.references) < 2 or (self.config.clustering_method == "none" and not force):
            return False
        try:
            vectors = np.array([r.vector for r in self.references])
            if self.config.clustering_method == "kmeans":
                n_clusters = min(self.config.num_clusters, len(self.references))
                fit = self._kmeans_device_assign(vectors, n_clusters) if self.gpu_kmeans else None
                if fit is None:                      # degenerate input (empty cluster, duplicates): sklearn end to end
                    from sklearn.cluster import KMeans
                    km = KMeans(n_clusters=n_clusters, random_state=42, n_init=10)
                    labels = km.fit_predict(vectors)
                    self.cluster_centers = km.cluster_centers_
                else:
                    labels, self.cluster_centers = fit
            elif self.config.clustering_method == "dbscan":
                from sklearn.cluster import DBSCAN
                labels = DBSCAN(eps=0.5, min_samples=5).fit_predict(vectors)
                centers = [vectors[labels == lb].mean(axis=0) for lb in np.unique(labels) if lb != -1]
                self.cluster_centers = np.array(centers) if centers else None
            else:
                return False
            self.clusters.clear()
            for i, lb in enumerate(labels):
                if lb != -1:
                    self.references[i].cluster_id = int(lb)
                    self.clusters.setdefault(lb, []).append(i)
                else:
                    self.references[i].cluster_id = None
            self.stats["clustering_count"] += 1
            self.stats["last_clustering_time"] = time.time()
            return True
        except Exception as e:  # noqa: BLE001
            logger.error("clustering failed: %s", e)
            return False

    def _kmeans_device_assign(self, vectors: np.ndarray, n_clusters: int, n_init: int = 10, max_iter: int = 300,
                              tol: float = 1e-4, seed: int = 42):
        """KMeans(n_clusters, random_state=42, n_init=10).fit_predict(vectors) (src/ref_bank.py:296-300) with the
        ASSIGNMENT step - the O(B * C * d) part - as kernel (a): argmin_c |x - c|^2 = argmax_c (x.c - |c|^2 / 2) is
        an inner-product top-1 of the augmented rows [x, 1] against the augmented centres [c, -|c|^2 / 2]
        (SURVEY.md §8f rank 3).  Initialisation (k-means++) and the update step stay scikit-learn's / NumPy's,
        and the driver follows scikit-learn's Lloyd loop (sklearn/cluster/_kmeans.py, KMeans.fit and
        _kmeans_single_lloyd: centred data, one RandomState shared by the n_init restarts, stop on unchanged
        labels or centre shift <= tol * mean variance, a last assignment when stopped by tolerance, best inertia
        wins) so that the labels are scikit-learn's.  Rows whose two best scores are closer than fp32 can
        resolve are re-decided in fp64 on the host.  Returns (labels, centres) or None when a cluster runs
        empty (scikit-learn relocates centres there: the caller falls back to it)."""
        from sklearn.cluster import kmeans_plusplus
        from sklearn.utils import check_random_state
        X = np.ascontiguousarray(vectors, dtype=vectors.dtype if vectors.dtype in (np.float32, np.float64) else np.float64)
        X = X.copy()
        n, d = X.shape
        if n_clusters < 2 or n < n_clusters:
            return None
        tol_abs = float(np.mean(np.var(X, axis=0)) * tol)
        rs = check_random_state(seed)
        mean = X.mean(axis=0)
        X -= mean
        x_sq = np.einsum("ij,ij->i", X, X)
        q = np.empty((n, d + 1), np.float32)
        q[:, :d], q[:, d] = X, 1.0
        try:
            import torch
            q_dev = torch.from_numpy(q).cuda()
        except Exception:  # noqa: BLE001 - no torch: host rows, staged by the library on every call
            q_dev = None
        kk = 2 if n_clusters <= 16 else (11 if n_clusters <= 32 else 27)   # candidate width 16 / 32 / 64 >= centres
        kk = min(kk, n_clusters)

        def assign(centers):
            c = np.empty((n_clusters, d + 1), np.float32)
            c[:, :d] = centers
            c[:, d] = -0.5 * np.einsum("ij,ij->i", centers, centers)
            gal = Gallery(c, normalize=False)
            try:
                sims, idx = gal.search(q_dev if q_dev is not None else q, kk)
            finally:
                gal.close()
            if q_dev is not None:
                sims, idx = sims.cpu().numpy(), idx.cpu().numpy()
            labels = idx[:, 0].astype(np.int32)
            # fp32 cannot separate the two best centres of these rows: decide them as scikit-learn does (argmin of
            # the squared distance in the data's own precision, first minimum wins)
            scale = max(1.0, float(np.abs(sims[:, 0]).max()))
            close = np.nonzero((sims[:, 0] - sims[:, 1]) <= 1e-4 * scale)[0]
            for r0 in range(0, len(close), 4096):
                rows = close[r0:r0 + 4096]
                dist = x_sq[rows, None] - 2.0 * (X[rows] @ centers.T) + np.einsum("ij,ij->i", centers, centers)[None, :]
                labels[rows] = np.argmin(dist, axis=1)
            return labels

        def same_clustering(a, b):
            mapping = np.full(n_clusters, -1, np.int32)
            for i in range(n):
                if mapping[a[i]] == -1:
                    mapping[a[i]] = b[i]
                elif mapping[a[i]] != b[i]:
                    return False
            return True

        best = None
        for _ in range(n_init):
            centers, _ = kmeans_plusplus(X, n_clusters, x_squared_norms=x_sq, random_state=rs)
            centers = centers.astype(X.dtype, copy=True)
            labels_old = np.full(n, -1, np.int32)
            strict = False
            for _it in range(max_iter):
                labels = assign(centers)
                counts = np.bincount(labels, minlength=n_clusters)
                if (counts == 0).any():
                    return None
                new = np.zeros_like(centers)
                np.add.at(new, labels, X)
                new /= counts[:, None]
                shift = float(((new - centers) ** 2).sum())
                centers = new
                if np.array_equal(labels, labels_old):
                    strict = True
                    break
                if shift <= tol_abs:
                    break
                labels_old = labels
            if not strict:
                labels = assign(centers)
            inertia = float(((X - centers[labels]) ** 2).sum())
            if best is None or (inertia < best[0] and not same_clustering(labels, best[1])):
                best = (inertia, labels, centers)
        return best[1], best[2] + mean

    # ------------------------------------------------------------------ insert / evict
    def _is_too_similar(self, vector: np.ndarray) -> bool:
        """Any stored vector with cosine > similarity_threshold (src/ref_bank.py:341-363), as one exact
        top-1 search instead of a random 100-sample."""
        if not self.references:
            return False
        v = np.asarray(vector, dtype=np.float32).reshape(1, -1)
        if not np.any(v):
            return False  # zero vector: cosine defined as 0.0 (:497-501)
        sims, rows = self._search(v, 1, -np.inf)
        return bool(rows[0, 0] >= 0 and sims[0, 0] > self.config.similarity_threshold)

    def _remove_at(self, idx: int):
        item = self.references.pop(idx)
        self._dev_remove(item)
        self._update_clusters_after_removal(idx)
        if self.config.persistence_enabled:
            self._journal({"op": "remove", "index": int(idx)})

    def _remove_reference(self):
        """src/ref_bank.py:365-399."""
        if not self.references:
            return
        s = self.config.update_strategy
        if s == "fifo":
            self._remove_at(0)
        elif s == "lru":
            if self._access:
                oldest, _ = self._access.popitem(last=False)
                if oldest < len(self.references):
                    self._remove_at(oldest)
            else:
                self._remove_at(0)
        elif s == "random":
            self._remove_at(int(np.random.randint(len(self.references))))
        elif s == "similarity":
            self._remove_most_similar()
        self.stats["total_removed"] += 1

    def _remove_most_similar(self):
        """src/ref_bank.py:401-427: drop one member of the most similar pair.  The O(B^2) Python loop
        becomes one self-search (top-1 excluding self) on the GPU."""
        if len(self.references) < 2:
            return
        n = len(self._row_items)
        rows = self._gallery.get_rows(np.arange(n, dtype=np.int64))
        sims, nbr = self._gallery.search(rows, 1, skip_self=True)
        r = int(np.argmax(sims[:, 0]))
        a, b = self._row_items[r], self._row_items[int(nbr[r, 0])]
        pos = {id(it): i for i, it in enumerate(self.references)}
        i, j = sorted((pos[id(a)], pos[id(b)]))
        victim = i if self.references[i].access_count <= self.references[j].access_count else j
        self._remove_at(victim)

    def _update_clusters_after_removal(self, removed_idx: int):
        """src/ref_bank.py:429-460."""
        new_clusters = {}
        for cid, members in self.clusters.items():
            kept = [m if m < removed_idx else m - 1 for m in members if m != removed_idx]
            if kept:
                new_clusters[cid] = kept
        self.clusters = new_clusters
        self._access = OrderedDict((m if m < removed_idx else m - 1, None) for m in self._access if m != removed_idx)
        self._pos_cache = None

    def _compute_similarities(self, query_vector: np.ndarray) -> np.ndarray:
        """src/ref_bank.py:462-484: cosine of the query against every reference, in `references` order."""
        if not self.references:
            return np.array([])
        n = len(self._row_items)
        sims = self._gallery.similarity_matrix(np.asarray(query_vector, np.float32).reshape(1, -1),
                                               normalize_queries=True)[0]
        order = np.array([self._row_of[id(it)] for it in self.references], dtype=np.int64)
        return sims[:n][order]

    def _cosine_similarity(self, vec1: np.ndarray, vec2: np.ndarray) -> float:
        n1, n2 = np.linalg.norm(vec1), np.linalg.norm(vec2)
        if n1 == 0 or n2 == 0:
            return 0.0
        return float(np.dot(vec1, vec2) / (n1 * n2))

    # ------------------------------------------------------------------ persistence (:505-576)
    _JOURNAL = "references.journal.jsonl"

    def _journal(self, op: Dict[str, Any]):
        """Append one operation; fold the journal into the JSON files every journal_compact_every ops."""
        try:
            self._seq += 1
            with open(Path(self.config.save_path) / self._JOURNAL, "a") as f:
                f.write(json.dumps(dict(op, seq=self._seq)) + "\n")
            self._journal_ops += 1
            if self._journal_ops >= self.journal_compact_every:
                self._save_to_disk()
        except Exception as e:  # noqa: BLE001
            logger.error("journal write failed: %s", e)

    def flush(self):
        """Fold the journal into references.json / clusters.json / stats.json / config.json (the layout the
        reference reads back, src/ref_bank.py:537-576)."""
        with self._lock:
            if self.config.persistence_enabled:
                self._save_to_disk()

    def _save_to_disk(self):
        try:
            root = Path(self.config.save_path)
            root.mkdir(parents=True, exist_ok=True)

            def put(name, text):                       # a reader (or a crash) never sees half a file
                tmp = root / (name + ".tmp")
                tmp.write_text(text)
                os.replace(tmp, root / name)

            # The folded sequence number travels INSIDE references.json (the one file whose replacement commits
            # the fold): as an extra key of the first item - ReferenceItem.from_dict, here and in the reference
            # (src/ref_bank.py:73-83), reads named keys only - or, for an empty bank, in the {"references": []}
            # wrapper the shipped cache/ref_bank snapshot uses.  A crash between the replace and the unlink of
            # the journal then replays nothing twice: lines with seq <= the folded number are skipped on load.
            data = [r.to_dict() for r in self.references]
            if data:
                data[0] = dict(data[0], _journal_seq=self._seq)
            put("clusters.json", json.dumps({
                "clusters": {str(k): v for k, v in self.clusters.items()},
                "cluster_centers": self.cluster_centers.tolist() if self.cluster_centers is not None else None},
                indent=2))
            put("stats.json", json.dumps(self.stats, indent=2))
            put("config.json", json.dumps(asdict(self.config), indent=2))
            put("references.json", json.dumps(data if data else {"references": [], "_journal_seq": self._seq}, indent=2))
            j = root / self._JOURNAL
            if j.exists():
                j.unlink()
            self._journal_ops = 0
        except Exception as e:  # noqa: BLE001
            logger.error("save to disk failed: %s", e)

    def _load_from_disk(self):
        try:
            root = Path(self.config.save_path)
            refs_file = root / "references.json"
            if not self.config.persistence_enabled or not (refs_file.exists() or (root / self._JOURNAL).exists()):
                return
            data = json.loads(refs_file.read_text()) if refs_file.exists() else []
            folded = 0
            if isinstance(data, dict):  # the snapshot shipped in cache/ref_bank wraps the list
                folded = int(data.get("_journal_seq", 0))
                data = data.get("references", [])
            elif data and isinstance(data[0], dict):
                folded = int(data[0].get("_journal_seq", 0))
            self._seq = folded
            self.references = [ReferenceItem.from_dict(d) for d in data]
            self._pos_cache = None
            cf = root / "clusters.json"
            if cf.exists():
                cd = json.loads(cf.read_text())
                self.clusters = {int(k): v for k, v in cd.get("clusters", {}).items()}
                if cd.get("cluster_centers"):
                    self.cluster_centers = np.array(cd["cluster_centers"])
            sf = root / "stats.json"
            if sf.exists():
                self.stats.update(json.loads(sf.read_text()))
            jf = root / self._JOURNAL
            if jf.exists():                                 # operations made after the last fold
                for line in jf.read_text().splitlines():
                    if not line.strip():
                        continue
                    op = json.loads(line)
                    if "seq" in op:
                        if int(op["seq"]) <= folded:        # already inside references.json (crash after the fold)
                            continue
                        self._seq = max(self._seq, int(op["seq"]))
                    if op.get("op") == "add":
                        self.references.append(ReferenceItem.from_dict(op["item"]))
                        self.stats["total_added"] += 1
                    elif op.get("op") == "remove" and 0 <= op["index"] < len(self.references):
                        self.references.pop(op["index"])
                        self._update_clusters_after_removal(op["index"])
                        self.stats["total_removed"] += 1
                    self._journal_ops += 1
            self._dev_rebuild()
        except Exception as e:  # noqa: BLE001
            logger.error("load from disk failed: %s", e)

    def clear(self):
        with self._lock:
            self.references.clear()
            self.clusters.clear()
            self.cluster_centers = None
            self._access.clear()
            self._pos_cache = None
            self.stats = self._fresh_stats()
            self._dev_rebuild()
            if self.config.persistence_enabled:
                self._save_to_disk()      # the journal on disk describes the bank that was just dropped

    def get_statistics(self) -> Dict[str, Any]:
        with self._lock:
            st = dict(self.stats)
            st.update({"current_size": len(self.references), "num_clusters": len(self.clusters),
                       "max_size": self.config.max_size, "clustering_method": self.config.clustering_method,
                       "update_strategy": self.config.update_strategy})
            if self.references:
                ac = [r.access_count for r in self.references]
                st.update({"avg_access_count": np.mean(ac), "max_access_count": np.max(ac),
                           "min_access_count": np.min(ac)})
            return st

    def export_references(self, export_path: str, format: str = "json") -> bool:
        try:
            with self._lock:
                p = Path(export_path)
                p.parent.mkdir(parents=True, exist_ok=True)
                if format == "json":
                    p.write_text(json.dumps([r.to_dict() for r in self.references], indent=2))
                elif format == "numpy":
                    np.savez(p, vectors=np.array([r.vector for r in self.references]),
                             metadata=np.array([r.metadata for r in self.references], dtype=object))
                elif format == "pickle":
                    with open(p, "wb") as f:
                        pickle.dump(self.references, f)
                else:
                    raise ValueError(f"unsupported export format: {format}")
                return True
        except Exception as e:  # noqa: BLE001
            logger.error("export failed: %s", e)
            return False

    def import_references(self, import_path: str, format: str = "json") -> bool:
        try:
            p = Path(import_path)
            if not p.exists():
                raise FileNotFoundError(str(p))
            with self._lock:
                if format == "json":
                    items = [ReferenceItem.from_dict(d) for d in json.loads(p.read_text())]
                elif format == "numpy":
                    z = np.load(p, allow_pickle=True)
                    items = [ReferenceItem(vector=v, metadata=m, timestamp=time.time())
                             for v, m in zip(z["vectors"], z["metadata"])]
                elif format == "pickle":
                    with open(p, "rb") as f:
                        items = pickle.load(f)
                else:
                    raise ValueError(f"unsupported import format: {format}")
                room = max(0, self.config.max_size - len(self.references))
                taken = items[:room]
                self.references.extend(taken)
                self._pos_cache = None
                self._dev_append(taken)
                self.stats["total_added"] += len(taken)
                if self.config.persistence_enabled:
                    self._save_to_disk()  # later journal lines index into a list that now holds these rows
                return True
        except Exception as e:  # noqa: BLE001
            logger.error("import failed: %s", e)
            return False


def create_reference_bank(max_size: int = 10000, similarity_threshold: float = 0.9,
                          clustering_method: str = "kmeans", **kwargs) -> ReferenceBank:
    return ReferenceBank(ReferenceBankConfig(max_size=max_size, similarity_threshold=similarity_threshold,
                                             clustering_method=clustering_method, **kwargs))


__all__ = ["ReferenceBankConfig", "ReferenceItem", "ReferenceBank", "create_reference_bank"]
