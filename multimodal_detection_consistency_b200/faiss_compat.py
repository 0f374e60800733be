"""FAISS-shaped front for libtvc galleries.

The reference reaches its top-k arithmetic through exactly this surface
(`faiss.IndexFlatIP(d)`, `.add`, `.search -> (D, I)`, `.ntotal`, `.is_trained`, `write_index`,
`read_index`, `get_num_gpus`, `index_cpu_to_gpu`: src/retrieval.py:100-112,136,225-226,253,495-518,
652-655,781-783; experiments/defenses/retrieval_ref.py:140-156,252; scripts/build_faiss_indices.py:
138-158,244).  Installing this module as `sys.modules["faiss"]` routes the reference's own files to
the B200 kernels with no source change.  Only exact inner-product search is
provided: IVF / HNSW / PQ are approximate and north_star demands exact top-k, so those constructors
return the same exact index.
"""
from __future__ import annotations

import pickle
from typing import Optional

import numpy as np

from ._native import Gallery

METRIC_INNER_PRODUCT = 0


class IndexFlatIP:
    """Exact inner-product index resident in HBM (bf16 GEMM operand + fp32 master for re-rank)."""

    def __init__(self, d: int, *_, **__):
        self.d = int(d)
        self.is_trained = True
        self._gallery: Optional[Gallery] = None

    @property
    def ntotal(self) -> int:
        return 0 if self._gallery is None else len(self._gallery)

    def train(self, x) -> None:  # flat indices need no training (src/retrieval.py:513-515)
        self.is_trained = True

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected rows of dimension {self.d}, got {x.shape}")
        if self._gallery is None:
            self._gallery = Gallery(x)
        else:
            self._gallery.append(x)

    def search(self, x, k: int):
        """-> (D [n, k] float32 descending, I [n, k] int64), -1 / -inf padded like FAISS."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x[None]
        if self._gallery is None:
            return (np.full((x.shape[0], k), -np.inf, np.float32), np.full((x.shape[0], k), -1, np.int64))
        return self._gallery.search(x, int(k))

    def reconstruct_n(self, i0: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else n
        if self._gallery is None or n <= 0:
            return np.zeros((0, self.d), np.float32)
        return self._gallery.get_rows(np.arange(i0, i0 + n, dtype=np.int64))

    def reset(self) -> None:
        if self._gallery is not None:
            self._gallery.close()
        self._gallery = None


# approximate index families collapse to the exact one (see module docstring)
def IndexIVFFlat(quantizer, d, nlist=100, *_, **__):
    return IndexFlatIP(d)


def IndexHNSWFlat(d, m=32, *_, **__):
    return IndexFlatIP(d)


class StandardGpuResources:
    pass


def get_num_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def index_cpu_to_gpu(res, device, index):
    return index  # already resident on the GPU


def index_gpu_to_cpu(index):
    return index


def write_index(index: IndexFlatIP, path: str) -> None:
    """Sidecar written next to the retriever pickle (src/retrieval.py:781-783): the fp32 rows."""
    with open(path, "wb") as f:
        pickle.dump({"format": "tvc-flat-ip", "d": index.d, "rows": index.reconstruct_n()}, f)


def read_index(path: str) -> IndexFlatIP:
    with open(path, "rb") as f:
        blob = pickle.load(f)
    if not isinstance(blob, dict) or blob.get("format") != "tvc-flat-ip":
        raise ValueError(f"{path} is not a tvc flat-IP sidecar")
    idx = IndexFlatIP(int(blob["d"]))
    if len(blob["rows"]):
        idx.add(blob["rows"])
    return idx
