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
    """Flat inner-product index resident in HBM (bf16 GEMM operand + fp32 master for re-rank).  Exhaustive like
    faiss.IndexFlatIP (every row is scored); the k results are the fp32-re-scored best of the 16-64 best rows by
    bf16 score, so a row can differ from FAISS's only inside the ~1e-3 band around the k-th similarity
    (Gallery.search documents the bound); any k is accepted (k > 56 takes a chunked similarity matrix + top-k)."""

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


# ---- on-disk format ---------------------------------------------------------------------------
# FAISS's native flat-index layout (faiss/impl/index_write.cpp, write_index / write_index_header,
# v1.7.x - the version the reference pins, README.md:546), little endian:
#   fourcc "IxFI" (inner product) | "IxF2" (L2)          uint32
#   d                                                     int32
#   ntotal                                                int64
#   two reserved idx_t (1 << 20)                          2 x int64
#   is_trained                                            uint8
#   metric_type (0 = inner product, 1 = L2)               int32
#   [metric_arg float32 when metric_type > 1]
#   number of float32 values (ntotal * d)                 uint64
#   the rows                                              float32 [ntotal, d]
# Files written here are meant to be readable by a real faiss.read_index and vice versa; faiss is not
# installable in the build container, so the layout is restated from the published source and covered
# only by our own round trip (tests/test_formats.py).
_FOURCC_IP, _FOURCC_L2 = b"IxFI", b"IxF2"


def _pack_flat(rows: np.ndarray, d: int, metric: int = METRIC_INNER_PRODUCT) -> bytes:
    import struct
    rows = np.ascontiguousarray(rows, dtype="<f4").reshape(-1, d) if rows.size else np.zeros((0, d), "<f4")
    head = (_FOURCC_IP if metric == METRIC_INNER_PRODUCT else _FOURCC_L2)
    head += struct.pack("<iqqqBi", d, rows.shape[0], 1 << 20, 1 << 20, 1, metric)
    return head + struct.pack("<Q", rows.size) + rows.tobytes()


def _unpack_flat(blob: bytes):
    import struct
    if blob[:4] not in (_FOURCC_IP, _FOURCC_L2):
        raise ValueError("not a FAISS flat index (fourcc %r): only IndexFlatIP / IndexFlatL2 files are exact" % blob[:4])
    d, ntotal, _, _, _, metric = struct.unpack_from("<iqqqBi", blob, 4)
    off = 4 + struct.calcsize("<iqqqBi")
    if metric > 1:
        off += 4
    (count,) = struct.unpack_from("<Q", blob, off)
    off += 8
    if count != ntotal * d:
        raise ValueError(f"corrupt flat index: {count} values for {ntotal} x {d}")
    rows = np.frombuffer(blob, dtype="<f4", count=count, offset=off).reshape(ntotal, d)
    return d, metric, rows


def write_index(index: IndexFlatIP, path: str) -> None:
    """Sidecar written next to the retriever pickle (src/retrieval.py:781-783) in FAISS's own layout."""
    with open(path, "wb") as f:
        f.write(_pack_flat(index.reconstruct_n(), index.d))


def read_index(path: str) -> IndexFlatIP:
    """FAISS flat-index files (see above); also the pickle sidecar earlier versions of this module wrote."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:4] in (_FOURCC_IP, _FOURCC_L2):
        d, _, rows = _unpack_flat(blob)
    else:
        old = pickle.loads(blob)
        if not isinstance(old, dict) or old.get("format") != "tvc-flat-ip":
            raise ValueError(f"{path} is neither a FAISS flat index nor a tvc sidecar")
        d, rows = int(old["d"]), old["rows"] if old.get("rows") is not None else np.zeros((0, int(old["d"])), np.float32)
    idx = IndexFlatIP(int(d))
    if len(rows):
        idx.add(np.ascontiguousarray(rows, dtype=np.float32))
    return idx
