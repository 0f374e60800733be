"""On-disk formats: the FAISS flat-index sidecar layout, retriever pickles written
by the reference's own class names, and the ReferenceBank 4-file JSON layout incl. the snapshot the
reference ships (cache/ref_bank)."""
import pickle
import struct
import sys
import types

import numpy as np
import pytest

from multimodal_detection_consistency_b200 import faiss_compat
from multimodal_detection_consistency_b200.retrieval import MultiModalRetriever, RetrievalConfig


def test_faiss_flat_layout_round_trip_and_header():
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((37, 24)).astype(np.float32)
    blob = faiss_compat._pack_flat(rows, 24)
    # faiss/impl/index_write.cpp: fourcc, d, ntotal, 2 reserved idx_t, is_trained, metric, count, data
    assert blob[:4] == b"IxFI"
    d, n, r0, r1, trained, metric = struct.unpack_from("<iqqqBi", blob, 4)
    assert (d, n, r0, r1, trained, metric) == (24, 37, 1 << 20, 1 << 20, 1, 0)
    assert struct.unpack_from("<Q", blob, 4 + 33)[0] == 37 * 24
    assert len(blob) == 4 + 33 + 8 + 37 * 24 * 4
    d2, metric2, back = faiss_compat._unpack_flat(blob)
    assert d2 == 24 and metric2 == 0 and np.array_equal(back, rows)
    with pytest.raises(ValueError):
        faiss_compat._unpack_flat(b"IxHN" + blob[4:])              # an HNSW file is not an exact index
    with pytest.raises(ValueError):
        faiss_compat._unpack_flat(blob[:37] + struct.pack("<Q", 5) + blob[45:])   # wrong value count


def test_retriever_pickle_written_under_the_reference_class_path(tmp_path, monkeypatch):
    """A pickle that names src.retrieval.RetrievalConfig (what the reference writes, src/retrieval.py:773-778)
    loads even when `src` is not importable; the .faiss sidecar alone can carry the features."""
    feats = np.random.default_rng(1).standard_normal((9, 16)).astype(np.float32)
    mod = types.ModuleType("src.retrieval")
    pkg = types.ModuleType("src")

    class RefConfig:                       # stands in for the reference's dataclass while pickling
        def __init__(self):
            self.top_k = 7
    RefConfig.__name__ = RefConfig.__qualname__ = "RetrievalConfig"
    RefConfig.__module__ = "src.retrieval"
    mod.RetrievalConfig = RefConfig
    monkeypatch.setitem(sys.modules, "src", pkg)
    monkeypatch.setitem(sys.modules, "src.retrieval", mod)
    path = tmp_path / "img_index.pkl"
    with open(path, "wb") as f:
        pickle.dump({"image_features": None, "image_paths": [f"p{i}" for i in range(9)], "config": RefConfig()}, f)
    with open(path.with_suffix(".faiss"), "wb") as f:
        f.write(faiss_compat._pack_flat(feats, 16))
    monkeypatch.delitem(sys.modules, "src.retrieval")
    monkeypatch.delitem(sys.modules, "src")
    data, got = MultiModalRetriever(RetrievalConfig(), clip_model=object())._load(path, "image_features")
    assert isinstance(data["config"], RetrievalConfig) and data["image_paths"][3] == "p3"
    assert np.array_equal(got, feats)
