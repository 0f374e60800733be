"""CPU-side checks: the C-ABI library loads and exports every symbol include/tvc.h declares, the
host mirrors keep the reference's configuration surface, and the product path fails loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes
import dataclasses
import re
from pathlib import Path

import numpy as np
import pytest

import multimodal_detection_consistency_b200 as tvc
from multimodal_detection_consistency_b200 import _native as N

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "tvc.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tvc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = tvc.load_library()
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tvc.h but not exported by libtvc.so"
    assert set(names) == set(N.EXPORTED_SYMBOLS), set(names) ^ set(N.EXPORTED_SYMBOLS)
    assert lib.tvc_version() == 100


def test_header_constants_match_binding():
    text = (ROOT / "include" / "tvc.h").read_text()
    for name, val in [("TVC_MAX_K", N.MAX_K), ("TVC_MAX_VARIANTS", N.MAX_VARIANTS), ("TVC_MAX_REFS", N.MAX_REFS),
                      ("TVC_NSCORES", N.NSCORES)]:
        assert int(re.search(rf"#define {name} (\d+)", text).group(1)) == val
    enum = re.findall(r"TVC_S_[A-Z_]+ = (\d+)", text)
    assert [int(x) for x in enum] == list(range(N.NSCORES))
    assert len(N.SCORE_NAMES) == N.NSCORES
    assert ctypes.sizeof(N.DetectorParams) == 72


def test_default_params_are_the_reference_defaults():
    p = tvc.default_params().as_dict()
    assert (p["n_variants"], p["n_retrieval"], p["n_generative"]) == (5, 10, 3)
    assert p["methods"] == 7 and p["aggregation"] == 0 and p["voting"] == 1
    assert np.allclose([p["w_text_variants"], p["w_sd_reference"], p["w_consistency"]], [0.4, 0.4, 0.2])
    assert p["detection_threshold"] == 0.5 and p["cc_base_threshold"] == 0.5 and p["cc_adaptive"] == 1
    assert np.allclose(p["cc_weights"], 0.25) and abs(p["dedup_threshold"] - 0.95) < 1e-6


def test_config_dataclasses_keep_reference_fields():
    from multimodal_detection_consistency_b200 import (DetectionConfig, DetectorConfig, ReferenceBankConfig,
                                                       RetrievalConfig)
    fields = lambda c: [f.name for f in dataclasses.fields(c)]  # noqa: E731
    assert fields(RetrievalConfig) == ["clip_model", "device", "batch_size", "top_k", "similarity_metric",
                                       "index_type", "faiss_index_type", "n_clusters", "enable_cache", "cache_dir",
                                       "normalize_features", "use_gpu_index"]
    assert fields(ReferenceBankConfig) == ["max_size", "similarity_threshold", "clustering_method", "num_clusters",
                                           "update_strategy", "persistence_enabled", "save_path", "auto_clustering",
                                           "clustering_interval", "feature_dim"]
    assert fields(DetectorConfig) == ["clip_model", "device", "detection_methods", "use_text_variants",
                                      "num_text_variants", "text_similarity_threshold", "use_sd_reference",
                                      "num_reference_images", "reference_similarity_threshold",
                                      "consistency_threshold", "consistency_weight", "detection_threshold",
                                      "adaptive_threshold", "threshold_percentile", "score_aggregation",
                                      "enable_cache", "cache_size", "batch_size"]
    assert fields(DetectionConfig)[:4] == ["use_text_variants", "text_variant_count", "use_retrieval_ref",
                                           "retrieval_top_k"]
    assert DetectorConfig().detection_methods == ["text_variants", "sd_reference", "consistency"]
    assert RetrievalConfig().top_k == 10 and ReferenceBankConfig().similarity_threshold == 0.9
    with pytest.raises(ValueError):
        ReferenceBankConfig(update_strategy="bogus")


def _no_cuda():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_cuda(), reason="only meaningful without a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    lib = tvc.load_library()
    h = ctypes.c_void_p()
    assert lib.tvc_ctx_create(0, ctypes.byref(h)) == N.TVC_ERR_NO_DEVICE
    with pytest.raises(tvc.TvcError):
        tvc.Gallery(np.zeros((4, 8), np.float32))
    with pytest.raises(tvc.TvcError):
        tvc.Context(0)
    # the reference-shaped methods keep the reference's never-raise convention instead
    from multimodal_detection_consistency_b200 import MultiModalRetriever, RetrievalConfig
    r = MultiModalRetriever(RetrievalConfig(), clip_model=object())
    assert r.retrieve_images_by_text("a cat") == ([], [])
    assert r.get_stats()["image_count"] == 0


def test_product_code_does_not_import_the_oracle():
    pkg = ROOT / "multimodal_detection_consistency_b200"
    for f in pkg.glob("*.py"):
        src = f.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), f.name


def test_header_is_plain_c_and_binds_with_dlopen(tmp_path):
    """include/tvc.h compiles as C11 with -Wall -Werror, and a C program bound with dlopen alone can call
    the device-free entry points (defaults, widths, status strings, argument checks, tvc_ctx_create)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    tvc.load_library()
    exe = tmp_path / "abi_probe"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "c" / "abi_probe.c"), "-o", str(exe), "-ldl", "-lm"], check=True)
    out = subprocess.run([str(exe), str(N.LIB_PATH)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("abi ok")


def test_host_batch_piece_bounds():
    """TVCScorer._piece_bounds: weighted / equal splits cover the batch exactly, fall back to two halves and
    then to one piece when a piece would be smaller than min_chunk_queries."""
    import types
    from multimodal_detection_consistency_b200.pipeline import TVCScorer
    sc = types.SimpleNamespace(host_chunks=(1, 6, 1), host_chunks_pageable=2, min_chunk_queries=1024, world=1)
    pb = lambda q: TVCScorer._piece_bounds(sc, q)  # noqa: E731
    assert pb(16384) == [(0, 2048), (2048, 14336), (14336, 16384)]
    assert pb(8192) == [(0, 1024), (1024, 7168), (7168, 8192)]
    assert pb(8000) == [(0, 4000), (4000, 8000)]            # 1/8 piece too small -> two halves
    assert pb(2047) == [(0, 2047)]                          # halves too small -> no pipelining
    sc.host_chunks = 4
    assert pb(16384) == [(i * 4096, (i + 1) * 4096) for i in range(4)]
    assert pb(10001)[0][0] == 0 and pb(10001)[-1][1] == 10001 and all(a[1] == b[0] for a, b in zip(pb(10001), pb(10001)[1:]))
    sc.host_chunks = 1
    assert pb(16384) == [(0, 16384)]
    # a rank's slice on several GPUs: a quarter-size first piece where the (1, 6, 1) split would fall under a wave
    sc.host_chunks, sc.world = (1, 6, 1), 8
    assert pb(2048) == [(0, 512), (512, 2048)]
    assert pb(1024) == [(0, 1024)]
    assert TVCScorer._piece_bounds(sc, 4096, pinned=False) == [(0, 2048), (2048, 4096)]
