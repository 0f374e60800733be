"""bench.py's drop-in block (latency / dropin: the reference's API with table encoders) runs end to end on the CPU
with the native layer replaced by the oracle-backed test double - the block's own host code (table encoder, timing
loops, worker threads) is what is checked here; the rates it prints mean nothing on the double."""
import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_native  # noqa: E402

import bench  # noqa: E402


def test_dropin_block_runs_on_the_double():
    rng = np.random.default_rng(0)

    def unit(*shape):
        x = rng.standard_normal(shape).astype(np.float32)
        return x / np.linalg.norm(x, axis=-1, keepdims=True)

    args = types.SimpleNamespace(variants=5, topk=10, dim=32)
    n, d = 48, 32
    g_host, b_host = torch.from_numpy(unit(400, d)), torch.from_numpy(unit(64, d))
    img, txt, var = unit(n, d), unit(n, d), unit(n, 5, d)
    with fake_native.installed():
        latency, dropin = bench.dropin_blocks(np, args, g_host, b_host, img, txt, var, n_items=n)
    assert latency["single_query_ms"]["calls"] == n - 8 and latency["five_variant_batch_ms"]["calls"] == n - 8
    assert dropin["queries"] == n and dropin["threads4_queries"] == n
    for key in ("batch_api_queries_per_s", "batch_retrieve_queries_per_s", "batch_detect_queries_per_s",
                "threads4_micro_batched_queries_per_s", "sequential_single_calls_queries_per_s"):
        assert dropin[key] > 0


def test_table_encoder_gathers_rows():
    t = np.arange(12, dtype=np.float32).reshape(6, 2)
    i = np.arange(8, dtype=np.float32).reshape(4, 2) + 100
    x = np.arange(6, dtype=np.float32).reshape(3, 2) + 1000
    enc = bench._TableEncoder({"a": 0, "b": 5}, t, i, x, extra_base=1_000_000_000)
    assert np.array_equal(enc.encode_text(["b", "a", "b"]), t[[5, 0, 5]])
    assert np.array_equal(enc.encode_text("a"), t[[0]])
    assert np.array_equal(enc.encode_image([3, 1_000_000_002, 0]), np.stack([i[3], x[2], i[0]]))
    assert np.array_equal(enc.encode_image(2), i[[2]])
