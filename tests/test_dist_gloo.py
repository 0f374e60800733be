"""World-size-2/3 gloo runs of TVCScorer's sharded path on CPU (oracle engine): the per-slice results
must equal the unsharded single-process run — indices and histogram bit-exact."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(seed=0, n=700, b=90, q=37, d=32, v=5):
    from oracle import tvc_oracle as O
    g = O.synth_gallery(n, d, seed=seed, clusters=16, dup_rate=0.01)
    bank = O.synth_gallery(b, d, seed=seed + 1, clusters=16)
    img, txt, var = O.synth_queries(g, q, v, seed=seed + 2)
    return g, bank, img, txt, var


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_engine import OracleEngine
    from multimodal_detection_consistency_b200.pipeline import TVCScorer, shard_bounds
    g, bank, img, txt, var = _data()
    glo, ghi = shard_bounds(len(g), world, rank)
    blo, bhi = shard_bounds(len(bank), world, rank)
    sc = TVCScorer(g[glo:ghi], bank[blo:bhi], k=10, total_gallery_rows=len(g), total_bank_rows=len(bank),
                   engine=OracleEngine())
    out = sc.score_batch(torch.from_numpy(img), torch.from_numpy(txt), torch.from_numpy(var))
    out2 = sc.score_batch(img, txt, var)  # second batch accumulates the histogram
    lo, hi = out["slice"]
    np.savez(Path(out_dir) / f"r{rank}.npz", lo=lo, hi=hi, scores=out["scores"].numpy(), flags=out["flags"].numpy(),
             topk_idx=out["topk_idx"].numpy(), topk_sim=out["topk_sim"].numpy(), bank_idx=out["bank_idx"].numpy(),
             hub=sc.k_occurrence.numpy(), scores2=out2["scores"].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(tmp_path, world):
    from oracle_engine import OracleEngine
    from multimodal_detection_consistency_b200.pipeline import TVCScorer
    g, bank, img, txt, var = _data()
    ref = TVCScorer(g, bank, k=10, engine=OracleEngine())
    want = ref.score_batch(img, txt, var)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        covered += hi - lo
        assert np.array_equal(z["topk_idx"], want["topk_idx"].numpy()[lo:hi])
        assert np.array_equal(z["topk_sim"], want["topk_sim"].numpy()[lo:hi])
        assert np.array_equal(z["bank_idx"], want["bank_idx"].numpy()[lo:hi])
        assert np.abs(z["scores"] - want["scores"].numpy()[lo:hi]).max() <= 1e-6
        assert np.array_equal(z["flags"], want["flags"].numpy()[lo:hi])
        assert np.array_equal(z["scores"], z["scores2"])
        assert np.array_equal(z["hub"], 2 * ref.k_occurrence.numpy())   # all-reduced, two batches
    assert covered == len(img)


def test_shard_and_slice_bounds():
    from multimodal_detection_consistency_b200.pipeline import shard_bounds, slice_bounds
    for n in [0, 1, 7, 8, 9, 1000003]:
        for w in [1, 2, 3, 8]:
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert [slice_bounds(n, w, r) for r in range(w)] == spans
