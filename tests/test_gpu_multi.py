"""Multi-GPU runs (world = 2, 4, 8 - whatever the box has) of the sharded TVCScorer (candidates stored into the
owner's HBM over NVLink and re-ranked there from local + peer fp32 masters, kernel (b) reading peer shards over
CUDA IPC, histogram all-reduce; a further pass forces the NCCL all-to-all + merge path) against the single-GPU
result: indices, flags and the histogram bit-identical, and the same output digest bench.py prints.
Skipped for world sizes the box cannot serve (run with `gpurun --gpus N`)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


TAGS = ("peer", "peer2", "peer3", "pipe", "pull", "nccl")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    sys.path.insert(0, str(ROOT))
    from oracle import tvc_oracle as O
    g = O.synth_gallery(20000, 256, seed=3, clusters=128, dup_rate=1e-3)
    bank = O.synth_gallery(3000, 256, seed=4, clusters=128)
    img, txt, var = O.synth_queries(g, 1500, 5, seed=5)
    return g, bank, img, txt, var


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    # file rendezvous: a port probed free by the parent can be taken again before rank 0 listens on it
    dist.init_process_group("nccl", init_method=f"file://{out_dir}/rendezvous", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    sys.path.insert(0, str(ROOT))
    from multimodal_detection_consistency_b200.pipeline import TVCScorer, shard_bounds
    g, bank, img, txt, var = _data()
    glo, ghi = shard_bounds(len(g), world, rank)
    blo, bhi = shard_bounds(len(bank), world, rank)
    sc = TVCScorer(g[glo:ghi], bank[blo:bhi], k=10, total_gallery_rows=len(g), total_bank_rows=len(bank),
                   device=f"cuda:{rank}")
    assert sc._gallery_group is not None          # the peer-memory path, not the staged fetch
    assert sc._exchange is not None
    ex = sc._exchange
    import bench
    for tag in TAGS:
        if tag == "nccl":
            sc._exchange = None                   # candidate all-to-all + merge kernel instead
        sc.rescore_at_shards = tag != "pull"      # "pull": round-1 phase 2 (owner reads peer masters)
        sc.min_chunk_queries = 32 if tag == "pipe" else 1024     # "pipe": host batch pipelined in pieces
        sc.reset_hubness()
        if tag == "peer3":
            # skewed arrival: every rank enters the batch at a different time (the double-buffered receive
            # areas + stream-ordered barriers must not let a fast rank overwrite what a slow one still reads)
            torch.cuda._sleep(int(2e8) * (rank % 3))
            sc.score_batch(img, txt, var)
            torch.cuda._sleep(int(3e8) * ((world - rank) % 4))
            sc.reset_hubness()
        out = sc.score_batch(img, txt, var, to_host=True)   # (peer2: the second receive buffer)
        lo, hi = out["slice"]
        torch.cuda.synchronize()
        v, k = int(var.shape[1]), 10
        dg = (bench.digest64(torch, out["topk_idx"], lo * v * k, 1) + bench.digest64(torch, out["bank_idx"], lo * v * k, 2)
              + bench.digest64(torch, out["flags"], lo, 3)).reshape(1).cuda()
        dist.all_reduce(dg)
        dg = int((dg.cpu() + bench.digest64(torch, sc.k_occurrence.cpu(), 0, 4)).item())
        np.savez(Path(out_dir) / f"{tag}_r{rank}.npz", lo=lo, hi=hi, scores=out["scores"].numpy(),
                 flags=out["flags"].numpy(), topk_idx=out["topk_idx"].numpy(), topk_sim=out["topk_sim"].numpy(),
                 bank_idx=out["bank_idx"].numpy(), hub=sc.k_occurrence.cpu().numpy(), digest=dg)
    sc._exchange = ex
    sc.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import bench
    from multimodal_detection_consistency_b200.pipeline import TVCScorer
    g, bank, img, txt, var = _data()
    ref = TVCScorer(g, bank, k=10, device="cuda:0")
    want = ref.score_batch(img, txt, var, to_host=True)
    want = {k: (v.numpy().copy() if hasattr(v, "numpy") else v) for k, v in want.items()}
    hub = ref.k_occurrence.cpu().numpy()
    t = {n: torch.from_numpy(want[n]) for n in ("topk_idx", "bank_idx", "flags")}
    want_digest = int((bench.digest64(torch, t["topk_idx"], 0, 1) + bench.digest64(torch, t["bank_idx"], 0, 2)
                       + bench.digest64(torch, t["flags"], 0, 3) + bench.digest64(torch, torch.from_numpy(hub), 0, 4)).item())
    del ref
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for tag in TAGS:
        covered = 0
        for r in range(world):
            z = np.load(tmp_path / f"{tag}_r{r}.npz")
            lo, hi = int(z["lo"]), int(z["hi"])
            covered += hi - lo
            assert np.array_equal(z["topk_idx"], want["topk_idx"][lo:hi]), tag
            assert np.array_equal(z["bank_idx"], want["bank_idx"][lo:hi]), tag
            assert np.abs(z["topk_sim"] - want["topk_sim"][lo:hi]).max() <= 1e-6
            assert np.abs(z["scores"] - want["scores"][lo:hi]).max() <= 1e-5
            assert np.array_equal(z["flags"], want["flags"][lo:hi])
            assert np.array_equal(z["hub"], hub)
            assert int(z["digest"]) == want_digest, (tag, r)
        assert covered == len(img)
