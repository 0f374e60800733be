"""End-to-end (pinned host in, host out) step time of TVCScorer.score_batch for several host-batch
pipelining splits on the bench workload, one process.  python scripts/e2e_splits.py [--queries Q]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from multimodal_detection_consistency_b200.pipeline import TVCScorer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    args = argparse.Namespace(queries=a.queries, variants=5, gallery=1_000_000, bank=100_000, dim=768, topk=10)
    dev = torch.device("cuda", 0)
    g, centers = bench.synth_device(torch, args, dev, args.gallery, 42)
    b, _ = bench.synth_device(torch, args, dev, args.bank, 43, centers=centers)
    scorer = TVCScorer(g, b, k=10, device=dev)
    del g, b
    img, txt, var = bench.synth_queries(torch, args, dev, centers, 123)
    h = [t.cpu().pin_memory() for t in (img, txt, var)]
    scorer.min_chunk_queries = 512

    def timed(fn, steps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    res = {"resident": timed(lambda: scorer.score_batch(img, txt, var), a.steps, 3)}
    for split in [1, 2, 4, 8, (1, 7), (1, 4, 3), (1, 3, 3, 1), (1, 5, 2), (1, 2, 2, 2, 1), (2, 5, 1), (1, 6, 1)]:
        scorer.host_chunks = split
        res[str(split)] = timed(lambda: scorer.score_batch(*h, to_host=True), a.steps)
    res["resident_again"] = timed(lambda: scorer.score_batch(img, txt, var), a.steps, 1)
    print(json.dumps({k: round(v, 3) for k, v in res.items()}))


if __name__ == "__main__":
    main()
