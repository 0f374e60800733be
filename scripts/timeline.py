"""Kernel timeline of one TVC step (torch.profiler / CUPTI): start, duration and the idle gap before every
kernel, to see what the non-GEMM share of the step is made of.  usage: timeline.py [queries]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from multimodal_detection_consistency_b200.pipeline import TVCScorer

import os
import torch.distributed as dist
from multimodal_detection_consistency_b200.pipeline import shard_bounds
sys.argv = sys.argv[:1]
args = bench.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
glo, ghi = shard_bounds(args.gallery, world, rank)
blo, bhi = shard_bounds(args.bank, world, rank)
g, centers = bench.synth_device(torch, args, dev, args.gallery, 42, glo, ghi)
b, _ = bench.synth_device(torch, args, dev, args.bank, 43, blo, bhi, centers)
sc = TVCScorer(g, b, k=args.topk, device=dev, total_gallery_rows=args.gallery, total_bank_rows=args.bank)
del g, b
img, txt, var = bench.synth_queries(torch, args, dev, centers, 123)
for _ in range(3):
    sc.score_batch(img, txt, var)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        sc.score_batch(img, txt, var)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
if rank != 0:
    dist.destroy_process_group()
    sys.exit(0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
tot_k = 0.0
print(f"{'start us':>10s} {'dur us':>10s} {'gap us':>8s}  kernel")
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    gap = s - prev_end
    print(f"{s - t0:10.1f} {en - s:10.1f} {gap:8.1f}  {e.name[:90]}")
    prev_end = max(prev_end, en)
    tot_k += en - s
if world > 1:
    dist.destroy_process_group()
print(f"span {prev_end - t0:.1f} us, kernels {tot_k:.1f} us, idle {prev_end - t0 - tot_k:.1f} us")
