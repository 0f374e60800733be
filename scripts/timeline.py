"""Kernel timeline of one TVC step (torch.profiler / CUPTI): start, duration and the idle gap before every
kernel, to see what the non-GEMM share of the step is made of.  usage: timeline.py [queries]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
from multimodal_detection_consistency_b200.pipeline import TVCScorer

args = bench.parse_args.__wrapped__() if hasattr(bench.parse_args, "__wrapped__") else None
sys.argv = sys.argv[:1]
args = bench.parse_args()
dev = torch.device("cuda:0")
g, centers = bench.synth_device(torch, args, dev, args.gallery, 42)
b, _ = bench.synth_device(torch, args, dev, args.bank, 43, centers=centers)
sc = TVCScorer(g, b, k=args.topk, device=dev)
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
print(f"span {prev_end - t0:.1f} us, kernels {tot_k:.1f} us, idle {prev_end - t0 - tot_k:.1f} us")
