"""Quick device-side timing of tvc_search on large shapes (inputs resident in HBM)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import multimodal_detection_consistency_b200 as tvc

ctx = tvc.Context.get(0)
ctx.set_timing(True)
dev = torch.device("cuda:0")
shapes = [(18944, 1_000_000, 768), (81920, 1_000_000, 768), (25000, 36000, 768), (5000, 5000, 512), (50000, 118287, 768)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for (m, n, d) in shapes:
    g = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
    q = torch.nn.functional.normalize(torch.randn(m, d, device=dev), dim=1)
    gal = tvc.Gallery(g, ctx=ctx)
    del g
    for _ in range(2):
        gal.search(q, 10)
    torch.cuda.synchronize()
    ctx.search_kernel_ms()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        gal.search(q, 10)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / reps
    kms, nl = ctx.search_kernel_ms()
    kms /= max(nl, 1)
    flops = 2.0 * m * n * d
    print(f"m={m} n={n} d={d}: search {total:.3f} ms/call, gemm_topk kernel {kms:.3f} ms "
          f"-> {flops / kms / 1e9:.1f} TFLOP/s kernel, {flops / total / 1e9:.1f} TFLOP/s call", flush=True)
    gal.close()
    del q
    torch.cuda.empty_cache()
