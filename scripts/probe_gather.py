"""Is kernel (b)'s embedding mode limited by random 3 KB row gathers?  Same launch with (1) the real
top-k indices, (2) sequential reference rows, plus the plain row-gather kernel as a reference point."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import multimodal_detection_consistency_b200 as tvc

ctx = tvc.Context.get(0)
dev = torch.device("cuda:0")
d, n, q, V, G = 768, 200_000, 16384, 5, 3


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
gal = tvc.Gallery(g, ctx=ctx)
img = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
txt = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
var = torch.nn.functional.normalize(txt[:, None, :] + 0.01 * torch.randn(q, V, d, device=dev), dim=2)
gen = torch.nn.functional.normalize(torch.randn(q, G, d, device=dev), dim=2)
_, ridx = gal.search(var, 10)
real = ridx.reshape(q, V * 10).contiguous()
seq = (torch.arange(q * 10, device=dev).reshape(q, 10) % n).repeat(1, V).contiguous()
rnd = torch.randint(0, n, (q, V * 10), device=dev)
p = tvc.default_params(dedup_threshold=-2.0)
byt = q * 4 * d * 20
for name, idx in (("top-k indices", real), ("sequential rows", seq), ("uniform random rows", rnd)):
    ms = timeit(lambda: ctx.consistency_emb(p, img, txt, var, ret_gallery=gal, ret_idx=idx, gen=gen))
    print(f"emb kernel, {name:20s}: {ms*1e3:7.1f} us  {byt/ms/1e6:7.1f} GB/s")
# plain gather of the same number of random rows (10 per query) and of sequential rows
for name, idx in (("random", rnd[:, :10].reshape(-1).contiguous()), ("sequential", seq[:, :10].reshape(-1).contiguous())):
    ms = timeit(lambda: gal.get_rows(idx))
    b = idx.numel() * d * 4 * 2
    print(f"gather_rows ({name:10s}) {idx.numel()} rows: {ms*1e3:7.1f} us  {b/ms/1e6:7.1f} GB/s (read+write)")
