"""Pipeline trace of kernel (b)'s embedding-mode kernel (block 0): per query, ns since the first event
of: producer start / producer armed / first unit after full / done passed / stage released / finished."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import multimodal_detection_consistency_b200 as tvc

ctx = tvc.Context.get(0)
dev = torch.device("cuda:0")
d, n, q, V, G = 768, 200_000, 16384, 5, 3
g = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
gal = tvc.Gallery(g, ctx=ctx)
img = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
txt = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
var = torch.nn.functional.normalize(txt[:, None, :] + 0.01 * torch.randn(q, V, d, device=dev), dim=2)
_, ridx = gal.search(var, 10)
ridx = ridx.reshape(q, V * 10).contiguous()
gen = torch.nn.functional.normalize(torch.randn(q, G, d, device=dev), dim=2)
p = tvc.default_params()
for _ in range(2):
    ctx.consistency_emb(p, img, txt, var, ret_gallery=gal, ret_idx=ridx, gen=gen)
trace = torch.full((512 * 8,), 2**62, dtype=torch.int64, device=dev)
trace.view(512, 8)[:, 6] = 0
ctx.set_option("emb_trace_ptr", trace.data_ptr())
ctx.consistency_emb(p, img, txt, var, ret_gallery=gal, ret_idx=ridx, gen=gen)
torch.cuda.synchronize()
ctx.set_option("emb_trace_ptr", 0)
t = trace.view(512, 8).cpu()
t0 = int(t[0, 0])
names = ["prod_start", "prod_armed", "first_unit", "done", "released", "finished"]
print("query " + " ".join(f"{n:>10s}" for n in names))
for i in list(range(0, 24)) + list(range(100, 111)):
    if int(t[i, 0]) >= 2**62:
        break
    print(f"{i:5d} " + " ".join(f"{int(t[i, k]) - t0:10d}" for k in range(6)) +
          f"   longest task {int(t[i, 6]) >> 8} ns (task {int(t[i, 6]) & 255})")
