"""Roofline probe for the bandwidth-bound kernels: (b) consistency reduction (similarity-fed and
embedding-fed) and (c) k-occurrence histogram.  Prints achieved algorithmic GB/s against the measured
HBM peak (MEASURED_PEAKS.json).  usage: bench_bc.py [quick]"""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import multimodal_detection_consistency_b200 as tvc

peak = 6533.8
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
ctx = tvc.Context.get(0)
dev = torch.device("cuda:0")
quick = len(sys.argv) > 1


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- (b) similarity-fed: 4*(1+V+R+G) + 4*X read, 96 + 1 written per query
V, R, G = 5, 10, 3
X = V * (V - 1) // 2
for q in ([1 << 20] if quick else [1 << 16, 1 << 20, 1 << 22]):
    s0 = torch.rand(q, device=dev)
    sv, sr, sg, sx = (torch.rand(q, w, device=dev) for w in (V, R, G, X))
    rc = torch.randint(0, R + 1, (q,), device=dev, dtype=torch.int32)
    gc = torch.randint(0, G + 1, (q,), device=dev, dtype=torch.int32)
    p = tvc.default_params()
    ms = timeit(lambda: ctx.consistency_sims(p, s0, sv, sr, rc, sg, gc, sx))
    byt = q * (4 * (1 + V + R + G + X) + 8 + 96 + 1)
    print(f"(b) sims  Q={q:8d}: {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s = {byt/ms/1e6/peak*100:5.1f}% of measured HBM peak  "
          f"({q/ms/1e3:.1f} M queries/s)", flush=True)

# ---- (b) embedding-fed: fp32 rows; bytes = 4*d*(2 + V + refs gathered) + outputs
d = 768
n = 200_000
g = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
gal = tvc.Gallery(g, ctx=ctx)
for q in ([16384] if quick else [2048, 16384, 65536]):
    img = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
    var = torch.nn.functional.normalize(txt[:, None, :] + 0.01 * torch.randn(q, V, d, device=dev), dim=2)
    _, ridx = gal.search(var, 10)
    ridx = ridx.reshape(q, V * 10).contiguous()
    gen = torch.nn.functional.normalize(torch.randn(q, G, d, device=dev), dim=2)
    p = tvc.default_params()
    res = {}

    def run():
        res["o"] = ctx.consistency_emb(p, img, txt, var, ret_gallery=gal, ret_idx=ridx, gen=gen)
    ms = timeit(run, 5)
    nret = float(res["o"][0][:, 21].mean())
    byt = q * (4 * d * (2 + V + G + nret) + V * 10 * 8 + 96 + 1)
    print(f"(b) emb   Q={q:8d}: {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s = {byt/ms/1e6/peak*100:5.1f}% of measured HBM peak  "
          f"({q/ms/1e3:.2f} M queries/s, {nret:.1f} refs/query)", flush=True)

# ---- (b) README reference-vector rule: 4*d*(1 + V*(k+m)) read per query, every row used once
kk, mm = 5, 3
for q in ([16384] if quick else [2048, 16384]):
    img = torch.nn.functional.normalize(torch.randn(q, d, device=dev), dim=1)
    ridx = torch.randint(0, n, (q, V, kk), device=dev)
    gen4 = torch.nn.functional.normalize(torch.randn(q, V, mm, d, device=dev), dim=3)
    ms = timeit(lambda: ctx.reference_vector_rule(img, gal, ridx, gen4, want_ref=False), 5)
    ms_ref = timeit(lambda: ctx.reference_vector_rule(img, gal, ridx, gen4), 5)
    byt = q * (4 * d * (1 + V * (kk + mm)) + V * kk * 8 + V * 4 + 9)
    print(f"(b) refvec Q={q:7d}: {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s = {byt/ms/1e6/peak*100:5.1f}% of measured HBM peak  "
          f"({q/ms/1e3:.2f} M queries/s; with the cross-variant Reference Vector {ms_ref*1e3:.1f} us)", flush=True)

# ---- (c) k-occurrence: 8*M*k read + 4*N written (+ 4*N zero fill)
for (m, k, nb) in ([(5_000_000, 10, 1_000_000)] if quick else [(50_000, 10, 118_287), (500_000, 10, 1_000_000), (5_000_000, 10, 1_000_000), (5_000_000, 10, 10_000)]):
    idx = (nb * torch.rand(m, k, device=dev) ** 3).long().clamp_(0, nb - 1)      # skewed: hubs
    counts = torch.zeros(nb, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ctx.k_occurrence(idx, nb))
    byt = 8 * m * k + 8 * nb
    print(f"(c) kocc  M*k={m*k:9d} N={nb:8d}: {ms*1e3:9.1f} us  {byt/ms/1e6:8.1f} GB/s = {byt/ms/1e6/peak*100:5.1f}% of measured HBM peak", flush=True)
