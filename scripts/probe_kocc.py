"""Kernel (c) probe: single-pass (RED) path against the bucketed two-pass path on index streams of a given size.
usage: probe_kocc.py [entries_in_millions] [bins]   (TVC_KOCC_PART_OCC=2|3 pins the partition kernel's occupancy)"""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import multimodal_detection_consistency_b200 as tvc

ctx = tvc.Context.get(0)
dev = torch.device("cuda:0")
peak = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
mm = float(sys.argv[1]) if len(sys.argv) > 1 else 50.0
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
m = int(mm * 1e6) // 10


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


streams = {
    "skewed N*u^3": lambda: (nb * torch.rand(m, 10, device=dev) ** 3).long().clamp_(0, nb - 1),
    "uniform": lambda: torch.randint(0, nb, (m, 10), device=dev),
}
for name, make in streams.items():
    idx = make()
    want = torch.bincount(idx.reshape(-1), minlength=nb).to(torch.int32)
    byt = 8 * m * 10 + 4 * nb
    for label, part_min in (("single pass", (1 << 63) - 1), ("bucketed", 0)):
        ctx.set_option("kocc_part_min", part_min)
        got = ctx.k_occurrence(idx, nb)
        torch.cuda.synchronize()
        ok = bool(torch.equal(got, want))
        cnt = torch.zeros(nb, dtype=torch.int32, device=dev)
        ms = timeit(lambda: ctx.k_occurrence(idx, nb, 0, cnt))
        print(f"(c) {name:13s} M*k={m*10:9d} N={nb:8d} {label:11s}: {ms*1e3:8.1f} us  {byt/ms/1e6:7.1f} GB/s = "
              f"{byt/ms/1e6/peak*100:5.1f}% of measured HBM peak  exact={ok}", flush=True)
    del idx, want
