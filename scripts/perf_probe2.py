"""Sustained-load probe: runs tvc_search back to back for ~3 s per configuration and reports TFLOP/s with
SM clocks / power sampled during the run.  usage: perf_probe2.py MxNxD [opt=value ...]"""
import subprocess, sys, threading, time, statistics
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import multimodal_detection_consistency_b200 as tvc

ctx = tvc.Context.get(0)
dev = torch.device("cuda:0")
m, n, d = (int(x) for x in sys.argv[1].split("x"))
g = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
q = torch.nn.functional.normalize(torch.randn(m, d, device=dev), dim=1)
gal = tvc.Gallery(g, ctx=ctx)
del g
configs = [c for c in sys.argv[2:]] or ["default"]
for cfg in configs:
    ctx.set_option("pair_min_rows", 4096); ctx.set_option("debug_flags", 0)
    if cfg != "default":
        for kv in cfg.split(","):
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
    for _ in range(2):
        gal.search(q, 10)
    torch.cuda.synchronize()
    lines = []
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    threading.Thread(target=lambda: [lines.append(l) for l in p.stdout], daemon=True).start()
    ctx.set_timing(True); ctx.search_kernel_ms()
    t0 = time.time(); reps = 0
    while time.time() - t0 < 3.0:
        gal.search(q, 10); reps += 1
        if reps % 4 == 0: torch.cuda.synchronize()
    torch.cuda.synchronize()
    kms, nl = ctx.search_kernel_ms(); ctx.set_timing(False)
    p.terminate()
    clk = [float(l.split(",")[0]) for l in lines if "," in l]; pw = [float(l.split(",")[1]) for l in lines if "," in l]
    clk2 = clk[len(clk)//3:] or [0]; pw2 = pw[len(pw)//3:] or [0]
    fl = 2.0 * m * n * d
    print(f"{cfg:40s} kernel {kms/nl:8.3f} ms  {fl/(kms/nl)/1e9:7.1f} TFLOP/s  sm {statistics.median(clk2):.0f} MHz  {statistics.median(pw2):.0f} W  "
          f"({fl/(kms/nl)/1e9/(148*8192*statistics.median(clk2)*1e6/1e12)*100:.1f}% of clock peak)", flush=True)
