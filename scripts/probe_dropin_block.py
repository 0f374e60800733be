"""Runs only bench.py's latency / dropin block (the reference's API with table encoders) on the bench's gallery."""
import json, sys, types
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench

args = types.SimpleNamespace(dim=768, queries=2048, variants=5, gallery=1_000_000, bank=100_000, topk=10)
dev = torch.device("cuda:0")
g, centers = bench.synth_device(torch, args, dev, args.gallery, 42)
b, _ = bench.synth_device(torch, args, dev, args.bank, 43, centers=centers)
img, txt, var = bench.synth_queries(torch, args, dev, centers, 123)
g_host, b_host = g.cpu(), b.cpu()
del g, b
latency, dropin = bench.dropin_blocks(np, args, g_host, b_host, *(t.cpu().numpy() for t in (img, txt, var)))
print(json.dumps(dict(latency=latency, dropin=dropin)))
