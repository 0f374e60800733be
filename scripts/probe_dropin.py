"""Times MultiModalRetriever.batch_retrieve_images_by_texts(2048 texts, top_k=5) against the bench's clustered 1M-row
gallery, call by call, and the bare search underneath."""
import sys, time, types
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
import multimodal_detection_consistency_b200 as tvc
from multimodal_detection_consistency_b200 import MultiModalRetriever, RetrievalConfig

args = types.SimpleNamespace(dim=768, queries=2048, variants=5, gallery=1_000_000, bank=0, topk=10)
dev = torch.device("cuda:0")
g_rows, centers = bench.synth_device(torch, args, dev, args.gallery, 42)
img, txt, var = bench.synth_queries(torch, args, dev, centers, 123)
g = g_rows.cpu().numpy(); t = txt.cpu().numpy()
q = t.shape[0]
rows = {f"t{i}": t[i] for i in range(q)}

class Enc:
    def encode_text(self, texts, normalize=True):
        return np.stack([rows[s] for s in texts])

r = MultiModalRetriever(RetrievalConfig(top_k=10, enable_cache=False), clip_model=Enc())
r.build_image_index_from_features(g, [f"img_{i}.jpg" for i in range(args.gallery)])
texts = [f"t{i}" for i in range(q)]
ctx = tvc.Context.get(0)
for m in (64, 2048, 2048, 512, 2048):
    l0 = ctx.launch_count()
    t0 = time.perf_counter(); res = r.batch_retrieve_images_by_texts(texts[:m], top_k=5); dt = time.perf_counter() - t0
    print(f"batch_retrieve m={m:5d}: {dt*1e3:9.2f} ms  ({m/dt:9.0f} queries/s, {ctx.launch_count() - l0} launches)", flush=True)
gal = tvc.Gallery(g_rows, ctx=ctx)
tq = txt[:2048].contiguous()
for k in (5, 10):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); s, i = gal.search(tq, k); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"Gallery.search m=2048 k={k} device tensors: {dt*1e3:.2f} ms", flush=True)
