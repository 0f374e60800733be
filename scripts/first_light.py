"""First-light diagnostic for libtvc on a GPU box: prints mismatch statistics instead of asserting."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import multimodal_detection_consistency_b200 as tvc
from oracle import tvc_oracle as O

ctx = tvc.Context.get(0)
rng = np.random.default_rng(0)
for (m, n, d, k) in [(128, 256, 64, 10), (128, 256, 128, 10), (200, 1000, 512, 10), (1000, 5000, 768, 10)]:
    g = O.l2_normalize(rng.standard_normal((n, d), dtype=np.float32))
    q = O.l2_normalize(rng.standard_normal((m, d), dtype=np.float32))
    gal = tvc.Gallery(g, ctx=ctx)
    S = gal.similarity_matrix(q)
    ref = O.bf16_round(q) @ O.bf16_round(g).T
    err = np.abs(S - ref)
    print(f"[simmat] m={m} n={n} d={d}: max err {err.max():.3e} mean {err.mean():.3e} "
          f"bad(>1e-4)={int((err > 1e-4).sum())}", flush=True)
    if err.max() > 1e-3:
        r, c = np.unravel_index(err.argmax(), err.shape)
        print("   worst at", r, c, "got", S[r, c], "want", ref[r, c])
        print("   row err by col block of 32:", [(int(b), float(err[:, b:b+32].max())) for b in range(0, min(n, 256), 32)])
        print("   err by row block of 32:", [(int(b), float(err[b:b+32].max())) for b in range(0, min(m, 128), 32)])
    t0 = time.time()
    sims, idx = gal.search(q, k)
    rs, ri = O.search(q, g, k)
    print(f"[search] m={m} n={n} d={d} k={k}: idx mismatch {(idx != ri).mean():.4f} "
          f"sim err {np.abs(sims - rs).max():.3e}  ({time.time() - t0:.2f}s)", flush=True)
print("launches", ctx.launch_count())
