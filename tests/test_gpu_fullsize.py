"""Parity at BASELINE.json's full sizes.  C1 and C2 are checked element by element against the oracle;
C3 and the 1M-row north_star gallery are checked through size-independent properties (sharded merge ==
unsharded search, every similarity is the fp32 dot of its index, planted duplicates come back first with
ties to the lower index, sortedness, histogram identities) plus an oracle check on a row sample."""
import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu


def _band_check(sims, idx, ref_s, ref_i):
    bad = idx != ref_i
    assert np.abs(sims - ref_s).max() <= 2e-3
    if bad.any():
        assert np.abs(sims[bad] - ref_s[bad]).max() <= 1e-3
    return bad.mean()


def test_c1_cpu_reference_config(tvc_ctx):
    """configs[0]: 1k queries x 5 variants, top-10 over a 5k-image 512-d gallery."""
    import multimodal_detection_consistency_b200 as tvc
    g = O.synth_gallery(5000, 512, seed=42)
    img, txt, var = O.synth_queries(g, 1000, 5, seed=123)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(var, 10)
    rs, ri = O.search(var.reshape(-1, 512), g, 10)
    assert _band_check(sims.reshape(-1, 10), idx.reshape(-1, 10), rs, ri) < 0.005
    counts = tvc_ctx.k_occurrence(idx.reshape(-1, 10), 5000)
    assert np.array_equal(counts, O.k_occurrence(idx, 5000))


def test_c2_flickr_scale_full_pipeline(tvc_ctx):
    """configs[1]: 5k queries x 5 variants vs 1k images + 25k captions (768-d) + 10k reference bank."""
    from multimodal_detection_consistency_b200.pipeline import TVCScorer
    d = 768
    g = O.synth_gallery(26000, d, seed=1)
    bank = O.synth_gallery(10000, d, seed=2)
    img, txt, var = O.synth_queries(g, 5000, 5, seed=3)
    sc = TVCScorer(g, bank, k=10, device="cuda:0")
    out = sc.score_batch(img, txt, var, to_host=True)
    rows = var.reshape(-1, d)
    rs, ri = O.search(rows, g, 10)
    frac = _band_check(out["topk_sim"].numpy().reshape(-1, 10), out["topk_idx"].numpy().reshape(-1, 10), rs, ri)
    assert frac < 0.005
    bs, bi = O.search(rows, bank, 10)
    _band_check(out["bank_sim"].numpy().reshape(-1, 10), out["bank_idx"].numpy().reshape(-1, 10), bs, bi)
    # scoring is checked on the GPU's own candidate lists (decouples it from in-band index swaps)
    sub = slice(0, 600)
    ref, rflags, _ = O.consistency_emb(img[sub], txt[sub], var[sub], ret_rows=g,
                                       ret_idx=out["topk_idx"].numpy()[sub].reshape(600, 50), gen_rows=bank,
                                       gen_idx=out["bank_idx"].numpy()[sub].reshape(600, 50))
    assert np.abs(out["scores"].numpy()[sub] - ref).max() <= 1e-5
    ok = (np.abs(ref[:, O.S_DET_AGG] - 0.5) > 1e-5) & (np.abs(ref[:, O.S_CC_OVERALL] - ref[:, O.S_CC_THRESHOLD]) > 1e-5)
    assert np.array_equal((out["flags"].numpy()[sub] & 3)[ok], (rflags & 3)[ok])
    assert np.array_equal(sc.k_occurrence.cpu().numpy(), O.k_occurrence(out["topk_idx"].numpy(), 26000))


def test_c3_hubness_coco_train_scale(tvc_ctx):
    """configs[2]: k-occurrence (k=10) of 50k queries over a 118,287-image 768-d gallery."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    n, m, d, k = 118287, 50000, 768, 10
    gen = torch.Generator(device="cuda").manual_seed(7)
    g = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=gen), dim=1)
    hub = torch.nn.functional.normalize(torch.randn(d, device="cuda", generator=gen), dim=0)
    q = torch.nn.functional.normalize(torch.randn(m, d, device="cuda", generator=gen) / d ** 0.5 + 0.6 * hub, dim=1)
    g[777] = hub                                       # an adversarial hub: near every query
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(q, k)
    counts = tvc_ctx.k_occurrence(idx, n)
    torch.cuda.synchronize()
    assert int(counts.sum()) == m * k and int(counts[777]) == m          # the hub is in every top-k
    assert torch.equal(counts.cpu(), torch.bincount(idx.reshape(-1).cpu(), minlength=n).int())
    assert bool((sims[:, :-1] >= sims[:, 1:]).all())                      # sorted
    dots = (q[:, None, :] * g[idx]).sum(-1)                               # every sim is the fp32 dot of its index
    assert float((dots - sims).abs().max()) <= 1e-5
    sel = np.random.default_rng(0).choice(m, 1500, replace=False)
    rs, ri = O.search(q[sel].cpu().numpy(), g.cpu().numpy(), k)
    assert _band_check(sims[sel].cpu().numpy(), idx[sel].cpu().numpy(), rs, ri) < 0.005
    # self k-NN of a gallery slice: never returns itself
    s2, i2 = gal.search(g[:4096], k, skip_self=True)
    torch.cuda.synchronize()
    assert bool((i2 != torch.arange(4096, device="cuda")[:, None]).all())


def test_1m_gallery_properties(tvc_ctx):
    """north_star gallery (1M x 768): sharded search + merge == unsharded, planted rows, sampled oracle."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    n, d, k, m = 1_000_000, 768, 10, 6000
    gen = torch.Generator(device="cuda").manual_seed(11)
    g = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=gen), dim=1)
    q = torch.nn.functional.normalize(torch.randn(m, d, device="cuda", generator=gen), dim=1)
    # plant every 10th query as an exact gallery row, twice (tie -> the lower index must come first)
    planted = torch.arange(0, m, 10, device="cuda")
    lo_idx = 1000 + 37 * torch.arange(len(planted), device="cuda")
    hi_idx = 900_000 + 11 * torch.arange(len(planted), device="cuda")
    g[lo_idx] = q[planted]
    g[hi_idx] = q[planted]
    full = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = full.search(q, k)
    torch.cuda.synchronize()
    assert torch.equal(idx[planted, 0], lo_idx) and torch.equal(idx[planted, 1], hi_idx)
    assert float((sims[planted, 0] - 1).abs().max()) <= 1e-5 and torch.equal(sims[planted, 0], sims[planted, 1])
    assert bool((sims[:, :-1] >= sims[:, 1:]).all())
    dots = (q[:, None, :] * g[idx]).sum(-1)
    assert float((dots - sims).abs().max()) <= 1e-5
    # 4 row shards searched independently, merged: identical to the unsharded result
    parts_s, parts_i = [], []
    for r in range(4):
        a, b = r * 250_000, (r + 1) * 250_000
        shard = tvc.Gallery(g[a:b], global_row_offset=a, ctx=tvc_ctx)
        s, i = shard.search(q, k)
        parts_s.append(s)
        parts_i.append(i)
        torch.cuda.synchronize()
        shard.close()
    ms, mi = tvc_ctx.merge_topk(torch.stack(parts_s, 1), torch.stack(parts_i, 1), k)
    torch.cuda.synchronize()
    assert torch.equal(mi, idx) and torch.equal(ms, sims)
    # sampled rows against a plain torch fp32 reference (full row of similarities)
    sel = torch.randperm(m, device="cuda", generator=gen)
    sel = sel[sel % 10 != 0][:256]          # planted rows tie exactly; torch.topk's tie order is unspecified
    ref = torch.topk(q[sel] @ g.T, k, dim=1)
    same = ref.indices == idx[sel]
    assert float(same.float().mean()) > 0.995
    assert float((ref.values - sims[sel]).abs().max()) <= 1e-3
    diff = (ref.values - sims[sel]).abs()
    assert (not bool((~same).any())) or float(diff[~same].max()) <= 1e-3


def test_pipelined_host_batches_equal_one_shot(tvc_ctx):
    """TVCScorer's chunked upload/compute/download pipeline returns exactly the one-shot result."""
    import torch
    from multimodal_detection_consistency_b200.pipeline import TVCScorer
    d = 256
    g = O.synth_gallery(20000, d, seed=4)
    bank = O.synth_gallery(3000, d, seed=5)
    img, txt, var = (torch.from_numpy(x).pin_memory() for x in O.synth_queries(g, 1000, 5, seed=6))
    outs = []
    for chunks in (1, 4, (1, 6, 1)):
        sc = TVCScorer(g, bank, k=10, device="cuda:0")
        sc.host_chunks, sc.min_chunk_queries = chunks, 100
        o = sc.score_batch(img, txt, var, to_host=True)
        outs.append(({k: v.clone() for k, v in o.items() if k != "slice"}, sc.k_occurrence.cpu().clone()))
        assert o["slice"] == (0, 1000)
        o2 = sc.score_batch(img, txt, var)            # device-resident results, same pipeline
        torch.cuda.synchronize()
        for name, t in outs[-1][0].items():
            assert torch.equal(o2[name].cpu(), t), name
    for other in outs[1:]:
        for name, t in outs[0][0].items():
            assert torch.equal(t, other[0][name]), name
        assert torch.equal(outs[0][1], other[1])


def test_c4_cc3m_scale_properties(tvc_ctx):
    """configs[3]: 100k queries x 5 variants vs a 3M-image 768-d gallery (one GPU holds it: 4.6 GB bf16 +
    9.2 GB fp32).  Size-independent properties on all 500k rows, the two-phase sharded search on a
    sample, a torch fp32 reference on a sample, and the hubness histogram identities."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    n, d, k, nq, v = 3_000_000, 768, 10, 100_000, 5
    gen = torch.Generator(device="cuda").manual_seed(21)
    g = torch.empty((n, d), device="cuda")
    for a in range(0, n, 500_000):
        g[a:a + 500_000] = torch.nn.functional.normalize(torch.randn(500_000, d, device="cuda", generator=gen), dim=1)
    base = g[torch.randint(0, n, (nq,), device="cuda", generator=gen)]
    var = torch.nn.functional.normalize(
        base[:, None, :] + (0.6 / d ** 0.5) * torch.randn(nq, v, d, device="cuda", generator=gen), dim=2)
    rows = var.view(nq * v, d)
    full = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = full.search(rows, k)
    torch.cuda.synchronize()
    assert bool((idx >= 0).all()) and bool((idx < n).all())
    assert bool((sims[:, :-1] >= sims[:, 1:]).all())
    tie = sims[:, :-1] == sims[:, 1:]
    assert bool((idx[:, :-1][tie] < idx[:, 1:][tie]).all())           # ties to the lower index
    assert bool((idx.sort(dim=1).values[:, 1:] != idx.sort(dim=1).values[:, :-1]).all())   # no row twice
    for a in range(0, nq * v, 50_000):                                  # every similarity is its index's fp32 dot
        dots = (rows[a:a + 50_000, None, :] * g[idx[a:a + 50_000]]).sum(-1)
        assert float((dots - sims[a:a + 50_000]).abs().max()) <= 1e-5
    # hubness histogram: integer identities + the oracle on the same indices
    counts = tvc_ctx.k_occurrence(idx, n)
    torch.cuda.synchronize()
    assert int(counts.sum()) == nq * v * k
    assert torch.equal(counts, torch.bincount(idx.reshape(-1), minlength=n).to(torch.int32))
    # two-phase sharded search (3 shards, 2 slice owners) on a 20k-row sample == the unsharded result
    from multimodal_detection_consistency_b200._native import Scatter
    sub = rows[:20_000]
    per, kp, rps = 1_000_000, tvc_ctx.candidate_width(k), 10_000
    parts = [tvc.Gallery.wrap_rows(g[r * per:(r + 1) * per], r * per, ctx=tvc_ctx) for r in range(3)]
    group = tvc.Gallery.group(parts)
    val = [torch.empty((3, rps, kp), device="cuda") for _ in range(2)]
    ix = [torch.empty((3, rps, kp), dtype=torch.int64, device="cuda") for _ in range(2)]
    for r in range(3):
        shard = tvc.Gallery(g[r * per:(r + 1) * per], global_row_offset=r * per, keep_master=False, ctx=tvc_ctx)
        sc = Scatter()
        sc.n_slices, sc.slot, sc.rows_per_slice = 2, r, rps
        for j in range(2):
            sc.val[j], sc.idx[j] = val[j].data_ptr(), ix[j].data_ptr()
        shard.search_candidates(sub, k, sc)
        torch.cuda.synchronize()
        shard.close()
    for j in range(2):
        s2, i2 = tvc_ctx.rerank_candidates(group, sub[j * rps:(j + 1) * rps], val[j].data_ptr(), ix[j].data_ptr(), 3, kp, k)
        torch.cuda.synchronize()
        assert torch.equal(i2, idx[j * rps:(j + 1) * rps]) and torch.equal(s2, sims[j * rps:(j + 1) * rps])
    # torch fp32 reference on sampled rows
    sel = torch.randperm(nq * v, device="cuda", generator=gen)[:128]
    ref = torch.topk(rows[sel] @ g.T, k, dim=1)
    same = ref.indices == idx[sel]
    assert float(same.float().mean()) > 0.99
    assert float((ref.values - sims[sel]).abs().max()) <= 1e-3


def test_more_than_2_20_query_rows_are_chunked(tvc_ctx):
    """tvc_search cuts > 2^20 query rows into launches (kMaxRowsPerLaunch): rows either side of the cut
    must equal the same rows searched on their own, for device buffers (asynchronous) and for host
    buffers (the staging workspace is reused by the next chunk), and match the oracle on a sample."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(20)
    m, n, d, k = (1 << 20) + 777, 700, 64, 10
    g = O.l2_normalize(rng.standard_normal((n, d), dtype=np.float32))
    g[400] = g[40]
    q = O.l2_normalize(rng.standard_normal((m, d), dtype=np.float32))
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    probe = np.r_[0:300, (1 << 20) - 300:(1 << 20) + 777]          # first rows, the cut, the ragged tail
    ps, pi = gal.search(q[probe], k)
    rs, ri = O.search(q[probe], g, k)
    assert _band_check(ps, pi, rs, ri) < 0.01
    hs, hi = gal.search(q, k)                                        # host buffers
    assert np.array_equal(hi[probe], pi) and np.array_equal(hs[probe], ps)
    ds, di = gal.search(torch.from_numpy(q).cuda(), k)               # device buffers
    assert np.array_equal(di.cpu().numpy(), hi) and np.array_equal(ds.cpu().numpy(), hs)
    counts = tvc_ctx.k_occurrence(di, n).cpu().numpy()
    assert int(counts.sum()) == m * k and np.array_equal(counts, np.bincount(hi.reshape(-1), minlength=n))


def test_self_search_skips_self_across_the_chunk_cut(tvc_ctx):
    """k-occurrence of a set against itself (README hubness spec, `[:, 1:k+1]`): with more than 2^20
    rows the second launch must exclude row r0+i, not row i.  Planted twins find each other first."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    n, d, k = (1 << 20) + 4096, 64, 4
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    f = torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=gen), dim=1)
    twins = [(5, 900_000), (1 << 20, (1 << 20) + 7), ((1 << 20) + 4000, 123)]
    for a, b in twins:
        f[b] = f[a]
    gal = tvc.Gallery(f, ctx=tvc_ctx)
    sims, idx = gal.search(f, k, skip_self=True)
    rows = torch.arange(n, device="cuda")[:, None]
    assert not bool((idx == rows).any())                             # nobody retrieves itself ...
    assert bool((idx >= 0).all()) and bool((sims[:, :-1] >= sims[:, 1:]).all())
    for a, b in twins:                                               # ... but exact twins retrieve each other
        assert int(idx[a, 0]) == b and int(idx[b, 0]) == a
        assert abs(float(sims[a, 0]) - 1.0) < 1e-5
    # every similarity is the fp32 dot of its index (sample across both launches)
    probe = torch.tensor([0, 5, 524_288, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, n - 1], device="cuda")
    want = (f[probe][:, None, :] * f[idx[probe]]).sum(-1)
    assert float((want - sims[probe]).abs().max()) <= 1e-5
    _, plain_i = gal.search(f[probe], k + 1)                         # without the flag every row finds itself
    for r, p in enumerate(probe.tolist()):                           # (first, or second behind a lower-index twin)
        assert p in plain_i[r, :2].tolist()
