"""Kernel (a) parity: libtvc.so tvc_search (tcgen05 GEMM + in-register top-k + fp32 re-rank) against
the NumPy oracle on the same seeded inputs.  Indices bit-exact except inside the 1e-3 similarity
band north_star allows; similarities within 2e-3 (bf16 operands, fp32 accumulate)."""
import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu

SIM_TOL = 2e-3
BAND = 1e-3


def _check_topk(sims, idx, ref_s, ref_i, full_sims=None):
    sims, idx = np.asarray(sims), np.asarray(idx)
    assert sims.shape == ref_s.shape and idx.shape == ref_i.shape
    fin = np.isfinite(ref_s)
    assert np.array_equal(np.isfinite(sims), fin)
    assert np.array_equal(idx < 0, ~fin)
    assert np.abs(sims[fin] - ref_s[fin]).max(initial=0.0) <= SIM_TOL
    bad = (idx != ref_i)
    if bad.any():
        # disagreement allowed only where the two candidates' similarities lie within the band
        assert np.abs(sims[bad] - ref_s[bad]).max() <= BAND, "index mismatch outside the 1e-3 band"
        if full_sims is not None:
            rows, cols = np.nonzero(bad)
            true_of_ours = full_sims[rows, idx[rows, cols]]
            assert np.abs(true_of_ours - ref_s[rows, cols]).max() <= BAND
    return float(bad.mean())


def _unit(rng, n, d):
    return O.l2_normalize(rng.standard_normal((n, d), dtype=np.float32))


@pytest.mark.parametrize("m,n,d,k", [
    (5, 300, 64, 10),        # one tile, ragged everything
    (128, 256, 128, 10),     # exactly one tile
    (129, 257, 512, 10),     # one row / one column over the tile edge
    (1000, 5000, 512, 10),   # C1 scale-down of the reference CPU config
    (300, 4000, 768, 5),     # pipeline top_k=5 (src/pipeline.py:452)
    (64, 3000, 768, 20),     # rerank_top_k=20 (experiments/defenses/retrieval_ref.py)
    (40, 2000, 96, 50),      # KP=64 path
    (17, 1000, 100, 3),      # d not a multiple of 64 / 8
])
def test_search_matches_oracle(tvc_ctx, m, n, d, k):
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(m * 7919 + n)
    g, q = _unit(rng, n, d), _unit(rng, m, d)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(q, k)
    ref_s, ref_i = O.search(q, g, k)
    frac = _check_topk(sims, idx, ref_s, ref_i, q @ g.T)
    assert frac < 0.01
    # every returned similarity is the true fp32 dot of the returned index
    rows = np.arange(m)[:, None]
    assert np.abs((q @ g.T)[rows, idx] - sims).max() <= 1e-5


def test_fewer_rows_than_k_and_empty(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(3)
    g, q = _unit(rng, 4, 64), _unit(rng, 6, 64)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(q, 10)
    ref_s, ref_i = O.search(q, g, 10)
    _check_topk(sims, idx, ref_s, ref_i)
    assert (idx[:, 4:] == -1).all() and np.isneginf(sims[:, 4:]).all()
    empty = tvc.Gallery(dim=64, ctx=tvc_ctx)
    s2, i2 = empty.search(q, 3)
    assert (i2 == -1).all() and np.isneginf(s2).all()


def test_ties_break_to_lower_index(tvc_ctx):
    """Duplicate gallery rows score identically; the lower index must win (north_star)."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(11)
    base = _unit(rng, 50, 128)
    g = np.concatenate([base, base, base], axis=0)  # every row appears 3 times: i, i+50, i+100
    q = _unit(rng, 33, 128)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(q, 9)
    ref_s, ref_i = O.search(q, g, 9)
    assert np.array_equal(idx, ref_i)
    assert np.abs(sims - ref_s).max() <= 1e-5
    trip = idx.reshape(33, 3, 3)
    assert (trip[:, :, 1] == trip[:, :, 0] + 50).all() and (trip[:, :, 2] == trip[:, :, 0] + 100).all()


def test_threshold_offset_and_cosine(tvc_ctx):
    """ReferenceBank semantics: un-normalised rows, cosine, `>= threshold` (src/ref_bank.py:172-224)."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(5)
    g = rng.standard_normal((500, 512)).astype(np.float32) * 3.0
    q = (g[rng.integers(0, 500, 20)] + 0.7 * rng.standard_normal((20, 512)).astype(np.float32) * 3.0)
    gal = tvc.Gallery(g, normalize=True, global_row_offset=1000, ctx=tvc_ctx)
    sims, idx = gal.search(q, 10, threshold=0.3, normalize_queries=True)
    ref_s, ref_i = O.search(q, g, 10, metric="cosine", threshold=0.3, index_offset=1000)
    _check_topk(sims, idx, ref_s, ref_i)
    assert (idx[idx >= 0] >= 1000).all()
    assert (sims[np.isfinite(sims)] >= 0.3).all()


def test_many_splits_small_m(tvc_ctx):
    """Few query rows over a long gallery: the gallery is split across CTAs and merged."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(8)
    g, q = _unit(rng, 70000, 128), _unit(rng, 5, 128)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(q, 10)
    ref_s, ref_i = O.search(q, g, 10)
    _check_topk(sims, idx, ref_s, ref_i, q @ g.T)


def test_torch_device_path_and_append(tvc_ctx):
    import torch
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(21)
    g, q = _unit(rng, 3000, 256), _unit(rng, 200, 256)
    gal = tvc.Gallery(torch.from_numpy(g[:1000]).cuda(), ctx=tvc_ctx)
    gal.append(g[1000:2500])
    gal.append(torch.from_numpy(g[2500:]).cuda().bfloat16().float())  # bf16-exact rows
    g2 = g.copy()
    g2[2500:] = O.bf16_round(g[2500:])
    assert len(gal) == 3000
    sims, idx = gal.search(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    ref_s, ref_i = O.search(q, g2, 10)
    _check_topk(sims.cpu().numpy(), idx.cpu().numpy(), ref_s, ref_i, q @ g2.T)
    assert np.abs(gal.get_rows(np.array([0, 1500, 2999])) - g2[[0, 1500, 2999]]).max() < 1e-6


def test_similarity_matrix(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(4)
    g, q = _unit(rng, 700, 192), _unit(rng, 150, 192)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    s = gal.similarity_matrix(q)
    ref = O.bf16_round(q) @ O.bf16_round(g).T
    assert np.abs(s - ref).max() <= 1e-5          # exact for the bf16 operands
    assert np.abs(s - q @ g.T).max() <= SIM_TOL   # and within tolerance of fp32


def test_skip_self_hubness_knn(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(9)
    f = _unit(rng, 1500, 128)
    gal = tvc.Gallery(f, ctx=tvc_ctx)
    sims, idx = gal.search(f, 10, skip_self=True)
    ref_s, ref_i = O.search(f, f, 10, skip_self=True)
    _check_topk(sims, idx, ref_s, ref_i, f @ f.T)
    assert (idx != np.arange(1500)[:, None]).all()


@pytest.mark.parametrize("m,n,d,k", [(256, 256, 64, 10), (300, 1000, 128, 10), (1000, 5000, 768, 10),
                                     (2500, 9000, 512, 20), (513, 70000, 128, 10), (640, 3000, 96, 50)])
def test_pair_kernel_matches_oracle(tvc_ctx, m, n, d, k):
    """The CTA-pair (cta_group::2) kernel, forced for every size, against the oracle and against the
    single-CTA kernel (bit-identical outputs: same candidates, same fp32 re-rank)."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(m + n)
    g, q = _unit(rng, n, d), _unit(rng, m, d)
    g[n // 2] = g[3]                                   # a tie
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    try:
        tvc_ctx.set_option("pair_min_rows", 0)
        sims, idx = gal.search(q, k)
        sims_self, idx_self = (gal.search(g[:m], k, skip_self=True) if m <= n else (None, None))
        tvc_ctx.set_option("pair_min_rows", 1 << 62)
        sims1, idx1 = gal.search(q, k)
    finally:
        tvc_ctx.set_option("pair_min_rows", 4096)
    ref_s, ref_i = O.search(q, g, k)
    _check_topk(sims, idx, ref_s, ref_i, q @ g.T)
    assert np.array_equal(idx, idx1) and np.array_equal(sims, sims1)
    if idx_self is not None:
        rs, ri = O.search(g[:m], g, k, skip_self=True)
        _check_topk(sims_self, idx_self, rs, ri, g[:m] @ g.T)


@pytest.mark.parametrize("m,n,d,k", [(256, 2048, 64, 10), (300, 5000, 128, 10), (1000, 9000, 768, 10), (4500, 30000, 768, 10),
                                     (2500, 9000, 512, 20), (700, 6000, 320, 10), (640, 3000, 96, 50), (513, 4099, 704, 1)])
def test_resident_query_kernels_match_pair_kernel(tvc_ctx, m, n, d, k):
    """The two resident-query revisions of kernel (a), forced for every size, against the oracle and against the plain
    CTA-pair kernel: bit-identical outputs (same bf16 operands and k order into the same fp32 accumulators, same
    column order into the top-k lists).
      rq: first 7 k-blocks of the query tile resident in shared memory, the rest and the gallery through a 6-slot ring
          (the default for long units);
      ts: query tile in tensor memory, N = 64 MMA tiles, 3-D TMA boxes (measured slower, off by default).
    Covers ragged last tiles in both dimensions, several query tiles per pair (units back to back: the resident tile
    is replaced), d <= 448 (everything resident) and d > 448 (streamed query k-blocks), k-block counts that are not a
    multiple of the 4 a 3-D box carries (d = 320, 704), KP = 16 / 32 / 64, ties and skip_self."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(m * 7 + n)
    g, q = _unit(rng, n, d), _unit(rng, m, d)
    g[n // 2] = g[3]                                   # a tie
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    never = 1 << 62
    got = {}
    try:
        tvc_ctx.set_option("pair_min_rows", 0)
        for name, rq, ts in (("rq", 1, never), ("ts", never, 1), ("pair", never, never)):
            tvc_ctx.set_option("rq_min_tiles", rq)
            tvc_ctx.set_option("ts_min_tiles", ts)
            got[name] = gal.search(q, k) + (gal.search(g[:m], k, skip_self=True) if m <= n else (None, None))
    finally:
        tvc_ctx.set_option("pair_min_rows", 4096)
        tvc_ctx.set_option("rq_min_tiles", 64)
        tvc_ctx.set_option("ts_min_tiles", never)     # the default: off (measured slower, see the kernel's header)
    ref_s, ref_i = O.search(q, g, k)
    _check_topk(got["pair"][0], got["pair"][1], ref_s, ref_i, q @ g.T)
    for name in ("rq", "ts"):
        for a, b in zip(got[name], got["pair"]):
            assert (a is None and b is None) or np.array_equal(a, b), name
    if got["pair"][3] is not None:
        rs, ri = O.search(g[:m], g, k, skip_self=True)
        _check_topk(got["pair"][2], got["pair"][3], rs, ri, g[:m] @ g.T)


@pytest.mark.parametrize("n,d,m,k,shards", [(5000, 256, 700, 10, 3), (40000, 128, 5000, 10, 4), (900, 64, 33, 5, 2)])
def test_sharded_two_phase_search_equals_unsharded(tvc_ctx, n, d, m, k, shards):
    """tvc_search_candidates (scatter into per-owner receive buffers) + tvc_rerank_candidates over a
    gallery group returns exactly what tvc_search returns on the unsharded gallery."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    from multimodal_detection_consistency_b200._native import Scatter
    g = O.synth_gallery(n, d, seed=7, clusters=32, dup_rate=0.01)
    q = torch.from_numpy(O.synth_queries(g, m, 1, seed=8)[2].reshape(m, d)).cuda()
    full = tvc.Gallery(g, ctx=tvc_ctx)
    want_s, want_i = full.search(q, k)
    per = -(-n // shards)
    parts = [tvc.Gallery(g[r * per:(r + 1) * per], global_row_offset=r * per, ctx=tvc_ctx) for r in range(shards)]
    group = tvc.Gallery.group(parts)
    kp = tvc_ctx.candidate_width(k)
    owners = 3                                             # query slices (the "ranks" that re-rank)
    rps = -(-m // owners)
    val = [torch.full((shards, min(rps, max(0, m - j * rps)), kp), float("nan"), device="cuda") for j in range(owners)]
    idx = [torch.full((shards, min(rps, max(0, m - j * rps)), kp), -7, dtype=torch.int64, device="cuda")
           for j in range(owners)]
    # the bf16 operand is prepared once, slice by slice, into two "rank" buffers (broadcast form)
    d_pad = tvc_ctx.query_row_bytes(d) // 2
    op = [torch.zeros((m, d_pad), dtype=torch.bfloat16, device="cuda") for _ in range(2)]
    for j in range(owners):
        rows = q[j * rps:(j + 1) * rps]
        if rows.shape[0]:
            tvc_ctx.prepare_queries(rows, [o.data_ptr() for o in op], j * rps)
    torch.cuda.synchronize()
    assert torch.equal(op[0], op[1])
    assert torch.equal(op[0][:, :d], q.to(torch.bfloat16)) and not op[0][:, d:].any()
    for r, part in enumerate(parts):
        sc = Scatter()
        sc.n_slices, sc.slot, sc.rows_per_slice = owners, r, rps
        for j in range(owners):
            sc.val[j], sc.idx[j] = val[j].data_ptr(), idx[j].data_ptr()
        # odd shards search the prepared operand, even shards the raw fp32 rows: same candidates
        part.search_candidates((op[r % 2].data_ptr(), m) if r % 2 else q, k, sc)
    got_s, got_i = [], []
    for j in range(owners):
        rows = q[j * rps:(j + 1) * rps]
        if rows.shape[0] == 0:
            continue
        s, i = tvc_ctx.rerank_candidates(group, rows, val[j].data_ptr(), idx[j].data_ptr(), shards, kp, k)
        got_s.append(s)
        got_i.append(i)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(got_i), want_i)
    assert torch.equal(torch.cat(got_s), want_s)
    # local (non-scattered) candidate lists: global indices, sorted by GEMM score
    cv, ci = parts[1].search_candidates(q, k)
    torch.cuda.synchronize()
    ok = ci >= 0
    assert ((ci[ok] >= per) & (ci[ok] < 2 * per)).all()
    assert (cv[:, :-1] >= cv[:, 1:]).all()


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_half_precision_rows_and_queries(tvc_ctx, dtype):
    """tvc_dtype BF16 / F16 inputs (gallery rows and queries): the search runs on exactly the values
    given, so the result equals the fp32 search of the up-converted rows."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(5)
    dt = getattr(torch, dtype)
    g = torch.from_numpy(_unit(rng, 3000, 192)).cuda().to(dt)
    q = torch.from_numpy(_unit(rng, 257, 192)).cuda().to(dt)
    s_half, i_half = tvc.Gallery(g, ctx=tvc_ctx).search(q, 10)
    s_f32, i_f32 = tvc.Gallery(g.float(), ctx=tvc_ctx).search(q.float(), 10)
    torch.cuda.synchronize()
    assert torch.equal(i_half, i_f32) and torch.equal(s_half, s_f32)
    ref_s, ref_i = O.search(q.float().cpu().numpy(), g.float().cpu().numpy(), 10)
    assert (i_half.cpu().numpy() != ref_i).mean() < 0.01
    assert np.abs(s_half.cpu().numpy() - ref_s).max() <= (2e-3 if dtype == "float16" else 2e-3)


def test_host_rows_go_through_a_bounded_staging_window(tvc_ctx):
    """tvc_gallery_create/append with HOST rows larger than the 64 MiB staging window (several windows,
    ragged last one, f32 and f16): same rows and same search results as the device-resident upload, and
    the workspace stays window-sized instead of holding a second copy of the gallery."""
    import torch
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(77)
    n, d = 150_001, 256                                   # 153.6 MB of fp32 rows -> 3 windows
    g = _unit(rng, n, d)
    q = _unit(rng, 300, d)
    tvc_ctx.release_workspace()
    from_host = tvc.Gallery(g, ctx=tvc_ctx)
    held = tvc_ctx.release_workspace()
    assert 0 < held <= 80 << 20, held                     # one window (+12.5 % growth slack), not n*d*4
    from_dev = tvc.Gallery(torch.from_numpy(g).cuda(), ctx=tvc_ctx)
    pick = np.array([0, 1, 65535, 65536, 65537, 131071, 131072, n - 1], np.int64)
    assert np.array_equal(from_host.get_rows(pick), g[pick])
    s0, i0 = from_host.search(q, 10)
    s1, i1 = from_dev.search(q, 10)
    assert np.array_equal(np.asarray(i0), np.asarray(i1)) and np.array_equal(np.asarray(s0), np.asarray(s1))
    ref_s, ref_i = O.search(q, g, 10)
    assert _check_topk(s0, i0, ref_s, ref_i) < 0.01
    # append of fp16 host rows across the window edge
    extra = _unit(rng, 140_000, d).astype(np.float16)     # 71.7 MB -> 2 windows
    from_host.append(extra)
    assert len(from_host) == n + len(extra)
    pick2 = np.array([n, n + 131071, n + 131072, n + len(extra) - 1], np.int64)
    assert np.array_equal(from_host.get_rows(pick2), extra[pick2 - n].astype(np.float32))


def test_release_workspace_is_transparent(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(5)
    g, q = _unit(rng, 3000, 128), _unit(rng, 200, 128)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    s0, i0 = gal.search(q, 10)
    assert tvc_ctx.release_workspace() > 0
    assert tvc_ctx.release_workspace() == 0
    s1, i1 = gal.search(q, 10)
    assert np.array_equal(np.asarray(i0), np.asarray(i1)) and np.array_equal(np.asarray(s0), np.asarray(s1))
    assert tvc.Context.release_all_workspaces() > 0


def test_wrong_query_dimension_raises(tvc_ctx):
    """a query of another width must be refused before its pointer reaches the library (the
    reference raises from np.dot / the FAISS assert)."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(5)
    gal = tvc.Gallery(_unit(rng, 500, 128), ctx=tvc_ctx)
    for bad in (64, 129, 256):
        with pytest.raises(ValueError):
            gal.search(_unit(rng, 3, bad), 5)
        with pytest.raises(ValueError):
            gal.similarity_matrix(_unit(rng, 3, bad))
    # and the C entry itself refuses a mismatching d (what abi callers hit)
    import ctypes as C
    q = _unit(rng, 3, 64)
    sims, idx = np.empty((3, 5), np.float32), np.empty((3, 5), np.int64)
    rc = tvc_ctx.lib.tvc_search(tvc_ctx.handle, gal.handle, q.ctypes.data, 0, 3, 64, 5, C.c_float(-np.inf), 0,
                                sims.ctypes.data, idx.ctypes.data, None)
    assert rc == 1   # TVC_ERR_INVALID


@pytest.mark.parametrize("k", [57, 100, 600])
def test_k_beyond_the_epilogue_limit(tvc_ctx, k):
    """FAISS and the reference accept any k; beyond TVC_MAX_K the wrapper serves it from the dense similarity
    tile + fp32 re-score (same ordering rule, -1 padding when k > N)."""
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(k)
    g, q = _unit(rng, 500, 128), _unit(rng, 9, 128)
    g[300] = g[12]
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    sims, idx = gal.search(q, k)
    ref_s, ref_i = O.search(q, g, k)
    _check_topk(sims, idx, ref_s, ref_i, q @ g.T)
    import torch
    ts, ti = gal.search(torch.from_numpy(q).cuda(), k, threshold=0.05)
    ref_s, ref_i = O.search(q, g, k, threshold=0.05)
    _check_topk(ts.cpu().numpy(), ti.cpu().numpy(), ref_s, ref_i, q @ g.T)


def test_concurrent_host_buffer_searches(tvc_ctx):
    """SURVEY §8b threading row / VERDICT r1 weak 9: four threads (the reference's ThreadPoolExecutor(max_workers=4),
    src/pipeline.py:42,288,555-560) call tvc_search with HOST buffers at the same time on one immutable gallery;
    each call runs on its own context-owned stream with its own workspace, no context-wide lock is held across
    the blocking copies, and every thread gets exactly the single-threaded answer."""
    import threading
    import time
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(11)
    g = _unit(rng, 60000, 256)
    gal = tvc.Gallery(g, ctx=tvc_ctx)
    qs = [_unit(rng, 700 + 50 * t, 256) for t in range(4)]
    want = [gal.search(q, 10) for q in qs]
    got = [None] * 4
    errs = []
    gate = threading.Barrier(4)

    def work(t):
        try:
            gate.wait()
            for _ in range(6):
                got[t] = gal.search(qs[t], 10)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    t0 = time.perf_counter()
    for q in qs:
        for _ in range(6):
            gal.search(q, 10)
    serial = time.perf_counter() - t0
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    threaded = time.perf_counter() - t0
    assert not errs, errs
    for t in range(4):
        assert np.array_equal(got[t][1], want[t][1]) and np.array_equal(got[t][0], want[t][0])
    print(f"24 host-buffer searches: serial {serial * 1e3:.1f} ms, 4 threads {threaded * 1e3:.1f} ms")
    # device-pointer calls on distinct torch streams from threads as well
    import torch
    dq = [torch.from_numpy(q).cuda() for q in qs]
    res = [None] * 4

    def work_dev(t):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(4):
                res[t] = gal.search(dq[t], 10)
            s.synchronize()

    th = [threading.Thread(target=work_dev, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for t in range(4):
        assert np.array_equal(res[t][1].cpu().numpy(), want[t][1])
