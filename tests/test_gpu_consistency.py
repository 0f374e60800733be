"""Kernel (b) parity: tvc_consistency_sims / tvc_consistency_emb against the oracle and against the
reference-generated golden fixtures.  Scores within 2e-3 (north_star; observed ~1e-6), decisions
identical outside a 1e-5 exclusion zone around the thresholds."""
from pathlib import Path

import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
TOL = 2e-3
TIGHT = 5e-6


def _decisions_match(flags, ref_flags, ref_scores, params):
    p = dict(O.DEFAULT_PARAMS)
    p.update(params or {})
    flags, ref_flags = np.asarray(flags), np.asarray(ref_flags)
    ok_det = np.abs(ref_scores[:, O.S_DET_AGG] - p["detection_threshold"]) > 1e-5
    ok_cc = np.abs(ref_scores[:, O.S_CC_OVERALL] - ref_scores[:, O.S_CC_THRESHOLD]) > 1e-5
    ok_cc &= np.abs(ref_scores[:, O.S_CROSS_MODAL_VAR] - 0.1) > 1e-6
    ok_sig = np.abs(ref_scores[:, O.S_REF_SIGMA] - p["sigma_threshold"]) > 1e-5
    for bit, ok in [(O.FLAG_DET_ADV, ok_det), (O.FLAG_CC_ADV, ok_cc), (O.FLAG_SIGMA_ADV, ok_sig)]:
        assert np.array_equal((flags & bit)[ok], (ref_flags & bit)[ok])


@pytest.mark.parametrize("q,V,R,G", [(1, 5, 10, 3), (127, 5, 10, 3), (1000, 5, 10, 3), (513, 3, 4, 2), (300, 16, 16, 16)])
@pytest.mark.parametrize("voting,agg", [(1, 0), (0, 1), (2, 2), (1, 3)])
def test_sims_mode(tvc_ctx, q, V, R, G, voting, agg):
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(q * 31 + V)
    s0 = rng.uniform(-0.2, 0.9, q).astype(np.float32)
    sv = (s0[:, None] + rng.normal(0, 0.15, (q, V))).astype(np.float32)
    sr = rng.uniform(-0.1, 0.9, (q, R)).astype(np.float32)
    sg = rng.uniform(-0.1, 0.9, (q, G)).astype(np.float32)
    r_cnt = rng.integers(0, R + 1, q).astype(np.int32)
    g_cnt = rng.integers(0, G + 1, q).astype(np.int32)
    sxv = rng.uniform(0.5, 1.0, (q, V * (V - 1) // 2)).astype(np.float32)
    over = dict(n_variants=V, n_retrieval=R, n_generative=G, voting=voting, aggregation=agg)
    params = tvc.default_params(**over)
    scores, flags = tvc_ctx.consistency_sims(params, s0, sv, sr, r_cnt, sg, g_cnt, sxv)
    ref, rflags = O.consistency_sims(s0, sv, sr, r_cnt, sg, g_cnt, sxv, over)
    assert np.abs(scores - ref).max() <= TIGHT
    _decisions_match(flags, rflags, ref, over)


def test_sims_mode_device_tensors_and_optional_inputs(tvc_ctx):
    import torch
    import multimodal_detection_consistency_b200 as tvc
    rng = np.random.default_rng(1)
    q = 700
    s0 = rng.uniform(0, 0.9, q).astype(np.float32)
    sv = rng.uniform(0, 0.9, (q, 5)).astype(np.float32)
    params = tvc.default_params(n_retrieval=0, n_generative=0)
    sc, fl = tvc_ctx.consistency_sims(params, torch.from_numpy(s0).cuda(), torch.from_numpy(sv).cuda())
    torch.cuda.synchronize()
    ref, rfl = O.consistency_sims(s0, sv, params=dict(n_retrieval=0, n_generative=0))
    assert np.abs(sc.cpu().numpy() - ref).max() <= TIGHT
    _decisions_match(fl.cpu().numpy(), rfl, ref, {})


def test_emb_mode_against_oracle(tvc_ctx):
    import multimodal_detection_consistency_b200 as tvc
    d, nq, V, k = 768, 333, 5, 10
    g = O.synth_gallery(4000, d, seed=1, clusters=64)
    bank = O.synth_gallery(900, d, seed=2, clusters=64)
    img, txt, var = O.synth_queries(g, nq, V, seed=3)
    gal, bnk = tvc.Gallery(g, ctx=tvc_ctx), tvc.Gallery(bank, global_row_offset=50, ctx=tvc_ctx)
    _, ridx = gal.search(var, k)
    _, gidx = bnk.search(var, k)
    params = tvc.default_params()
    scores, flags, (sv, sr, sg) = tvc_ctx.consistency_emb(params, img, txt, var, ret_gallery=gal,
                                                          ret_idx=ridx.reshape(nq, V * k), gen_gallery=bnk,
                                                          gen_idx=gidx.reshape(nq, V * k), return_sims=True)
    ref, rflags, (rsv, rsr, rsg) = O.consistency_emb(img, txt, var, ret_rows=g, ret_idx=ridx.reshape(nq, V * k),
                                                     gen_rows=bank, gen_idx=gidx.reshape(nq, V * k), gen_offset=50)
    assert np.array_equal(scores[:, O.S_N_RET], ref[:, O.S_N_RET])
    assert np.array_equal(scores[:, O.S_N_GEN], ref[:, O.S_N_GEN])
    assert np.abs(scores - ref).max() <= TIGHT
    for i in range(nq):
        assert np.abs(sr[i, :len(rsr[i])] - np.array(rsr[i], np.float32)).max(initial=0) <= TIGHT
        assert np.abs(sg[i, :len(rsg[i])] - np.array(rsg[i], np.float32)).max(initial=0) <= TIGHT
        assert np.abs(sv[i] - np.array(rsv[i], np.float32)).max() <= TIGHT
    _decisions_match(flags, rflags, ref, {})


def test_emb_mode_golden_reference_detectors(tvc_ctx):
    """Same inputs the reference's AdversarialDetector / MultiModalDefenseDetector were run on."""
    import multimodal_detection_consistency_b200 as tvc
    z = np.load(GOLD / "detectors.npz")
    gal = tvc.Gallery(z["gallery"], ctx=tvc_ctx)
    for mode_i in range(4):
        params = tvc.default_params(aggregation=mode_i)
        scores, flags = tvc_ctx.consistency_emb(params, z["img"], z["txt"], z["var"], ret_gallery=gal,
                                                ret_idx=z["cand"], gen=z["gen"], g_cnt=z["g_cnt"])
        want = z["det_scores"][mode_i]
        assert np.abs(scores[:, O.S_DET_TV] - want[:, 0]).max() <= TOL
        assert np.abs(scores[:, O.S_DET_AGG] - want[:, 3]).max() <= TOL
        assert np.abs(scores[:, O.S_DET_AGG] - want[:, 3]).max() <= 1e-5
        margin = np.abs(want[:, 3] - 0.5) > 1e-5
        assert np.array_equal((flags & O.FLAG_DET_ADV).astype(bool)[margin], want[margin, 4].astype(bool))
    keys = [str(k) for k in z["cs_keys"]]
    col = dict(original_similarity=O.S_ORIGINAL, text_variant_consistency=O.S_TV_MEAN, text_variant_std=O.S_TV_STD,
               retrieval_consistency=O.S_RET_MEAN, retrieval_std=O.S_RET_STD, generative_consistency=O.S_GEN_MEAN,
               generative_std=O.S_GEN_STD, cross_modal_variance=O.S_CROSS_MODAL_VAR)
    for j, kname in enumerate(keys):
        assert np.abs(scores[:, col[kname]] - z["cs"][:, j]).max() <= 1e-5, kname
    assert np.array_equal(scores[:, O.S_N_RET].astype(np.int64), z["n_ret"])


def test_checker_golden(tvc_ctx):
    """ConsistencyChecker.make_decision outputs of the reference, fed as degenerate similarity lists."""
    import multimodal_detection_consistency_b200 as tvc
    z = np.load(GOLD / "consistency_checker.npz")
    keys = [str(k) for k in z["keys"]]
    c = {k: i for i, k in enumerate(keys)}
    S = z["scores"]
    n = S.shape[0]

    def pair(mean, std):  # two similarities with exactly this mean and population std
        return np.stack([mean - std, mean + std], 1).astype(np.float32)

    s0 = S[:, c["original_similarity"]].astype(np.float32)
    sv = pair(S[:, c["text_variant_consistency"]], S[:, c["text_variant_std"]])
    sr = pair(S[:, c["retrieval_consistency"]], S[:, c["retrieval_std"]])
    sg = pair(S[:, c["generative_consistency"]], S[:, c["generative_std"]])
    r_cnt = np.where(S[:, c["retrieval_consistency"]] == 0, 0, 2).astype(np.int32)
    g_cnt = np.where(S[:, c["generative_consistency"]] == 0, 0, 2).astype(np.int32)
    for voting, vi in [("simple", 0), ("weighted", 1), ("adaptive", 2)]:
        params = tvc.default_params(n_variants=2, n_retrieval=2, n_generative=2, voting=vi, cc_adaptive=0)
        scores, flags = tvc_ctx.consistency_sims(params, s0, sv, sr, r_cnt, sg, g_cnt)
        want = z[f"{voting}_0"]
        # the kernel derives cross_modal_variance itself; the fixture drew it at random, so compare
        # only what does not depend on it when adaptive thresholds are off: overall score + decision
        assert np.abs(scores[:, O.S_CC_OVERALL] - want[:, 0]).max() <= 1e-5
        margin = np.abs(want[:, 0] - 0.5) > 1e-5
        assert np.array_equal((flags & O.FLAG_CC_ADV).astype(bool)[margin], want[margin, 3].astype(bool))
