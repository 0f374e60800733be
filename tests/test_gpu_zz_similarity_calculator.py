"""SimilarityCalculator mirror (src/utils/metrics.py:107-164) on the GPU: the batched cosine matrix is
tvc_similarity_matrix (bf16 operands, fp32 accumulate) - within 2e-3 of the fp32 reference arithmetic."""
import numpy as np
import pytest

from oracle import tvc_oracle as O

pytestmark = pytest.mark.gpu


def test_batch_cosine_similarity_matches_oracle(tvc_ctx):
    from multimodal_detection_consistency_b200.metrics import SimilarityCalculator
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((300, 512)) * 4).astype(np.float32)          # un-normalised, as callers pass them
    y = (rng.standard_normal((700, 512)) * 0.3).astype(np.float32)
    got = SimilarityCalculator.batch_cosine_similarity(x, y)
    want = O.similarity_matrix(x, y, "cosine")
    assert got.shape == (300, 700) and np.abs(got - want).max() <= 2e-3
    import torch
    got_t = SimilarityCalculator.batch_cosine_similarity(torch.from_numpy(x), torch.from_numpy(y).cuda())
    assert np.array_equal(got_t, got)
    assert abs(SimilarityCalculator.cosine_similarity(x[0], y[0]) - O.scalar_cosine(x[0], y[0])) <= 1e-6
    assert SimilarityCalculator.cosine_similarity(np.zeros(8), np.ones(8)) == 0.0
