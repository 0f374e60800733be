"""The committed fixtures under tests/golden/ ARE the reference's outputs: where the reference tree is
present (the build container; never the GPU box) re-run tests/golden/make_golden.py — the
reference's own classes, unmodified — into a temp directory and compare with what is committed.
Integer arrays (indices, counts, decisions) must be identical; floating-point arrays may move by
BLAS threading only (1e-6)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")


@pytest.mark.skipif(not (REF / "src" / "retrieval.py").exists(), reason="reference tree not present on this machine")
def test_committed_fixtures_are_the_reference_outputs(tmp_path):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "golden" / "make_golden.py"), "--out", str(tmp_path)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    committed = sorted((ROOT / "tests" / "golden").glob("*.npz"))
    assert committed and sorted(p.name for p in tmp_path.glob("*.npz")) == [p.name for p in committed]
    for p in committed:
        want, got = np.load(p), np.load(tmp_path / p.name)
        assert sorted(want.files) == sorted(got.files), p.name
        for key in want.files:
            a, b = want[key], got[key]
            assert a.shape == b.shape and a.dtype == b.dtype, (p.name, key)
            if a.dtype.kind == "f":
                np.testing.assert_allclose(b, a, rtol=0, atol=1e-6, equal_nan=True, err_msg=f"{p.name}:{key}")
            else:
                assert np.array_equal(a, b), (p.name, key)
