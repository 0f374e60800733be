"""The Python mirrors (ReferenceBank, ConsistencyChecker, hubness, RetrievalEvaluator, retriever, detector)
next to the reference's own classes on the same random scenarios (tests/golden/mirrors_live.py, subprocess).
Build container only: needs /root/reference; native layer = oracle-backed test double (tests/fake_native.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")


@pytest.mark.skipif(not (REF / "src" / "ref_bank.py").exists(), reason="reference tree not present on this machine")
def test_mirrors_behave_like_the_reference_classes():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "golden" / "mirrors_live.py"), "31", "32"],
                       capture_output=True, text=True, timeout=1200, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "mirrors live check ok" in r.stdout
