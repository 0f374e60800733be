"""Reference-in-the-loop: the oracle against the reference's own classes on fresh random inputs
(tests/golden/live_check.py), in a subprocess.  Runs only where the reference tree exists (the build
container); the GPU box has no /root/reference and skips."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")


@pytest.mark.skipif(not (REF / "src" / "retrieval.py").exists(), reason="reference tree not present on this machine")
def test_oracle_matches_live_reference_on_fresh_seeds():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "golden" / "live_check.py"), "2024", "2025", "2026"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-2500:])
    assert "live reference check ok" in r.stdout
