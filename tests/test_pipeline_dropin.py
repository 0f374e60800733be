"""The reference's own orchestrator (src/pipeline.py, unmodified) run with this repo's mirrors swapped in
for its two imports, next to a run with the reference's own classes, on the same table encoders
(tests/golden/pipeline_dropin.py, in a subprocess).  Build container only: needs /root/reference; the
native layer is the oracle-backed test double there (no GPU), so this pins the HOST side of the drop-in."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")


@pytest.mark.skipif(not (REF / "src" / "pipeline.py").exists(), reason="reference tree not present on this machine")
def test_reference_pipeline_runs_unchanged_on_the_mirrors():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "golden" / "pipeline_dropin.py"), "7", "8"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "pipeline drop-in ok" in r.stdout
