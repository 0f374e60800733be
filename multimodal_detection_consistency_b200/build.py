"""Builds libtvc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the repo snapshot, so the GPU tests and
bench.py load exactly what was cross-compiled here.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libtvc.so"
STAMP = PKG_DIR / ".libtvc.stamp"
SOURCES = ["tvc_gemm_topk.cu", "tvc_gemm_topk_pair.cu", "tvc_gemm_topk_ts.cu", "tvc_aux.cu", "tvc_consistency.cu", "tvc_api.cu"]
HEADERS = ["tvc_ptx.cuh", "tvc_topk.cuh", "tvc_internal.h", "../../include/tvc.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--shared",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libtvc.so")
    return cand


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into libtvc.so unless an up-to-date build exists."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH)] + [str(CSRC / s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout + proc.stderr)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
