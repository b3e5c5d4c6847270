"""Build libdce_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m deep_contact_estimator_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdce_b200.so")
SOURCES = ["dce.cu"]
HEADERS = ["dce_common.cuh", "dce_fp32.cuh", "dce_tc.cuh", "dce_tc_ptx.cuh", "dce_tc_block1.cuh", "dce_tc_block2.cuh", "dce_tc_run.cuh", "dce_small.cuh", "dce_latency.cuh", os.path.join("..", "..", "include", "dce.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libdce_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/dce.cu -> libdce_b200.so; returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    extra = ["-DDCE_TRACE=1"] if os.environ.get("DCE_TRACE") == "1" else []      # clock64 timelines (tools/trace_*.py)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
