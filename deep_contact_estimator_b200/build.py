"""Build the two in-tree shared objects (both cross-compile without a GPU):

    libdce_b200.so   csrc/dce.cu      nvcc, sm_100a: the kernels + the C ABI of include/dce.h (no torch types)
    _dce_torch.so    csrc/dce_torch.cpp   g++: the torch extension (TORCH_LIBRARY `dce`: torch.ops.dce.forward / stream /
                     accuracy_counts) that binds that C ABI to the PyTorch dispatcher; links libdce_b200.so ($ORIGIN rpath)

    python -m deep_contact_estimator_b200.build [--force] [-v]
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


TORCH_LIB_PATH = os.path.join(PKG_DIR, "_dce_torch.so")
TORCH_SOURCES = ["dce_torch.cpp"]


def torch_extension_is_stale() -> bool:
    if not os.path.exists(TORCH_LIB_PATH):
        return True
    t = os.path.getmtime(TORCH_LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in TORCH_SOURCES] + [os.path.join(PKG_DIR, "..", "include", "dce.h"), LIB_PATH]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_torch_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/dce_torch.cpp -> _dce_torch.so against this interpreter's torch headers (no nvcc: it holds no kernels)."""
    build_library(force=False)
    if not force and not torch_extension_is_stale():
        return TORCH_LIB_PATH
    import torch
    from torch.utils import cpp_extension
    # the system g++, NOT $CXX: this image's $CXX (/opt/gcc/bin/g++) links libstdc++ statically, and an extension carrying
    # its own copy of the C++ runtime crashes as soon as a c10::Error (TORCH_CHECK) unwinds through it
    cxx = shutil.which("g++") or os.environ.get("CXX") or "g++"
    cuda_home = os.environ.get("CUDA_HOME") or os.path.dirname(os.path.dirname(_nvcc()))
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += ["-I" + p for p in cpp_extension.include_paths()] + ["-I" + os.path.join(cuda_home, "include")]
    cmd += TORCH_SOURCES + ["-o", TORCH_LIB_PATH]
    cmd += ["-L" + p for p in cpp_extension.library_paths()] + ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch"]
    cmd += ["-L" + PKG_DIR, "-l:libdce_b200.so", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    needed = subprocess.run(["readelf", "-d", TORCH_LIB_PATH], capture_output=True, text=True).stdout
    if needed and "libstdc++.so" not in needed:
        os.remove(TORCH_LIB_PATH)
        raise RuntimeError(f"{cxx} linked libstdc++ statically into _dce_torch.so: exceptions could not cross it; use the system g++")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return TORCH_LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_extension(force="--force" in sys.argv, verbose="-v" in sys.argv))
