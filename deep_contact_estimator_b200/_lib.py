"""ctypes binding of include/dce.h (libdce_b200.so).

The product path has NO fallback: if the library is missing, or the device is
not a B200, every call raises.  (CPU tensors never reach this module: the
``contact_cnn`` module routes them to the stock PyTorch layers, exactly what
the reference does on a CPU.)
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p, c_float

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libdce_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "dce.h")

DCE_OK = 0
DCE_PREC_FP32 = 0
DCE_PREC_BF16X3 = 1
DCE_NUM_COUNTS = 277
PRECISIONS = {"fp32": DCE_PREC_FP32, "bf16x3": DCE_PREC_BF16X3}

_lib = None


class DceError(RuntimeError):
    def __init__(self, code: int, what: str, cuda_error: int = 0):
        self.code, self.cuda_error = code, cuda_error
        msg = f"{what}: {dce_strerror(code)} (code {code})"
        if cuda_error:
            msg += f", cudaError {cuda_error}"
        super().__init__(msg)


_SIGNATURES = {
    "dce_version": (c_int, []),
    "dce_strerror": (c_char_p, [c_int]),
    "dce_last_cuda_error": (c_int, []),
    "dce_last_launch_count": (c_int, []),
    "dce_weights_create": (c_int, [POINTER(c_void_p), c_int]),
    "dce_weights_destroy": (c_int, [c_void_p]),
    "dce_weights_pack": (c_int, [c_void_p, POINTER(c_void_p), c_void_p]),
    "dce_weights_packed_bytes": (c_size_t, [c_void_p]),
    "dce_weights_packed_ptr": (c_void_p, [c_void_p]),
    "dce_weights_adopt": (c_int, [c_void_p]),
    "dce_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "dce_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "dce_stream": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_size_t, c_int, c_void_p]),
    "dce_forward_profile": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p,
                                    c_int, POINTER(c_float), POINTER(c_char_p), POINTER(c_int)]),
    "dce_stream_profile": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                   c_int, c_void_p, c_int, POINTER(c_float), POINTER(c_char_p), POINTER(c_int)]),
    "dce_latency_server_start": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         ctypes.c_double, c_void_p]),
    "dce_latency_row_server_start": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, ctypes.c_double, c_void_p]),
    "dce_weights_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "dce_debug_read_trace": (c_int, [c_void_p, c_void_p, c_int]),
    "dce_decimal2binary": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "dce_accuracy_counts": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "dce_ingest_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
}


def header_symbols() -> list:
    """Every function include/dce.h declares (used by the CPU symbol test)."""
    with open(HEADER_PATH) as f:
        return sorted(set(re.findall(r"^DCE_API[^;(]*?\b(dce_\w+)\s*\(", f.read(), flags=re.M)))


def load(build_if_missing: bool = False):
    """dlopen libdce_b200.so and type its entry points.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from .build import build_library
            build_library()
        else:
            raise RuntimeError(
                f"{LIB_PATH} not found: the B200 contact-estimator path has no fallback. "
                "Build it with `python -m deep_contact_estimator_b200.build` (needs nvcc).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


TORCH_LIB_PATH = os.path.join(PKG_DIR, "_dce_torch.so")
_torch_ops = False          # False: not tried yet; None: unavailable


def torch_ops():
    """``torch.ops.dce`` (csrc/dce_torch.cpp: forward / stream / accuracy_counts registered with the PyTorch dispatcher
    over the same C ABI), or None when ``_dce_torch.so`` has not been built or ``DCE_BINDING=ctypes`` asks for the
    plain ctypes binding.  Either way the kernels run: the two bindings call the same ``dce_forward`` / ``dce_stream``."""
    global _torch_ops
    if _torch_ops is False:
        _torch_ops = None
        if os.environ.get("DCE_BINDING", "torch") != "ctypes" and os.path.exists(TORCH_LIB_PATH):
            import torch
            load()                                   # libdce_b200.so first: _dce_torch.so links against it
            torch.ops.load_library(TORCH_LIB_PATH)
            _torch_ops = torch.ops.dce
    return _torch_ops


def dce_strerror(code: int) -> str:
    if _lib is None:
        return "?"
    return _lib.dce_strerror(code).decode()


def check(code: int, what: str):
    if code != DCE_OK:
        raise DceError(code, what, _lib.dce_last_cuda_error() if code == -4 else 0)
