"""
ctypes binding of the C ABI declared in ``include/emulator_b200.h``.

There is no CPU fallback: if the CUDA shared library is missing or fails to load this
module raises, and every caller above it fails with it.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (EB200_LIB: another build of the same library, e.g. a development variant, for A/B measurements)
LIB_PATH = os.environ.get("EB200_LIB") or os.path.join(_HERE, "libemulator_b200.so")

MAX_DIM = 16


class NativeLibraryError(RuntimeError):
    """The sm_100a CUDA library is missing or broken; there is no other code path."""


def result_len(d: int) -> int:
    return d + 5


_dp = POINTER(c_double)

# name -> (restype, argtypes); mirrors include/emulator_b200.h one to one.
SIGNATURES = {
    "eb200_version": (c_int, []),
    "eb200_last_error": (c_char_p, []),
    "eb200_device_count": (c_int, []),
    "eb200_device_mem_info": (c_int, [c_int, POINTER(ctypes.c_ulonglong), POINTER(ctypes.c_ulonglong)]),
    "eb200_gp_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, _dp, _dp]),
    "eb200_gp_destroy": (c_int, [c_void_p]),
    "eb200_gp_n": (c_int, [c_void_p]),
    "eb200_gp_dim": (c_int, [c_void_p]),
    "eb200_gp_set_residual_host": (c_int, [c_void_p, _dp]),
    "eb200_gp_set_residual_device": (c_int, [c_void_p, c_void_p, c_void_p]),
    "eb200_gp_eval_host": (c_int, [c_void_p, _dp, c_int, _dp]),
    "eb200_gp_eval_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "eb200_gp_eval_multi_host": (c_int, [POINTER(c_void_p), c_int, _dp, c_int, c_int, _dp, c_int]),
    "eb200_gp_eval_multi_device": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "eb200_gp_eval_value_batch_host": (c_int, [c_void_p, c_int, _dp, _dp]),
    "eb200_gp_predict_host": (c_int, [c_void_p, c_int, _dp, c_int, _dp, _dp]),
    "eb200_gp_predict_grid_host": (c_int, [c_void_p, c_int, _dp, c_int, _dp, c_int, c_int, _dp, _dp]),
    "eb200_gp_predict_grid_device": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "eb200_gp_predict_device": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "eb200_gp_synchronize": (c_int, [c_void_p]),
    "eb200_gp_debug_stage": (c_int, [c_void_p, _dp, c_int]),
    "eb200_gp_padded_n": (c_int, [c_void_p]),
    "eb200_gp_debug_matrix": (c_int, [c_void_p, c_int, _dp]),
    "eb200_gp_debug_alpha": (c_int, [c_void_p, _dp]),
    "eb200_gp_eval_launches": (c_int, [c_void_p, c_int]),
    "eb200_gp_debug_factor_trace": (c_int, [c_void_p, _dp, POINTER(ctypes.c_ulonglong), POINTER(c_int)]),
    "eb200_gp_factor_tasks": (c_int, [c_void_p]),
    "eb200_gp_debug_group_trace": (c_int, [POINTER(c_void_p), c_int, _dp, c_int, POINTER(ctypes.c_ulonglong), POINTER(c_int)]),
    "eb200_gp_debug_value_trace": (c_int, [c_void_p, _dp, POINTER(ctypes.c_ulonglong), POINTER(c_int)]),
    "eb200_gp_value_tasks": (c_int, [c_void_p]),
    "eb200_gp_profile_stages": (c_int, [c_void_p, _dp, c_int, POINTER(ctypes.c_float)]),
    "eb200_gp_profile_group_stages": (c_int, [POINTER(c_void_p), c_int, _dp, c_int, c_int, POINTER(ctypes.c_float)]),
    "eb200_bench_accumulate": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "eb200_bench_gemm_nt": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the library (once) and attach the prototypes.  Raises NativeLibraryError."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). emulator_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().eb200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {last_error()}")


def as_double_ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def require_gpu():
    """Fail loudly when no CUDA device is visible (the product has no CPU path)."""
    lib = load()
    if lib.eb200_device_count() < 1:
        raise NativeLibraryError("no CUDA device visible; emulator_b200 has no CPU fallback")
    return lib
