"""ctypes binding of libdfine_b200.so (the C-ABI declared in include/dfine_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception.  Nothing here imports or calls the CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
from typing import Optional

from . import build as _build

F32, BF16 = 0, 1
MSDA_FUSED_INPUTS = 1
MSDA_GRAD_VALUE_BF16 = 2
MSDA_FORCE_ATOMIC = 4
MSDA_GRAD_SAMP_BF16 = 8
MSDA_RECORDS_VALID = 16
MSDA_GRAD_VALUE_ACCUMULATE = 32
MSDA_BWD_DOTS_ONLY = 128
MSDA_BWD_VALUE_ONLY = 256
E_NULL, E_SHAPE, E_UNSUPPORTED, E_ALIGN = -1, -2, -3, -4

_I32P = ctypes.POINTER(c_int32)
_lib: Optional[ctypes.CDLL] = None

_SIGNATURES = {
    "dfine_version": (c_int, []),
    "dfine_last_error": (c_char_p, []),
    "dfine_msda_fwd": (c_int, [c_void_p, c_int64, c_int64, _I32P, _I32P, _I32P, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int64,
                               c_int64, c_void_p, c_void_p]),
    "dfine_msda_bwd": (c_int, [c_void_p, c_int64, c_int64, _I32P, _I32P, _I32P, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
                               c_void_p]),
    "dfine_msda_bwd_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int]),
    "dfine_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "dfine_colsum": (c_int, [c_void_p, c_int, c_int64, c_int, c_int64, c_void_p, c_void_p]),
    "dfine_fdr_project": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dfine_fdr_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int64, c_int, c_void_p]),
    "dfine_fdr_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_int64, c_int, c_void_p]),
    "dfine_linear_wgrad": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "dfine_linear_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                 c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "dfine_gate_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_float, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "dfine_ffn_out_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                                  c_float, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "dfine_ffn_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float,
                              c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "dfine_lqe_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dfine_multicast_add": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "dfine_pack_linear": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                  c_void_p, c_int, c_void_p]),
    "dfine_mask_gemm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p]),
    "dfine_mask_loss_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "dfine_mask_loss_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_int, c_void_p]),
    "dfine_lsap": (c_int, [c_void_p, c_int64, c_int64, c_int64, _I32P, c_int, c_int, c_void_p, c_void_p, c_int64,
                           c_void_p]),
    "dfine_mask_gemm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p]),
}


class DfineB200Error(RuntimeError):
    """A C-ABI call returned non-zero (message from dfine_last_error())."""


def library_path() -> str:
    return _build.LIB_PATH


def lib() -> ctypes.CDLL:
    """Load libdfine_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise DfineB200Error(
                f"{path} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for this path.")
        handle = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dfine_last_error().decode("utf-8", "replace")
        raise DfineB200Error(f"{what or 'libdfine_b200'} failed (code {rc}): {msg}")


def exported_symbols():
    return sorted(_SIGNATURES)


def i32_array(values):
    arr = (c_int32 * len(values))(*[int(v) for v in values])
    return arr
