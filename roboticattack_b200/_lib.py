"""ctypes loader for the C-ABI library. There is no CPU fallback: if the library is missing or a call fails,
the caller gets an exception."""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_longlong, c_void_p
from pathlib import Path

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "libvla_b200.so"


class VLAError(RuntimeError):
    pass


def _declare(lib):
    lib.vla_last_error.restype = c_char_p
    lib.vla_last_error.argtypes = []
    lib.vla_abi_version.restype = c_int
    lib.vla_launch_count.restype = c_longlong
    lib.vla_gemm_bf16_tn.restype = c_int
    lib.vla_gemm_bf16_tn.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise VLAError(
                f"{LIB_PATH} is missing: build it with `python -m roboticattack_b200.build` "
                "(there is no CPU fallback for the attack hot path)")
        handle = ctypes.CDLL(str(LIB_PATH))
        _declare(handle)
        _LIB = handle
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().vla_last_error().decode("utf-8", "replace")
        raise VLAError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
