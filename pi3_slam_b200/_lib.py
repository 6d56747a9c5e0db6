"""ctypes binding of libpi3b200.so — the C-ABI boundary declared in include/pi3b200.h.

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
PyTorch is only used by callers for device memory and streams; nothing torch-typed crosses this file.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libpi3b200.so"
_lib = None

# name -> (restype, argtypes); mirrors include/pi3b200.h one to one.
_P = c_void_p
SIGNATURES = {
    "pi3b_version": (c_int, []),
    "pi3b_last_error": (c_char_p, []),
    "pi3b_gemm": (c_int, [_P, c_int, _P, c_int64, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_int64, _P,
                          c_int64, c_int, c_int]),
    "pi3b_attn_fwd_hd64": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, _P, c_int64,
                                   c_int, c_int, c_int, c_float]),
}


class Pi3bError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load(build_if_missing: bool = False) -> ctypes.CDLL:
    """Load libpi3b200.so (optionally compiling it first). Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from .build import build_lib
        build_lib()
    if not _LIB_PATH.exists():
        raise Pi3bError(
            f"{_LIB_PATH} not found: build it with `python -m pi3_slam_b200.build` "
            "(there is no CPU or PyTorch fallback for the pi3 chunk path)")
    try:
        import torch  # noqa: F401  (brings libcudart.so.12 into the process before dlopen)
    except Exception:  # pragma: no cover
        pass
    lib = ctypes.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pi3b_last_error()
        raise Pi3bError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
