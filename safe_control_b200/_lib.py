"""Loader for libscb.so -- the hand-written sm_100a kernels behind include/scb.h.

There is NO CPU fallback: if the library is missing (not built) this raises, and every
solve raises if no CUDA device is present.  Build with `python -m safe_control_b200.build`
(or `__graft_entry__.build()`).
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCB_LIB selects another build of the same library (tuning variants, profiling builds); never a different backend
LIB_PATH = os.environ.get("SCB_LIB") or os.path.join(_HERE, "libscb.so")
_lib = None


class ScbError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScbError(
                f"{LIB_PATH} not found: the CUDA extension is not built. "
                "Run `python -m safe_control_b200.build` (needs nvcc). There is no CPU fallback.")
        _lib = _abi.bind(C.CDLL(LIB_PATH))
    return _lib


def check(rc, what="scb call"):
    if rc != 0:
        l = lib()
        msg = l.scb_strerror(rc).decode()
        if rc == -4:
            msg += f" [cudaError {l.scb_last_cuda_error()}]"
        raise ScbError(f"{what} failed: {msg} (code {rc})")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise ScbError("no CUDA device: safe_control_b200 has no CPU path (the numpy oracle under oracle/ is test-only)")
