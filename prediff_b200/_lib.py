"""ctypes binding of libprediff_b200.so (the C ABI in include/prediff_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PD_LIB_PATH: another build of the same library (A/B measurements of two builds on one box, tools/ab.sh)
LIB_PATH = os.environ.get("PD_LIB_PATH") or os.path.join(_HERE, "libprediff_b200.so")

_lib = None


class PDError(RuntimeError):
    pass


def lib():
    """Returns the loaded CDLL (loading it on first use)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PDError(
                f"{LIB_PATH} not found: build it with `python -m prediff_b200.build` "
                "(prediff_b200 has no CPU / PyTorch fallback path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pd_last_error.restype = ctypes.c_char_p
        _lib.pd_version.restype = ctypes.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().pd_last_error().decode("utf-8", "replace")
        raise PDError(f"prediff_b200 error {rc}: {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None as a c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_contiguous(), "prediff_b200 needs contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    """cudaStream_t of torch's current stream as a c_void_p."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def init():
    check(lib().pd_init())


c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_float = ctypes.c_float
c_double = ctypes.c_double
