"""ctypes binding of libdensebox_b200.so (the C ABI in include/densebox_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, we raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DBX_LIB: A/B measurements of two builds in one process run (tools/ only); the default is the in-tree build
LIB_PATH = os.environ.get("DBX_LIB") or os.path.join(_HERE, "csrc", "libdensebox_b200.so")
_lib = None


class DbxError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises DbxError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DbxError(
                "densebox_b200: %s is missing — run `python -m densebox_b200.build` "
                "(there is no CPU fallback for the CUDA path)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dbx_error_string.restype = ctypes.c_char_p
        _lib.dbx_error_string.argtypes = [ctypes.c_int]
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise DbxError("densebox_b200 %s failed: %s (code %d)" % (what, lib().dbx_error_string(rc).decode(), rc))


def ptr(t):
    """Device (or host) pointer of a torch tensor / None as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
