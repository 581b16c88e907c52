"""ctypes binding of the C ABI in include/monopsr_b200_*.h (libmonopsr_b200.so).

Fails loudly: a missing library is an ImportError-like RuntimeError, a non-zero status
from any entry point raises ``MpbError`` -- there is no fallback path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPB_LIB", os.path.join(_HERE, "libmonopsr_b200.so"))

c_f = ctypes.c_void_p   # device pointers travel as integers
c_i = ctypes.c_int


class MpbError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "mpb_nn_distance": [c_i, c_i, c_f, c_i, c_f, c_f, c_f, c_f, c_f, c_f],
    "mpb_nn_distance_grad": [c_i, c_i, c_f, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f],
    "mpb_approxmatch": [c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f],
    "mpb_matchcost": [c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f],
    "mpb_matchcostgrad": [c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f],
}


def load():
    """Load libmonopsr_b200.so (built by ``monopsr_b200.build``); raise if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "monopsr_b200: %s is missing -- run `python -m monopsr_b200.build` "
            "(or __graft_entry__.build()). There is no CPU/PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mpb_version.restype = ctypes.c_char_p
    lib.mpb_launch_count.restype = ctypes.c_ulonglong
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = sig
        fn.restype = c_i
    from . import lib_net
    lib_net.declare(lib)
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        if status > 0:
            raise MpbError("%s failed: CUDA error %d" % (what, status))
        raise MpbError("%s failed: invalid argument (status %d)" % (what, status))


def launch_count():
    return int(load().mpb_launch_count())


def require_cuda(t, name, dtype=None):
    """Validate a tensor argument: CUDA, contiguous, dtype.  No silent copies to/from CPU."""
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor on a CUDA device (got %r)" % (name, type(t)))
    if not t.is_cuda:
        raise MpbError("%s must live on a CUDA device: monopsr_b200 has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must have dtype %s (got %s)" % (name, dtype, t.dtype))
    return t.contiguous()


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
