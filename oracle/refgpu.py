"""oracle/refgpu.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the reference's UNMODIFIED CUDA kernels compiled for sm_100a
(oracle/_ref/libtfops_ref_gpu.so, built by oracle/build_ref.sh from
src/tf_ops/{nn_distance/tf_nndistance_g.cu, approxmatch/tf_approxmatch_g.cu}).  This is
both the strongest parity oracle on the GPU box (the reference itself, run there) and
"the kernel to beat" in bench.py.  Launchers keep their C++-mangled names and launch on
the legacy default stream, so callers synchronise around them.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libtfops_ref_gpu.so")
_lib = None

_NN = "_Z24NmDistanceKernelLauncheriiPKfiS0_PfPiS1_S2_"
_NNG = "_Z28NmDistanceGradKernelLauncheriiPKfiS0_S0_PKiS0_S2_PfS3_"
_AM = "_Z19approxmatchLauncheriiiPKfS0_PfS1_"
_MC = "_Z17matchcostLauncheriiiPKfS0_S0_Pf"
_MCG = "_Z21matchcostgradLauncheriiiPKfS0_S0_PfS1_"


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
        for name in (_NN, _NNG, _AM, _MC, _MCG):
            getattr(_lib, name).restype = None
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def nn_distance(x, y):
    b, n, _ = x.shape
    m = y.shape[1]
    d1 = torch.empty((b, n), device=x.device)
    i1 = torch.empty((b, n), device=x.device, dtype=torch.int32)
    d2 = torch.empty((b, m), device=x.device)
    i2 = torch.empty((b, m), device=x.device, dtype=torch.int32)
    torch.cuda.synchronize()
    getattr(lib(), _NN)(b, n, _p(x), m, _p(y), _p(d1), _p(i1), _p(d2), _p(i2))
    torch.cuda.synchronize()
    return d1, i1, d2, i2


def nn_distance_launch(x, y, d1, i1, d2, i2):
    """un-synchronised launch on the legacy stream (for timing)"""
    getattr(lib(), _NN)(x.shape[0], x.shape[1], _p(x), y.shape[1], _p(y), _p(d1), _p(i1), _p(d2), _p(i2))


def nn_distance_grad(x, y, gd1, i1, gd2, i2):
    b, n, _ = x.shape
    m = y.shape[1]
    g1 = torch.empty((b, n, 3), device=x.device)
    g2 = torch.empty((b, m, 3), device=x.device)
    torch.cuda.synchronize()
    getattr(lib(), _NNG)(b, n, _p(x), m, _p(y), _p(gd1), _p(i1), _p(gd2), _p(i2), _p(g1), _p(g2))
    torch.cuda.synchronize()
    return g1, g2


def approx_match(x, y):
    b, n, _ = x.shape
    m = y.shape[1]
    match = torch.empty((b, m, n), device=x.device)
    # the kernel indexes scratch by blockIdx.x of a fixed 32-CTA grid (tf_approxmatch_g.cu:2,181)
    temp = torch.empty((max(b, 32) * (n + m) * 2,), device=x.device)
    torch.cuda.synchronize()
    getattr(lib(), _AM)(b, n, m, _p(x), _p(y), _p(match), _p(temp))
    torch.cuda.synchronize()
    return match


def match_cost(x, y, match):
    b, n, _ = x.shape
    m = y.shape[1]
    out = torch.empty((b,), device=x.device)
    torch.cuda.synchronize()
    getattr(lib(), _MC)(b, n, m, _p(x), _p(y), _p(match), _p(out))
    torch.cuda.synchronize()
    return out


def match_cost_grad(x, y, match):
    b, n, _ = x.shape
    m = y.shape[1]
    g1 = torch.empty((b, n, 3), device=x.device)
    g2 = torch.empty((b, m, 3), device=x.device)
    torch.cuda.synchronize()
    getattr(lib(), _MCG)(b, n, m, _p(x), _p(y), _p(match), _p(g1), _p(g2))
    torch.cuda.synchronize()
    return g1, g2
