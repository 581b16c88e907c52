"""oracle/tfops.py -- TEST INFRASTRUCTURE ONLY: ctypes front-end for the CPU oracles.

* ``libtfops_oracle.so``  (oracle/tfops_oracle.c) -- the committed C restatement, both
  GPU-order (primary) and CPU-order flavours.
* ``_ref/libtfops_ref_cpu.so`` -- the reference's own CPU functions, compiled from the
  read-only checkout by oracle/build_ref.sh (present when that was run; optional).

All functions take/return numpy arrays.  Shapes follow the reference op API
(src/tf_ops/nn_distance/tf_nndistance.py:15-25, src/tf_ops/approxmatch/tf_approxmatch.py:15-43).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def _fp(a):
    return a.ctypes.data_as(_F)


def _ip(a):
    return a.ctypes.data_as(_I)


def build(force=False):
    """Compile the C restatement (and, when the reference checkout is present, _ref)."""
    so = os.path.join(_HERE, "libtfops_oracle.so")
    src = os.path.join(_HERE, "tfops_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libtfops_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(os.environ.get("MONOPSR_REFERENCE", "/root/reference")):
        if force or not all(os.path.exists(os.path.join(_HERE, "_ref", f))
                            for f in ("libtfops_ref_cpu.so", "evaluate_object_3d_offline")):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(os.path.join(_HERE, "libtfops_oracle.so"))
    return _lib


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libtfops_ref_cpu.so"))


def ref():
    """The reference's own CPU functions (oracle/_ref); raises if not built."""
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(os.path.join(_HERE, "_ref", "libtfops_ref_cpu.so"))
    return _ref


def _prep(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 3 and x.shape[2] == 3
    return x


def nn_distance(xyz1, xyz2, order="gpu"):
    """-> dist1 (b,n) f32, idx1 (b,n) i32, dist2 (b,m), idx2 (b,m).

    order='gpu' : fma rounding order of the reference CUDA kernel (primary oracle)
    order='cpu' : reference CPU nnsearch rounding order
    order='ref' : the reference's own nnsearch from oracle/_ref
    """
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    assert xyz2.shape[0] == b
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    if order == "ref":
        r = ref()
        r.ref_nnsearch(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1))
        r.ref_nnsearch(b, m, n, _fp(xyz2), _fp(xyz1), _fp(d2), _ip(i2))
    else:
        fn = lib().nn_distance_gpuorder if order == "gpu" else lib().nn_distance_cpuorder
        fn(b, n, _fp(xyz1), m, _fp(xyz2), _fp(d1), _ip(i1), _fp(d2), _ip(i2))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.ascontiguousarray(grad_dist1, np.float32)
    g2 = np.ascontiguousarray(grad_dist2, np.float32)
    i1 = np.ascontiguousarray(idx1, np.int32)
    i2 = np.ascontiguousarray(idx2, np.int32)
    o1 = np.empty((b, n, 3), np.float32)
    o2 = np.empty((b, m, 3), np.float32)
    lib().nn_distance_grad(b, n, _fp(xyz1), m, _fp(xyz2), _fp(g1), _ip(i1), _fp(g2), _ip(i2), _fp(o1), _fp(o2))
    return o1, o2


def approx_match(xyz1, xyz2, order="gpu"):
    """-> match.  order='gpu': (b,m,n) [l,k] layout, 10 levels (the documented API,
    tf_approxmatch.py:21).  order='cpu'/'ref': the CPU code's n-major fill, 11 levels,
    returned as a (b,n,m) array (quirks Q2/Q3)."""
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if order == "gpu":
        out = np.empty((b, m, n), np.float32)
        lib().approxmatch_gpuorder(b, n, m, _fp(xyz1), _fp(xyz2), _fp(out))
    else:
        out = np.empty((b, n, m), np.float32)
        fn = ref().ref_approxmatch_cpu if order == "ref" else lib().approxmatch_cpuorder
        fn(b, n, m, _fp(xyz1), _fp(xyz2), _fp(out))
    return out


def match_cost(xyz1, xyz2, match, order="gpu"):
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.ascontiguousarray(match, np.float32)
    out = np.empty((b,), np.float32)
    fn = {"gpu": lambda: lib().matchcost_gpuorder, "cpu": lambda: lib().matchcost_cpuorder,
          "ref": lambda: ref().ref_matchcost_cpu}[order]()
    fn(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match), _fp(out))
    return out


def match_cost_grad(xyz1, xyz2, match, order="gpu"):
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.ascontiguousarray(match, np.float32)
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    fn = {"gpu": lambda: lib().matchcostgrad_gpuorder, "cpu": lambda: lib().matchcostgrad_cpuorder,
          "ref": lambda: ref().ref_matchcostgrad_cpu}[order]()
    fn(b, n, m, _fp(xyz1), _fp(xyz2), _fp(match), _fp(g1), _fp(g2))
    return g1, g2
