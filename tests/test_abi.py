"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol the headers declare; the host-side mirror refuses CPU tensors (no fallback)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms += re.findall(r"\b(mpb_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(syms))


def test_library_builds_and_exports_every_declared_symbol():
    from monopsr_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = _declared_symbols()
    assert len(syms) >= 7
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    lib.mpb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mpb_version()


def test_only_sm100a_code_in_library():
    import subprocess
    from monopsr_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_ops_refuse_cpu_tensors():
    import torch
    from monopsr_b200.lib import MpbError
    from monopsr_b200.tf_ops.nn_distance import tf_nndistance
    from monopsr_b200.tf_ops.approxmatch import tf_approxmatch
    x = torch.zeros(1, 4, 3)
    with pytest.raises(MpbError):
        tf_nndistance.nn_distance(x, x)
    with pytest.raises(MpbError):
        tf_approxmatch.approx_match(x, x)


def test_product_package_never_imports_oracle():
    for p in glob.glob(os.path.join(ROOT, "monopsr_b200", "**", "*.py"), recursive=True):
        src = open(p).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p
