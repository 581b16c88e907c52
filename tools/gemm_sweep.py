"""Sweep tile width / split-K of the tcgen05 GEMM core on the block3 bottleneck shapes (CUDA events, warm L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep_lib import *  # noqa

SHAPES = [
    ("fwd3x3", TC_FWD, 1, 40, 152, 3, 4, 256, 256, 0),
    ("fwd1x1 1024>256", TC_FWD, 1, 40, 152, 1, 1, 1024, 256, 0),
    ("fwd1x1 256>1024 +res", TC_FWD, 1, 40, 152, 1, 1, 256, 1024, 1),
    ("dgrad3x3", TC_DGRAD, 1, 40, 152, 3, 4, 256, 256, 0),
    ("dgrad K=256 N=1024", TC_DGRAD, 1, 40, 152, 1, 1, 1024, 256, 1),
    ("dgrad K=1024 N=256", TC_DGRAD, 1, 40, 152, 1, 1, 256, 1024, 0),
    ("fwd3x3 crops", TC_FWD, 32, 12, 12, 3, 4, 256, 256, 0),
    ("fwd1x1 1024>256 crops", TC_FWD, 32, 12, 12, 1, 1, 1024, 256, 0),
    ("fwd1x1 256>1024+res crops", TC_FWD, 32, 12, 12, 1, 1, 256, 1024, 1),
    ("wgrad3x3", TC_WGRAD, 1, 40, 152, 3, 4, 256, 256, 0),
    ("wgrad1x1 256>1024", TC_WGRAD, 1, 40, 152, 1, 1, 256, 1024, 0),
    ("wgrad3x3 crops", TC_WGRAD, 32, 12, 12, 3, 4, 256, 256, 0),
]
for bn_ in (64, 128, 256):
    print("max resident clusters BN=%d:" % bn_, {(cx, ks): L.mpb_tc_max_clusters(bn_, cx, ks)
                                                  for cx, ks in ((1, 2), (1, 3), (2, 3), (1, 4), (2, 4), (1, 6), (1, 8))})
for name, op, nimg, H, W, k, dil, Cin, Cout, epi in SHAPES:
    p, keep = make(op, nimg, H, W, k, dil, Cin, Cout, epi)
    fl = 2.0 * p.M * Cin * Cout * k * k
    for bn in (64, 128, 256):
        for ks in (1, 2, 3, 4, 6, 8):
            for mode in ("cluster", "atomic"):
                if ks == 1 and mode != "cluster":
                    continue
                if mode == "atomic" and epi:
                    continue
                p.ksplit, p.atomic = ks, 1 if mode == "atomic" else 0
                us = time_it(p, bn)
                if us is None:
                    continue
                print("%-24s BN=%3d ksplit=%d %-7s %7.1f us  %6.1f TF/s" % (name, bn, ks, mode if ks > 1 else "-", us,
                                                                            fl / us / 1e6), flush=True)
