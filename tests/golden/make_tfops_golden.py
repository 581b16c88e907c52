"""Generate tests/golden/tfops_golden.npz from the REFERENCE'S OWN CPU functions.

Run in the authoring container (needs /root/reference; oracle/build_ref.sh compiles the
reference functions where they lie into oracle/_ref).  The fixtures pin the oracle's
cpu-order flavour to the real reference on seeded inputs, so they keep working on the
GPU box where /root/reference does not exist.

    python tests/golden/make_tfops_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tfops  # noqa: E402


def main():
    tfops.build(force=True)
    assert tfops.ref_available(), "oracle/_ref not built (no /root/reference?)"
    out = {}
    cases = {"a": (3, 37, 53, 11), "b": (2, 64, 64, 12), "c": (1, 5, 130, 13)}
    for name, (b, n, m, seed) in cases.items():
        rng = np.random.RandomState(seed)
        x = rng.randn(b, n, 3).astype(np.float32)
        y = rng.randn(b, m, 3).astype(np.float32)
        if name == "b":   # in-model distribution: 40 % of both clouds are the same (0,0,0) point
            mask = rng.rand(b, n, 1) < 0.4
            x = np.where(mask, 0, x).astype(np.float32)
            y = np.where(mask, 0, y).astype(np.float32)
        d1, i1, d2, i2 = tfops.nn_distance(x, y, "ref")
        out.update({f"{name}_x": x, f"{name}_y": y, f"{name}_d1": d1, f"{name}_i1": i1,
                    f"{name}_d2": d2, f"{name}_i2": i2})
        if name == "c":
            continue   # integer n/m ratio required by approxmatch (quirk Q5)
        mt = tfops.approx_match(x, y, "ref")          # (b,n,m) n-major, 11 levels
        out[f"{name}_match"] = mt
        out[f"{name}_cost"] = tfops.match_cost(x, y, mt, "ref")
        g1, g2 = tfops.match_cost_grad(x, y, mt, "ref")
        out[f"{name}_g1"], out[f"{name}_g2"] = g1, g2
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tfops_golden.npz"), **out)
    print("wrote tfops_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
