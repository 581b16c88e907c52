"""Per-shape timing of every tcgen05 GEMM launch of one training step: each distinct launch configuration is
replayed 20x as one CUDA graph (isolated, warm L2) and timed with CUDA events.  `sm_us` = time x the fraction of
the GPU's CTA slots the launch occupies -- the part of the machine it would need under perfect packing."""
import collections, ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from monopsr_b200.core import model_spec as ms
from monopsr_b200.core.engine import Engine
import gemm_sweep_lib as gs

dev = torch.device("cuda:0")
eng = Engine(dev, params=ms.init_params(0))
eng.set_inputs(ms.synthetic_sample(0))
eng.forward(True); eng.backward(); torch.cuda.synchronize()
eng._record = []
eng.forward(True); eng.backward()
rec, eng._record = eng._record, None
torch.cuda.synchronize()
groups = collections.OrderedDict()
for p, bn, fl in rec:
    key = (p.op, p.M, p.H, p.W, p.kh, p.dil, p.Cin, p.Cout, bn, p.ksplit, p.atomic, bool(p.res), bool(p.mask), bool(p.out_r))
    groups.setdefault(key, []).append((p, bn, fl))
rows = []
for key, items in groups.items():
    p, bn, fl = items[0]
    us = gs.time_it(p, abs(bn), x3=bn < 0)          # negative tile width: a 3xTF32 forward launch (MPB_PRECISION=x3)
    bn = abs(bn)
    ncols = {0: p.Cout, 1: p.Cin, 2: p.kh * p.kw * p.Cin}[p.op]
    nrows = p.Cout if p.op == 2 else p.M
    ctas = ((nrows + 127) // 128) * (ncols // bn) * p.ksplit
    slots = 148 * (2 if bn == 64 else 1)
    sm_us = us * min(1.0, ctas / slots)
    rows.append((us * len(items), us, len(items), fl / us / 1e6, ctas, sm_us * len(items), key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total %.1f us over %d launches, %d shapes; perfectly packed: %.1f us" % (tot, len(rec), len(rows), sum(r[5] for r in rows)))
print("%9s %8s %4s %8s %5s %9s  op M H W k dil Cin Cout BN ksplit atomic res mask out_r" % ("tot_us", "us", "n", "TFLOP/s", "ctas", "sm_us"))
for r in rows[:60]:
    print("%9.1f %8.1f %4d %8.1f %5d %9.1f  %s" % r)
