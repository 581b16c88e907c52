"""Per-shape timing of every tcgen05 GEMM launch of one training step (CUDA events, isolated, warm)."""
import collections, ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200.core import model_spec as ms
from monopsr_b200.core.engine import Engine

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
    key = (p.op, p.M, p.H, p.W, p.kh, p.dil, p.Cin, p.Cout, bn, p.ksplit, p.atomic)
    groups.setdefault(key, []).append((p, bn, fl))
rows = []
for key, items in groups.items():
    p, bn, fl = items[0]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        eng.L.mpb_tc_gemm(ctypes.byref(p), bn, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        eng.L.mpb_tc_gemm(ctypes.byref(p), bn, st)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    rows.append((us * len(items), us, len(items), fl / us / 1e6, key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total %.1f us over %d launches, %d shapes" % (tot, len(rec), len(rows)))
print("%9s %8s %4s %8s  op M H W k dil Cin Cout BN ksplit atomic" % ("tot_us", "us", "n", "TFLOP/s"))
for r in rows[:45]:
    print("%9.1f %8.1f %4d %8.1f  %s" % (r[0], r[1], r[2], r[3], r[4]))
