"""Median time of match_cost / match_cost_grad at cfg4 (32 x 1024^2) and the in-model size (32 x 2304^2);
L2 read-flushed before every launch.  Select a library variant with MPB_LIB=..., the register-ring kernel with MPB_MG_TMA=0."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monopsr_b200 import lib as mlib
from tools.bench_tfops import timeit
L = mlib.load(); dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = os.environ.get("TAG", "")
for n in (1024, 2304):
    g = torch.Generator(device="cpu").manual_seed(200)
    x = torch.randn(32, n, 3, generator=g).to(dev); y = torch.randn(32, n, 3, generator=g).to(dev)
    mt = torch.rand(32, n, n, device=dev) / n
    cost = torch.empty(32, device=dev); g1 = torch.empty(32, n, 3, device=dev); g2 = torch.empty(32, n, 3, device=dev)
    sp = mlib.stream_ptr
    c = timeit(lambda: L.mpb_matchcost(32, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), cost.data_ptr(), sp()), iters=100, warmup=10, flush=flush)
    gr = timeit(lambda: L.mpb_matchcostgrad(32, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), g1.data_ptr(), g2.data_ptr(), sp()), iters=100, warmup=10, flush=flush)
    gb = 32 * n * n * 4 / 1e3
    print("%s n=%d cost %.1f us (%.2f TB/s) grad %.1f us (%.2f TB/s) [p10 %.1f p90 %.1f]" % (
        tag, n, c["med_us"], gb / c["med_us"] / 1e3, gr["med_us"], gb / gr["med_us"] / 1e3, gr["p10_us"], gr["p90_us"]))
