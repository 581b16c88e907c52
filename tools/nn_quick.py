"""nn_distance forward time at cfg3 (32 x 2048^2), 256 x 2048^2 and the in-model 32 x 2304^2 + a checksum of the indices"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monopsr_b200 import lib as mlib
from tools.bench_tfops import timeit
L = mlib.load(); dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (b, n) in [(32, 2048), (256, 2048), (32, 2304)]:
    g = torch.Generator(device="cpu").manual_seed(100)
    x = torch.randn(b, n, 3, generator=g).to(dev); y = torch.randn(b, n, 3, generator=g).to(dev)
    d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, device=dev, dtype=torch.int32)
    d2 = torch.empty(b, n, device=dev); i2 = torch.empty(b, n, device=dev, dtype=torch.int32)
    f = lambda: L.mpb_nn_distance(b, n, x.data_ptr(), n, y.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), mlib.stream_ptr())
    r = timeit(f, iters=50, warmup=5, flush=flush)
    print(os.environ.get("TAG", ""), b, n, "%.1f us" % r["med_us"], "idx checksum", int(i1.long().sum()) + 3 * int(i2.long().sum()),
          "dist checksum %.9e" % float(d1.double().sum() + d2.double().sum()))
