"""The point-set ops at BASELINE cfg3 / cfg4, launched three times each after a read-flush of L2, for
    ncu --set full --clock-control none --import-source on \
        -k regex:'nn_distance_kernel|nn_distance_grad|approxmatch_cluster_kernel|matchcost_stream|matchcostgrad_stream|matchcostgrad_tma|ms_sum_planes' \
        -o gpurun_out/r2_ncu_tfops python tools/ncu_tfops.py
(summary: ncu -i rep --page raw --csv > raw.csv; python tools/ncu_table.py raw.csv)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200 import lib as mlib
L = mlib.load()
dev = torch.device("cuda:0")
sp = mlib.stream_ptr
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device="cpu").manual_seed(100)
b, n = 32, 2048
x, y = torch.randn(b, n, 3, generator=g).to(dev), torch.randn(b, n, 3, generator=g).to(dev)
d1, d2 = torch.empty(b, n, device=dev), torch.empty(b, n, device=dev)
i1, i2 = torch.empty(b, n, device=dev, dtype=torch.int32), torch.empty(b, n, device=dev, dtype=torch.int32)
g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev)
one = torch.ones(b, n, device=dev)
for _ in range(3):
    flush.sum()
    L.mpb_nn_distance(b, n, x.data_ptr(), n, y.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), sp())
    L.mpb_nn_distance_grad(b, n, x.data_ptr(), n, y.data_ptr(), one.data_ptr(), i1.data_ptr(), one.data_ptr(), i2.data_ptr(),
                           g1.data_ptr(), g2.data_ptr(), sp())
g = torch.Generator(device="cpu").manual_seed(200)
n = 1024
x, y = torch.randn(b, n, 3, generator=g).to(dev), torch.randn(b, n, 3, generator=g).to(dev)
mt = torch.empty(b, n, n, device=dev)
cost = torch.empty(b, device=dev)
g1, g2 = torch.empty(b, n, 3, device=dev), torch.empty(b, n, 3, device=dev)
for _ in range(3):
    flush.sum()
    L.mpb_approxmatch(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), None, sp())
    flush.sum()
    L.mpb_matchcost(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), cost.data_ptr(), sp())
    flush.sum()
    L.mpb_matchcostgrad(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), g1.data_ptr(), g2.data_ptr(), sp())
torch.cuda.synchronize()
print("done")
