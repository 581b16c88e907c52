"""Time the point-set ops (product vs the reference's own kernels recompiled for sm_100a).

CUDA-event timing, L2 flushed between iterations.  Prints one JSON object.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200 import lib as mlib  # noqa: E402
from oracle import refgpu  # noqa: E402


def timeit(fn, iters=20, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.sum()      # READ-flush (a write-flush leaves dirty lines whose write-back is charged to the next kernel)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return {"med_us": ts[len(ts) // 2], "p10_us": ts[len(ts) // 10], "p90_us": ts[(len(ts) * 9) // 10]}


def main():
    dev = torch.device("cuda:0")
    L = mlib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    sp = lambda: mlib.stream_ptr()
    for (b, n) in [(32, 2048), (256, 2048), (32, 2304)]:
        g = torch.Generator(device="cpu").manual_seed(100)
        x = torch.randn(b, n, 3, generator=g).to(dev)
        y = torch.randn(b, n, 3, generator=g).to(dev)
        d1 = torch.empty(b, n, device=dev); i1 = torch.empty(b, n, device=dev, dtype=torch.int32)
        d2 = torch.empty(b, n, device=dev); i2 = torch.empty(b, n, device=dev, dtype=torch.int32)
        g1 = torch.empty(b, n, 3, device=dev); g2 = torch.empty(b, n, 3, device=dev)
        ones = torch.ones(b, n, device=dev)
        f = lambda: L.mpb_nn_distance(b, n, x.data_ptr(), n, y.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), sp())
        out[f"nn_fwd_b{b}_n{n}"] = timeit(f, flush=flush)
        f()
        fg = lambda: L.mpb_nn_distance_grad(b, n, x.data_ptr(), n, y.data_ptr(), ones.data_ptr(), i1.data_ptr(), ones.data_ptr(), i2.data_ptr(), g1.data_ptr(), g2.data_ptr(), sp())
        out[f"nn_grad_b{b}_n{n}"] = timeit(fg, flush=flush)
        if refgpu.available():
            # reference launches on the legacy default stream; torch events on the current
            # (default) stream bracket it because the legacy stream serialises with it
            out[f"ref_nn_fwd_b{b}_n{n}"] = timeit(lambda: refgpu.nn_distance_launch(x, y, d1, i1, d2, i2), flush=flush)
    for (b, n) in [(32, 1024), (32, 2304)]:
        g = torch.Generator(device="cpu").manual_seed(200)
        x = torch.randn(b, n, 3, generator=g).to(dev)
        y = torch.randn(b, n, 3, generator=g).to(dev)
        mt = torch.empty(b, n, n, device=dev)
        cost = torch.empty(b, device=dev)
        g1 = torch.empty(b, n, 3, device=dev); g2 = torch.empty(b, n, 3, device=dev)
        out[f"am_match_b{b}_n{n}"] = timeit(lambda: L.mpb_approxmatch(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), None, sp()), iters=10, warmup=2, flush=flush)
        out[f"am_cost_b{b}_n{n}"] = timeit(lambda: L.mpb_matchcost(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), cost.data_ptr(), sp()), flush=flush)
        out[f"am_grad_b{b}_n{n}"] = timeit(lambda: L.mpb_matchcostgrad(b, n, n, x.data_ptr(), y.data_ptr(), mt.data_ptr(), g1.data_ptr(), g2.data_ptr(), sp()), flush=flush)
        if refgpu.available() and n == 1024:
            import ctypes
            rl = refgpu.lib()
            temp = torch.empty(32 * 4 * n, device=dev)
            p = lambda t: ctypes.c_void_p(t.data_ptr())
            out[f"ref_am_match_b{b}_n{n}"] = timeit(lambda: getattr(rl, refgpu._AM)(b, n, n, p(x), p(y), p(mt), p(temp)), iters=5, warmup=1, flush=flush)
            out[f"ref_am_cost_b{b}_n{n}"] = timeit(lambda: getattr(rl, refgpu._MC)(b, n, n, p(x), p(y), p(mt), p(cost)), iters=5, warmup=1, flush=flush)
            out[f"ref_am_grad_b{b}_n{n}"] = timeit(lambda: getattr(rl, refgpu._MCG)(b, n, n, p(x), p(y), p(mt), p(g1), p(g2)), iters=5, warmup=1, flush=flush)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
