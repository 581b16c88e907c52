"""Per-CTA clock64 timeline of the TMA GEMM kernel (needs a build with MPB_NVCC_EXTRA=-DMPB_TC_TRACE).

slots: 0 entry, 1 setup done, 2 smid, 4+i producer issues k-block i, 20+i MMA warp sees k-block i full,
36+i MMAs of k-block i issued, 52 accumulator complete (epilogue wakes), 53 first chunk stored, 54 epilogue done,
55 CTA exit, 56 nk."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from monopsr_b200 import lib as mlib
from monopsr_b200.lib_net import TC_DGRAD, TC_FWD, TC_WGRAD
import gemm_sweep_lib as gs

L = mlib.load()
L.mpb_tc_set_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
buf = torch.zeros(1024 * 64, dtype=torch.int64, device=dev)

CASES = [
    ("fwd3x3 BN64", TC_FWD, 1, 40, 152, 3, 4, 256, 256, 0, 64, 1),
    ("fwd3x3 BN256 ks3", TC_FWD, 1, 40, 152, 3, 4, 256, 256, 0, 256, 3),
    ("fwd3x3 BN256 ks1", TC_FWD, 1, 40, 152, 3, 4, 256, 256, 0, 256, 1),
    ("fwd1x1 1024>256 BN256", TC_FWD, 1, 40, 152, 1, 1, 1024, 256, 0, 256, 1),
    ("fwd1x1 1024>256 BN64", TC_FWD, 1, 40, 152, 1, 1, 1024, 256, 0, 64, 1),
    ("fwd1x1 256>1024 +res BN64", TC_FWD, 1, 40, 152, 1, 1, 256, 1024, 1, 64, 1),
    ("fwd1x1 256>1024 +res BN128", TC_FWD, 1, 40, 152, 1, 1, 256, 1024, 1, 128, 1),
]
for name, op, nimg, H, W, k, dil, Cin, Cout, epi, bn, ks in CASES:
    p, keep = gs.make(op, nimg, H, W, k, dil, Cin, Cout, epi)
    p.ksplit, p.atomic = ks, 1 if ks > 1 else 0
    L.mpb_tc_set_trace(None)
    us = gs.time_it(p, bn)
    buf.zero_()
    L.mpb_tc_set_trace(buf.data_ptr())
    L.mpb_tc_gemm(ctypes.byref(p), bn, mlib.stream_ptr())
    torch.cuda.synchronize()
    L.mpb_tc_set_trace(None)
    t = buf.cpu().numpy().reshape(1024, 64)
    live = t[:, 0] > 0
    t = t[live]
    n = len(t)
    t0 = t[:, 0].min()
    nk = int(t[0, 56])
    m = min(nk, 16)
    print("== %s: %.1f us, %d CTAs traced, nk=%d, SMs used %d" % (name, us, n, nk, len(set(t[:, 2]))))
    rel = lambda s: (t[:, s] - t[:, 0])
    g0 = t[:, 3].min()
    st_ns, en_ns = np.sort(t[:, 3] - g0), t[:, 57] - g0
    print("   CTA start (ns, globaltimer) pctl 50/75/90/100: %d %d %d %d ; last exit %d ns" % (
        st_ns[n // 2], st_ns[(3 * n) // 4], st_ns[(9 * n) // 10], st_ns[-1], en_ns.max()))
    percta = (t[:, 57] - t[:, 3])
    print("   per-CTA lifetime ns: med %d" % np.median(percta))
    print("   setup            : med %d" % np.median(rel(1)))
    print("   first TMA issue  : med %d" % np.median(rel(4)))
    print("   first full       : med %d  (TMA latency of k-block 0: %d)" % (np.median(rel(20)), np.median(t[:, 20] - t[:, 4])))
    lat = np.stack([t[:, 20 + i] - t[:, 4 + i] for i in range(m)], 1)
    print("   issue->full per k-block (med): %s" % np.median(lat, 0).astype(int).tolist())
    gap = np.stack([t[:, 20 + i + 1] - t[:, 20 + i] for i in range(m - 1)], 1)
    print("   full(i+1)-full(i) (med): %s" % np.median(gap, 0).astype(int).tolist())
    iss = np.stack([t[:, 4 + i + 1] - t[:, 4 + i] for i in range(m - 1)], 1)
    print("   issue(i+1)-issue(i) (med): %s" % np.median(iss, 0).astype(int).tolist())
    print("   mma issue cost full->issued (med): %d" % np.median(t[:, 36] - t[:, 20]))
    print("   acc complete     : med %d" % np.median(rel(52)))
    print("   first chunk done : med +%d" % np.median(t[:, 53] - t[:, 52]))
    print("   epilogue done    : med +%d" % np.median(t[:, 54] - t[:, 52]))
    print("   exit             : med %d  max %d" % (np.median(rel(55)), rel(55).max()))
    sys.stdout.flush()
