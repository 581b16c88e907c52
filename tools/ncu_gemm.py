"""The representative block3 GEMM launches of one training step, launched once each after a warm-up, for
`ncu --set full -k regex:tc_gemm_tma` (profiles/*_ncu_gemm*).  With --x3: the forward cases through the 3xTF32 kernel
(`ncu --set full -k regex:tc_gemm_x3 ... python tools/ncu_gemm.py --x3`)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep_lib import *  # noqa

CASES = [
    # name, op, nimg, H, W, k, dil, Cin, Cout, epi, BN, ksplit, atomic
    ("fwd 3x3 256>256 full-image, BN128", TC_FWD, 1, 40, 152, 3, 4, 256, 256, 0, 128, 1, 0),
    ("fwd 1x1 256>1024 +res +second output full-image, BN64", TC_FWD, 1, 40, 152, 1, 1, 256, 1024, 1, 64, 1, 0),
    ("fwd 1x1 1024>256 full-image, BN128", TC_FWD, 1, 40, 152, 1, 1, 1024, 256, 0, 128, 1, 0),
    ("dgrad 3x3 crops, BN64", TC_DGRAD, 32, 12, 12, 3, 4, 256, 256, 0, 64, 1, 0),
    ("wgrad 3x3 full-image, BN128 split-K 4 (RED)", TC_WGRAD, 1, 40, 152, 3, 4, 256, 256, 0, 128, 4, 1),
    ("fwd 3x3 512>256 decoder M=18432, BN256", TC_FWD, 32, 24, 24, 3, 1, 512, 256, 0, 256, 1, 0),
]
X3 = "--x3" in sys.argv
H3 = "--h3" in sys.argv          # forward cases through the fp16-split kernel (the engine's default precision)
for name, op, nimg, H, W, k, dil, Cin, Cout, epi, bn, ks, atomic in CASES:
    if X3 and op != TC_FWD:
        continue
    p, keep = make(op, nimg, H, W, k, dil, Cin, Cout, epi)
    p.ksplit, p.atomic = (1, 0) if (X3 or (H3 and op == TC_FWD)) else (ks, atomic)
    if H3 and op == TC_FWD:
        M = nimg * H * W
        x16, w16 = torch.empty(M, Cin, device=dev), torch.empty(Cout, k * k * Cin, device=dev)
        inv, o16 = torch.ones(Cout, device=dev), torch.empty(M, Cout, device=dev)
        L.mpb_split16(M, Cin, p.X, Cin, x16.data_ptr(), Cin, 0, None, mlib.stream_ptr())
        L.mpb_split16(Cout, k * k * Cin, p.Wt, k * k * Cin, w16.data_ptr(), k * k * Cin, 1, None, mlib.stream_ptr())
        p.X16, p.W16, p.scale = x16.data_ptr(), w16.data_ptr(), inv.data_ptr()
        if epi:                       # the residual-stream epilogue of the h3 engine: fp32 store + split copy, no rounded copy
            p.out_r, p.ldor, p.out16, p.ldo16 = None, 0, o16.data_ptr(), Cout
        keep += [x16, w16, inv, o16]
    for _ in range(2):
        if X3:
            rc = L.mpb_tc_gemm_x3(ctypes.byref(p), min(bn, 128), mlib.stream_ptr())
        elif H3 and op == TC_FWD:
            rc = L.mpb_tc_gemm_h3(ctypes.byref(p), bn, mlib.stream_ptr())
        else:
            rc = L.mpb_tc_gemm(ctypes.byref(p), bn, mlib.stream_ptr())
    torch.cuda.synchronize()
    print(name, rc)
