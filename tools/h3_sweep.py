"""Forward GEMM variants side by side on the forward shapes of the network: single-pass tf32, 3xTF32 (x3) and the fp16
hi/lo split kernel (h3) -- us per launch (20 launches per CUDA graph, warm L2) and relative error against fp64."""
import ctypes, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gemm_sweep_lib as gs
from test_h3_gpu import _split16, _weights16
from monopsr_b200.lib_net import TC_FWD, TcGemmParams

SHAPES = [  # nimg, H, W, k, dil, Cin, Cout   (Appendix A of SURVEY.md: block3 bottleneck, decoder, squash, FC)
    (1, 40, 152, 1, 1, 1024, 256), (1, 40, 152, 3, 2, 256, 256), (1, 40, 152, 1, 1, 256, 1024),
    (32, 12, 12, 1, 1, 1024, 256), (32, 12, 12, 3, 2, 256, 256), (32, 12, 12, 1, 1, 256, 1024),
    (32, 12, 12, 1, 1, 2048, 512), (32, 24, 24, 3, 1, 512, 256), (32, 48, 48, 3, 1, 256, 128), (32, 48, 48, 3, 1, 128, 128),
]
print("%-32s %4s %8s %8s %8s %9s %9s %9s" % ("shape", "BN", "tf32 us", "x3 us", "h3 us", "tf32 err", "x3 err", "h3 err"))
for nimg, H, W, k, dil, Cin, Cout in SHAPES:
    M = nimg * H * W
    x = torch.randn(nimg, H, W, Cin, device=gs.dev)
    w = torch.randn(Cout, k, k, Cin, device=gs.dev) / (k * k * Cin) ** 0.5
    out = torch.empty(M, Cout, device=gs.dev)
    x16 = _split16(x.reshape(M, Cin))
    w16, inv = _weights16(w.reshape(Cout, -1))
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), padding=dil * (k // 2), dilation=dil)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    for bn in (64, 128, 256):
        if Cout % bn:
            continue
        p = TcGemmParams()
        p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout, p.ksplit = TC_FWD, H, W, k, k, dil, M, Cin, Cout, 1
        p.X, p.ldx, p.Wt, p.ldw, p.out, p.ldo = x.data_ptr(), Cin, w.data_ptr(), k * k * Cin, out.data_ptr(), Cout
        res = {}
        for kind in ("tf32", "x3", "h3"):
            if kind == "x3" and bn == 256:
                res[kind] = (None, float("nan"))
                continue
            if kind == "h3":
                p.X16, p.W16, p.scale = x16.data_ptr(), w16.data_ptr(), inv.data_ptr()
            out.zero_()
            t = gs.time_it(p, bn, kind=kind)
            res[kind] = (t, float((out.double() - ref).norm() / ref.norm()))
            p.scale = None
        print("%-32s %4d %8.1f %8.1f %8.1f %9.2e %9.2e %9.2e" % ((nimg, H, W, k, dil, Cin, Cout), bn, res["tf32"][0] or -1,
              res["x3"][0] or -1, res["h3"][0] or -1, res["tf32"][1], res["x3"][1], res["h3"][1]))
