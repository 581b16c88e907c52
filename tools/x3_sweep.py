"""3xTF32 forward GEMM (mpb_tc_gemm_x3) next to the single-pass kernel on the forward shapes of the network:
us per launch (20 launches per CUDA graph, warm L2) and the relative error of both against an fp64 contraction."""
import ctypes, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gemm_sweep_lib as gs
from monopsr_b200.lib_net import TC_FWD, TcGemmParams

SHAPES = [  # nimg, H, W, k, dil, Cin, Cout   (Appendix A of SURVEY.md: block3 bottleneck, decoder, squash, FC)
    (1, 40, 152, 1, 1, 1024, 256), (1, 40, 152, 3, 2, 256, 256), (1, 40, 152, 1, 1, 256, 1024),
    (32, 12, 12, 1, 1, 1024, 256), (32, 12, 12, 3, 2, 256, 256), (32, 12, 12, 1, 1, 256, 1024),
    (32, 12, 12, 1, 1, 2048, 512), (32, 24, 24, 3, 1, 512, 256), (32, 48, 48, 3, 1, 256, 128), (32, 48, 48, 3, 1, 128, 128),
]
print("%-34s %9s %9s %7s %10s %10s" % ("shape", "tf32 us", "x3 us", "ratio", "tf32 err", "x3 err"))
for nimg, H, W, k, dil, Cin, Cout in SHAPES:
    M = nimg * H * W
    x = torch.randn(nimg, H, W, Cin, device=gs.dev)
    w = torch.randn(Cout, k, k, Cin, device=gs.dev) / (k * k * Cin) ** 0.5
    out = torch.empty(M, Cout, device=gs.dev)
    p = TcGemmParams()
    p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout, p.ksplit = TC_FWD, H, W, k, k, dil, M, Cin, Cout, 1
    p.X, p.ldx, p.Wt, p.ldw, p.out, p.ldo = x.data_ptr(), Cin, w.data_ptr(), k * k * Cin, out.data_ptr(), Cout
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), padding=dil * (k // 2), dilation=dil)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    bn = 128 if Cout % 128 == 0 else 64
    t1 = gs.time_it(p, bn)
    e1 = float((out.double() - ref).norm() / ref.norm())
    out.zero_()
    t3 = gs.time_it(p, bn, x3=True)
    e3 = float((out.double() - ref).norm() / ref.norm())
    print("%-34s %9.1f %9.1f %7.2f %10.2e %10.2e" % ((nimg, H, W, k, dil, Cin, Cout), t1 or -1, t3 or -1,
                                                     (t3 / t1) if t1 and t3 else 0, e1, e3))
