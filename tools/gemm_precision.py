"""Measure the arithmetic error of the tcgen05 tf32 path itself (inputs exactly representable in tf32)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monopsr_b200 import lib as mlib
from monopsr_b200.lib_net import TcGemmParams, TC_FWD

def q(t):
    i = t.contiguous().view(torch.int32); i = (i + 0x1000) & ~0x1FFF; return i.view(torch.float32)

dev = torch.device("cuda:0")
L = mlib.load()
for (M, K, N, pos) in [(4608, 256, 256, False), (4608, 2304, 256, False), (4608, 2304, 256, True), (4608, 18432, 1024, True)]:
    g = torch.Generator(device="cpu").manual_seed(0)
    a = torch.randn(M, K, generator=g); b = torch.randn(N, K, generator=g) * 0.05
    if pos: a = a.abs()
    a, b = q(a).to(dev), q(b).to(dev)
    out = torch.empty(M, N, device=dev)
    p = TcGemmParams(); p.op=TC_FWD; p.H=M; p.W=1; p.kh=p.kw=1; p.dil=1; p.M=M; p.Cin=K; p.Cout=N
    p.X=a.data_ptr(); p.ldx=K; p.Wt=b.data_ptr(); p.ldw=K; p.out=out.data_ptr(); p.ldo=N; p.ksplit=1
    mlib.check(L.mpb_tc_gemm(ctypes.byref(p), 64, mlib.stream_ptr()), "gemm"); torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    torch.backends.cuda.matmul.allow_tf32 = True
    cub = a @ b.T
    torch.backends.cuda.matmul.allow_tf32 = False
    f32 = a @ b.T
    rms = ref.pow(2).mean().sqrt()
    for name, o in (("tcgen05", out), ("cublas_tf32", cub), ("cublas_fp32", f32)):
        e = o.double() - ref
        print("M%d K%d N%d pos=%s %-12s rms_err/rms %.3e  max_err/rms %.3e  mean_err/rms %.3e" % (M, K, N, pos, name, e.pow(2).mean().sqrt()/rms, e.abs().max()/rms, e.mean()/rms))

print("---- unrounded A operand (does the tensor core truncate or round fp32 -> tf32?)")
for (M, K, N) in [(4608, 2304, 256)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(M, K, generator=g).abs().to(dev)            # NOT rounded
    b = q(torch.randn(N, K, generator=g).abs() * 0.05).to(dev)  # rounded, positive => no cancellation
    out = torch.empty(M, N, device=dev)
    p = TcGemmParams(); p.op=TC_FWD; p.H=M; p.W=1; p.kh=p.kw=1; p.dil=1; p.M=M; p.Cin=K; p.Cout=N
    p.X=a.data_ptr(); p.ldx=K; p.Wt=b.data_ptr(); p.ldw=K; p.out=out.data_ptr(); p.ldo=N; p.ksplit=1
    mlib.check(L.mpb_tc_gemm(ctypes.byref(p), 64, mlib.stream_ptr()), "gemm"); torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    ref_trunc = (a.view(torch.int32) & ~0x1FFF).view(torch.float32).double() @ b.double().T
    ref_rna = q(a).double() @ b.double().T
    rel = lambda o, r: (float(((o.double() - r) / r).mean()), float(((o.double() - r) / r).pow(2).mean().sqrt()))
    print("vs exact      mean rel %.3e rms rel %.3e" % rel(out, ref))
    print("vs truncated  mean rel %.3e rms rel %.3e" % rel(out, ref_trunc))
    print("vs rna        mean rel %.3e rms rel %.3e" % rel(out, ref_rna))
    print("corrected (1+2^-11) vs exact mean rel %.3e rms rel %.3e" % rel(out * (1 + 2.0 ** -11), ref))
