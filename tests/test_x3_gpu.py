"""GPU cases of the 3xTF32 forward path (csrc/tc_gemm.cu::tc_gemm_x3_kernel), the engine's default precision.
First run on a B200 in round 2 (profiles/r2_first_call_summary.txt): all cases pass; the error against fp64 grows
linearly with the reduction length (~1.4e-8 * K: the tensor core's fp32 accumulator truncates), 5e-6 at K = 2304."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monopsr_b200.core import model_spec as ms  # noqa: E402
from monopsr_b200.core.engine import Engine  # noqa: E402
from oracle import network as onet  # noqa: E402

X3_CASES = [
    # nimg, H, W, k, dil, Cin, Cout, BN, ksplit
    (1, 8, 16, 1, 1, 64, 64, 64, 1),          # one tile, two k-blocks (< pipeline depth)
    (2, 12, 12, 1, 1, 256, 128, 128, 1),      # tail rows
    (3, 12, 12, 3, 1, 64, 64, 64, 1),         # im2col, padding taps
    (2, 12, 12, 3, 4, 256, 256, 128, 1),      # atrous, 72 k-blocks, 2 column tiles
    (1, 40, 152, 3, 2, 128, 128, 128, 1),     # full-image geometry
    (32, 1, 1, 1, 1, 1056, 1024, 64, 4),      # FC, atomic split-K
]


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN,ksplit", X3_CASES)
def test_x3_gemm_matches_fp64(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    """mpb_tc_gemm_x3 on UNROUNDED fp32 operands vs an fp64 convolution, with the fused epilogue options on"""
    import ctypes
    import torch.nn.functional as F
    from monopsr_b200 import lib as mlib
    from monopsr_b200.lib_net import TC_FWD, TcGemmParams
    g = torch.Generator(device="cpu").manual_seed(5)
    M = nimg * H * W
    x = torch.randn(nimg, H, W, Cin, generator=g).to(cuda)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(cuda)
    shift = torch.randn(Cout, generator=g).to(cuda)
    res = torch.randn(M, Cout, generator=g).to(cuda)
    atomic = ksplit > 1
    out = torch.zeros(M, Cout, device=cuda) if atomic else torch.full((M, Cout), float("nan"), device=cuda)
    p = TcGemmParams()
    p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout = TC_FWD, H, W, k, k, dil, M, Cin, Cout
    p.X, p.ldx, p.Wt, p.ldw, p.out, p.ldo = x.data_ptr(), Cin, w.data_ptr(), k * k * Cin, out.data_ptr(), Cout
    p.ksplit, p.atomic = ksplit, 1 if atomic else 0
    if not atomic:
        p.shift, p.res, p.ldr, p.relu = shift.data_ptr(), res.data_ptr(), Cout, 1
    mlib.check(mlib.load().mpb_tc_gemm_x3(ctypes.byref(p), BN, mlib.stream_ptr()), "mpb_tc_gemm_x3")
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), padding=dil * (k // 2), dilation=dil)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    if not atomic:
        ref = torch.relu(ref + shift.double() + res.double())
    err = float((out.double() - ref).norm() / ref.norm())
    assert err < 2e-5, err          # measured 2e-6 (K = 256) .. 5e-6 (K = 2304); single-pass tf32 on these operands: ~4e-4


def test_x3_rejects_what_it_does_not_implement(cuda):
    import ctypes
    from monopsr_b200 import lib as mlib
    from monopsr_b200.lib_net import TC_DGRAD, TC_FWD, TcGemmParams
    x = torch.zeros(128, 64, device=cuda)
    p = TcGemmParams()
    p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout, p.ksplit = TC_FWD, 8, 16, 1, 1, 1, 128, 64, 64, 1
    p.X, p.ldx, p.Wt, p.ldw, p.out, p.ldo = x.data_ptr(), 64, x.data_ptr(), 64, x.data_ptr(), 64
    L = mlib.load()
    assert L.mpb_tc_gemm_x3(ctypes.byref(p), 256, mlib.stream_ptr()) == -1       # tile width
    p.ksplit = 2
    assert L.mpb_tc_gemm_x3(ctypes.byref(p), 64, mlib.stream_ptr()) == -1        # non-atomic split-K
    p.ksplit, p.op = 1, TC_DGRAD
    assert L.mpb_tc_gemm_x3(ctypes.byref(p), 64, mlib.stream_ptr()) == -1        # forward only


def test_x3_engine_forward_meets_the_parity_bar(cuda):
    """precision="x3": every output of the forward pass within 1e-3 (measured: < 1e-4) of the fp64 restatement --
    including the decoder's local xyz maps, which single-pass tf32 misses by 2.6x -- and a finite training step"""
    P, S = ms.init_params(0, randomize_bn=True), ms.synthetic_sample(0)
    eng = Engine(cuda, params=P, precision="x3")
    assert eng.x3
    eng.set_inputs(S)
    eng.forward(train=True)
    o = eng.outputs()
    out, _ = onet.forward(onet.to_torch(P, torch.float64, cuda), onet.to_torch(S, torch.float64, cuda), train=True)
    for k in ("inst_xyz_map_local", "centroids", "lwh", "alpha_bins", "alpha_regs", "cen_z_offs", "cen_y_offs",
              "proj_err_norm", "inst_depth_map_global"):
        a, b = o[k].double().reshape(-1), out[k].reshape(-1)
        assert float((a - b).norm() / b.norm()) < 1e-3, k
    eng.backward()
    eng.optimizer_step()
    torch.cuda.synchronize()
    assert np.isfinite(eng.losses()["total_loss"])


def test_engines_of_both_precisions_alternate_in_one_process(cuda):
    """the operand-rounding switch of the library is re-asserted per engine: a tf32 engine and an x3 engine used in turn
    give the same results as each alone"""
    P, S = ms.init_params(2), ms.synthetic_sample(2)
    a, b = Engine(cuda, params=P, precision="tf32"), Engine(cuda, params=P, precision="x3")
    outs = {}
    for rep in range(2):
        for name, e in (("tf32", a), ("x3", b)):
            e.set_inputs(S)
            e.forward(train=True)
            torch.cuda.synchronize()
            x = e.outputs()["inst_xyz_map_local"].clone()
            if rep == 0:
                outs[name] = x
            else:
                assert torch.equal(outs[name], x), name
    assert not torch.equal(outs["tf32"], outs["x3"])
