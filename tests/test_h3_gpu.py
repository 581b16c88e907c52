"""GPU cases of the fp16-split forward path (csrc/tc_gemm.cu H3 branch, csrc/split16.cu; Engine precision "h3")."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from monopsr_b200 import lib as mlib  # noqa: E402
from monopsr_b200.core import model_spec as ms  # noqa: E402
from monopsr_b200.core.engine import Engine  # noqa: E402
from monopsr_b200.lib_net import TC_FWD, TcGemmParams, W16Layer  # noqa: E402
from oracle import network as onet  # noqa: E402


def _split16(t, b_operand=0, flag=None):
    rows, C = t.reshape(-1, t.shape[-1]).shape
    out = torch.empty_like(t)
    mlib.check(mlib.load().mpb_split16(rows, C, t.data_ptr(), C, out.data_ptr(), C, b_operand,
                                       None if flag is None else flag.data_ptr(), mlib.stream_ptr()), "mpb_split16")
    return out


def _unsplit(s16, b_operand=0):
    """fp64 value hi + lo of every element of a split copy (host-side check of the layout)"""
    h = s16.contiguous().view(torch.float16).reshape(-1, s16.shape[-1] // 32, 2, 32).double()
    v = h[:, :, 0] + h[:, :, 1]
    return v.reshape(s16.shape)


def test_split16_layout_and_accuracy(cuda):
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.randn(37, 96, generator=g) * 4).to(cuda)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    for b_op in (0, 1):
        s = _split16(x, b_op, flag)
        h = s.view(torch.float16).reshape(37, 3, 2, 32)
        hi, lo = (h[:, :, 1], h[:, :, 0]) if b_op else (h[:, :, 0], h[:, :, 1])
        assert torch.equal(hi.reshape(37, 96), x.half())
        assert torch.equal(lo.reshape(37, 96), (x - x.half().float()).half())
        assert float((_unsplit(s) - x.double()).abs().max()) < 4 * 2.0 ** -21
    assert int(flag.item()) == 0
    x[3, 5] = 1e5
    _split16(x, 0, flag)
    assert int(flag.item()) == 1


def _weights16(w2d, gamma=None, var=None, eps=1e-5):
    """mpb_split16_weights_multi on one layer: returns (split copy, inv_scale)"""
    cout, K = w2d.shape
    dev = w2d.device
    w16 = torch.empty_like(w2d)
    inv = torch.empty(cout, device=dev)
    arr = (W16Layer * 1)()
    e = arr[0]
    e.w, e.w16, e.inv_scale = w2d.data_ptr(), w16.data_ptr(), inv.data_ptr()
    e.gamma = gamma.data_ptr() if gamma is not None else None
    e.var = var.data_ptr() if var is not None else None
    e.cout, e.K, e.row0 = cout, K, 0
    tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    r2l = torch.zeros(cout, dtype=torch.int32, device=dev)
    mlib.check(mlib.load().mpb_split16_weights_multi(cout, tab.data_ptr(), r2l.data_ptr(), eps, mlib.stream_ptr()),
               "mpb_split16_weights_multi")
    return w16, inv


def test_weight_split_scales_each_row_into_range(cuda):
    g = torch.Generator(device="cpu").manual_seed(1)
    w = torch.randn(48, 256, generator=g) * torch.logspace(-6, 2, 48).unsqueeze(1)
    w[7] = 0
    w = w.to(cuda)
    gamma, var = (torch.rand(48, generator=g) + 0.5).to(cuda), (torch.rand(48, generator=g) + 0.1).to(cuda)
    w16, inv = _weights16(w, gamma, var)
    s = gamma * torch.rsqrt(var + 1e-5)
    val = _unsplit(w16) * inv.double().unsqueeze(1)
    ref = w.double() * s.double().unsqueeze(1)
    assert float((val - ref).abs().max() / ref.abs().max()) < 1e-6
    rel = ((val - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300))
    assert float(rel[[i for i in range(48) if i != 7]].max()) < 1e-6         # every row, tiny or large, keeps ~22 bits
    mx = (_unsplit(w16).abs().max(dim=1).values)
    ok = (mx >= 2.0 ** 13 * 0.999) & (mx < 2.0 ** 14)
    ok[7] = True
    assert bool(ok.all())
    assert float(inv[7]) == 1.0
    lg = torch.log2(inv)
    assert torch.equal(lg, lg.round())                                        # powers of two


H3_CASES = [
    # nimg, H, W, k, dil, Cin, Cout, BN, ksplit
    (1, 8, 16, 1, 1, 64, 64, 64, 1),          # one tile, two k-blocks (< pipeline depth)
    (2, 12, 12, 1, 1, 256, 128, 128, 1),      # tail rows
    (3, 12, 12, 3, 1, 64, 64, 64, 1),         # im2col, padding taps
    (2, 12, 12, 3, 4, 256, 256, 128, 1),      # atrous, 72 k-blocks, 2 column tiles
    (2, 12, 12, 3, 4, 256, 256, 256, 1),      # widest tile
    (1, 40, 152, 3, 2, 128, 128, 128, 1),     # full-image geometry
    (32, 1, 1, 1, 1, 1088, 1024, 64, 4),      # FC, atomic split-K
]


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN,ksplit", H3_CASES)
def test_h3_gemm_matches_fp64(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    """mpb_tc_gemm_h3 on split copies of UNROUNDED fp32 operands vs an fp64 convolution, with the fused epilogue options
    on, and the split copy it writes of its own result"""
    import torch.nn.functional as F
    g = torch.Generator(device="cpu").manual_seed(5)
    M = nimg * H * W
    x = torch.randn(nimg, H, W, Cin, generator=g).to(cuda)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(cuda)
    shift = torch.randn(Cout, generator=g).to(cuda)
    res = torch.randn(M, Cout, generator=g).to(cuda)
    atomic = ksplit > 1
    out = torch.zeros(M, Cout, device=cuda) if atomic else torch.full((M, Cout), float("nan"), device=cuda)
    out16 = torch.full((M, Cout), float("nan"), device=cuda)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    x16 = _split16(x.reshape(M, Cin))
    w16, inv = _weights16(w.reshape(Cout, k * k * Cin))
    p = TcGemmParams()
    p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout = TC_FWD, H, W, k, k, dil, M, Cin, Cout
    p.ldx, p.ldw, p.out, p.ldo = Cin, k * k * Cin, out.data_ptr(), Cout
    p.X16, p.W16, p.scale, p.overflow = x16.data_ptr(), w16.data_ptr(), inv.data_ptr(), flag.data_ptr()
    p.ksplit, p.atomic = ksplit, 1 if atomic else 0
    if not atomic:
        p.shift, p.res, p.ldr, p.relu = shift.data_ptr(), res.data_ptr(), Cout, 1
        p.out16, p.ldo16 = out16.data_ptr(), Cout
    mlib.check(mlib.load().mpb_tc_gemm_h3(ctypes.byref(p), BN, mlib.stream_ptr()), "mpb_tc_gemm_h3")
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), padding=dil * (k // 2), dilation=dil)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    if not atomic:
        ref = torch.relu(ref + shift.double() + res.double())
    err = float((out.double() - ref).norm() / ref.norm())
    assert err < 2e-5, err          # x3 on the same cases: 2e-6 .. 5e-6; one tf32 pass: ~4e-4
    if not atomic:
        assert float((_unsplit(out16) - out.double()).abs().max()) < 1e-5
        assert int(flag.item()) == 0


def test_h3_rejects_what_it_does_not_implement(cuda):
    from monopsr_b200.lib_net import TC_DGRAD
    x = torch.zeros(128, 64, device=cuda)
    p = TcGemmParams()
    p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout, p.ksplit = TC_FWD, 8, 16, 1, 1, 1, 128, 64, 64, 1
    p.X16, p.ldx, p.W16, p.ldw, p.out, p.ldo = x.data_ptr(), 64, x.data_ptr(), 64, x.data_ptr(), 64
    L = mlib.load()
    assert L.mpb_tc_gemm_h3(ctypes.byref(p), 32, mlib.stream_ptr()) == -1        # tile width
    p.ksplit = 2
    assert L.mpb_tc_gemm_h3(ctypes.byref(p), 64, mlib.stream_ptr()) == -1        # non-atomic split-K
    p.ksplit, p.op = 1, TC_DGRAD
    assert L.mpb_tc_gemm_h3(ctypes.byref(p), 64, mlib.stream_ptr()) == -1        # forward only
    p.op, p.X16 = TC_FWD, None
    assert L.mpb_tc_gemm_h3(ctypes.byref(p), 64, mlib.stream_ptr()) == -1        # needs the split copies


def test_h3_engine_forward_meets_the_parity_bar(cuda):
    """precision="h3": every output of the forward pass within 1e-3 of the fp64 restatement, and a finite training step"""
    P, S = ms.init_params(0, randomize_bn=True), ms.synthetic_sample(0)
    eng = Engine(cuda, params=P, precision="h3")
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
    eng.check_overflow()
    assert np.isfinite(eng.losses()["total_loss"])


def test_h3_overflow_is_reported(cuda):
    P, S = ms.init_params(0), dict(ms.synthetic_sample(0))
    S["rgb_crops"] = np.asarray(S["rgb_crops"]) * 1e6          # drives the first activations beyond fp16's range
    eng = Engine(cuda, params=P, precision="h3")
    eng.set_inputs(S)
    eng.forward(train=True)
    torch.cuda.synchronize()
    with pytest.raises(mlib.MpbError):
        eng.check_overflow()
