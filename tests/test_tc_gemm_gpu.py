"""GPU tests of the tcgen05 implicit-GEMM core against torch fp64 convolution.

Test-side reference: torch.nn.functional.conv2d in float64 (the fp64 restatement of
slim.conv2d with SAME padding / atrous rate).  tcgen05 kind::tf32 keeps 10 mantissa
bits of each operand, so the tolerance is the tf32 bound |err| <= ~2^-10 * sum|a||b|
(inputs here are pre-rounded to tf32, making the product exact and leaving only fp32
accumulation error -> tight tolerance).
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from monopsr_b200 import lib as mlib  # noqa: E402
from monopsr_b200.lib_net import TC_DGRAD, TC_FWD, TC_WGRAD, TcGemmParams  # noqa: E402


@pytest.fixture(autouse=True, params=["tma", "tma_cluster", "cpasync"])
def producer(request):
    """run every case with all operand-staging variants of the kernel: TMA, TMA with the A tile
    multicast over a thread-block cluster of column tiles, and cp.async (LSU) producers"""
    L = mlib.load()
    mode = L.mpb_tc_set_producer(0 if request.param == "cpasync" else 1)
    if request.param != "cpasync" and mode != 1:
        pytest.skip("driver lacks cuTensorMapEncode*")
    L.mpb_tc_set_cluster(4 if request.param == "tma_cluster" else 1)
    yield request.param
    L.mpb_tc_set_producer(1)
    L.mpb_tc_set_cluster(1)


def tf32_round(t):
    """round-to-nearest-even-ish to 10 mantissa bits (matches cvt.rna up to ties)"""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def tapmask(nimg, H, W, k, dil, dev):
    out = torch.empty(nimg * H * W, dtype=torch.int16, device=dev)
    st = mlib.load().mpb_build_tapmask(nimg, H, W, k, k, dil, out.data_ptr(), mlib.stream_ptr())
    mlib.check(st, "tapmask")
    return out


def run(p, BN):
    st = mlib.load().mpb_tc_gemm(ctypes.byref(p), BN, mlib.stream_ptr())
    mlib.check(st, "mpb_tc_gemm")
    torch.cuda.synchronize()


def base_params(op, nimg, H, W, k, dil, Cin, Cout):
    p = TcGemmParams()
    p.op, p.H, p.W, p.kh, p.kw, p.dil = op, H, W, k, k, dil
    p.M, p.Cin, p.Cout = nimg * H * W, Cin, Cout
    p.ksplit = 1
    return p


CASES = [
    # nimg, H, W, k, dil, Cin, Cout, BN
    (1, 8, 16, 1, 1, 64, 64, 64),        # exactly one 128-row tile, plain GEMM
    (2, 12, 12, 1, 1, 256, 128, 128),    # M=288: tail rows
    (3, 12, 12, 3, 1, 64, 64, 64),
    (2, 12, 12, 3, 2, 128, 128, 64),
    (2, 12, 12, 3, 4, 256, 256, 256),
    (1, 40, 152, 3, 4, 64, 64, 64),      # full-image geometry
    (2, 12, 12, 3, 2, 128, 256, 64),     # 4 column tiles -> cluster of 4, multicast im2col A
    (1, 40, 152, 3, 4, 256, 128, 64),    # 2 column tiles -> cluster of 2; DGRAD: 4 tiles
    (3, 12, 12, 1, 1, 128, 512, 128),    # 1x1 with 4 column tiles (tiled multicast)
    (32, 1, 1, 1, 1, 1056, 1024, 128),   # FC: 32 rows
]


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN", CASES)
def test_fwd(cuda, nimg, H, W, k, dil, Cin, Cout, BN):
    _fwd(cuda, nimg, H, W, k, dil, Cin, Cout, BN, 1)


# cluster split-K (ksplit > 1, atomic = 0): the K slices of a tile form one thread-block cluster and are reduced
# through distributed shared memory inside the fused epilogue; output is written once, NaN-prefilled here
CSK_CASES = [
    # nimg, H, W, k, dil, Cin, Cout, BN, ksplit
    (2, 12, 12, 1, 1, 256, 128, 128, 2),
    (2, 12, 12, 3, 4, 256, 256, 256, 3),    # 8 chunks over 3 owners (3,3,2)
    (3, 12, 12, 3, 1, 64, 64, 64, 4),       # 2 chunks, 4 CTAs: two CTAs own nothing
    (1, 40, 152, 3, 4, 256, 128, 64, 3),
    (1, 40, 152, 1, 1, 1024, 256, 256, 8),
    (3, 12, 12, 1, 1, 128, 512, 128, 4),    # K = 4 k-blocks, one per CTA
]


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN,ksplit", CSK_CASES)
def test_fwd_fused_splitk(cuda, producer, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    if producer != "tma":
        pytest.skip("fused split-K exists for the TMA kernel only")
    _fwd(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit)


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN,ksplit", CSK_CASES[:5])
def test_dgrad_fused_splitk(cuda, producer, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    if producer != "tma":
        pytest.skip("fused split-K exists for the TMA kernel only")
    _dgrad(cuda, nimg, H, W, k, dil, Cout, Cin, BN, ksplit)     # (Cin, Cout) swapped: N = Cin must divide by BN


def test_cluster_splitk_rejects_empty_slice(cuda):
    p = base_params(TC_FWD, 1, 8, 16, 1, 1, 64, 64)     # 2 k-blocks cannot feed 3 slices
    x = torch.zeros(128, 64, device=cuda)
    p.X, p.ldx, p.Wt, p.ldw, p.out, p.ldo = x.data_ptr(), 64, x.data_ptr(), 64, x.data_ptr(), 64
    p.ksplit = 3
    assert mlib.load().mpb_tc_gemm(ctypes.byref(p), 64, mlib.stream_ptr()) == -1


def _fwd(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = tf32_round(torch.randn(nimg, H, W, Cin, generator=g)).to(cuda)
    w = tf32_round(torch.randn(Cout, k, k, Cin, generator=g) * 0.1).to(cuda)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(cuda)
    shift = torch.randn(Cout, generator=g).to(cuda)
    res = torch.randn(nimg, H, W, Cout, generator=g).to(cuda)
    out = torch.full((nimg, H, W, Cout), float("nan"), device=cuda)
    p = base_params(TC_FWD, nimg, H, W, k, dil, Cin, Cout)
    p.X, p.ldx = x.data_ptr(), Cin
    p.Wt, p.ldw = w.data_ptr(), k * k * Cin
    p.out, p.ldo = out.data_ptr(), Cout
    tm = tapmask(nimg, H, W, k, dil, cuda) if k > 1 else None
    p.tapmask = tm.data_ptr() if tm is not None else None
    p.scale, p.shift, p.res, p.ldr, p.relu = scale.data_ptr(), shift.data_ptr(), res.data_ptr(), Cout, 1
    p.ksplit = ksplit
    run(p, BN)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), padding=dil * (k // 2),
                   dilation=dil).permute(0, 2, 3, 1)
    ref = torch.relu(ref * scale.double() + shift.double() + res.double())
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN", CASES[:9])
def test_dgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN):
    _dgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN, 1)


def _dgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    g = torch.Generator(device="cpu").manual_seed(2)
    dy = tf32_round(torch.randn(nimg, H, W, Cout, generator=g)).to(cuda)
    w = tf32_round(torch.randn(Cout, k, k, Cin, generator=g) * 0.1).to(cuda)
    res = torch.randn(nimg, H, W, Cin, generator=g).to(cuda)
    mask = torch.randn(nimg, H, W, Cin, generator=g).to(cuda)
    s2 = (torch.rand(Cin, generator=g) + 0.5).to(cuda)
    out = torch.full((nimg, H, W, Cin), float("nan"), device=cuda)
    colsum = torch.zeros(Cin, device=cuda)
    p = base_params(TC_DGRAD, nimg, H, W, k, dil, Cin, Cout)
    p.X, p.ldx = dy.data_ptr(), Cout
    p.Wt, p.ldw = w.data_ptr(), k * k * Cin
    p.out, p.ldo = out.data_ptr(), Cin
    tm = tapmask(nimg, H, W, k, dil, cuda) if k > 1 else None
    p.tapmask = tm.data_ptr() if tm is not None else None
    p.res, p.ldr, p.mask, p.ldm, p.scale2 = res.data_ptr(), Cin, mask.data_ptr(), Cin, s2.data_ptr()
    p.colsum = colsum.data_ptr()
    p.ksplit = ksplit
    run(p, BN)
    xx = torch.zeros(nimg, Cin, H, W, dtype=torch.float64, device=cuda, requires_grad=True)
    yy = F.conv2d(xx, w.double().permute(0, 3, 1, 2), padding=dil * (k // 2), dilation=dil)
    yy.backward(dy.double().permute(0, 3, 1, 2))
    ref = (xx.grad.permute(0, 2, 3, 1) + res.double()) * (mask > 0).double() * s2.double()
    err = (out.double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
    cerr = (colsum.double() - ref.sum((0, 1, 2))).abs().max().item()
    assert cerr < 1e-3 * max(1.0, ref.sum((0, 1, 2)).abs().max().item()), cerr


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN,ksplit", [
    (1, 8, 16, 1, 1, 64, 128, 64, 1), (2, 12, 12, 1, 1, 256, 128, 128, 2), (3, 12, 12, 3, 1, 64, 64, 64, 3),
    (2, 12, 12, 3, 2, 128, 128, 64, 1), (4, 12, 12, 3, 4, 256, 256, 128, 4), (1, 40, 152, 3, 4, 64, 64, 64, 8),
    (32, 1, 1, 1, 1, 1088, 1024, 64, 1),
])
def test_wgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    _wgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit, 1)


@pytest.mark.parametrize("nimg,H,W,k,dil,Cin,Cout,BN,ksplit", [
    (2, 12, 12, 1, 1, 256, 128, 128, 2), (4, 12, 12, 3, 4, 256, 256, 256, 4), (1, 40, 152, 3, 4, 64, 64, 64, 8),
])
def test_wgrad_fused_splitk(cuda, producer, nimg, H, W, k, dil, Cin, Cout, BN, ksplit):
    if producer != "tma":
        pytest.skip("fused split-K exists for the TMA kernel only")
    _wgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit, 0)


def _wgrad(cuda, nimg, H, W, k, dil, Cin, Cout, BN, ksplit, atomic):
    g = torch.Generator(device="cpu").manual_seed(3)
    x = tf32_round(torch.randn(nimg, H, W, Cin, generator=g)).to(cuda)
    dy = tf32_round(torch.randn(nimg, H, W, Cout, generator=g)).to(cuda)
    dw = torch.zeros(Cout, k, k, Cin, device=cuda) if atomic else torch.full((Cout, k, k, Cin), float("nan"), device=cuda)
    p = base_params(TC_WGRAD, nimg, H, W, k, dil, Cin, Cout)
    p.X, p.ldx = x.data_ptr(), Cin
    p.Y, p.ldy = dy.data_ptr(), Cout
    p.Wt, p.ldw = None, k * k * Cin
    p.out = dw.data_ptr()
    tm = tapmask(nimg, H, W, k, dil, cuda) if k > 1 else None
    p.tapmask = tm.data_ptr() if tm is not None else None
    p.atomic, p.ksplit = atomic, ksplit
    run(p, BN)
    ww = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, device=cuda, requires_grad=True)
    yy = F.conv2d(x.double().permute(0, 3, 1, 2), ww, padding=dil * (k // 2), dilation=dil)
    yy.backward(dy.double().permute(0, 3, 1, 2))
    ref = ww.grad.permute(0, 2, 3, 1)
    err = (dw.double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
