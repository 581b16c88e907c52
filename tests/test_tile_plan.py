"""Host-side tile planning of the GEMM launches (no GPU): Engine._plan_tiles picks (tile width, cluster split-K)
from the output-tile grid and the reduction depth.  The expectations are the configurations that
tools/gemm_sweep.py measured fastest on B200 for the block3 bottleneck shapes (profiles/r1_gemm_sweep.txt)."""
import pytest

from monopsr_b200.core.engine import Engine


def planner(csk=1, csk_bn=128, ctas128=2, fill=0.9, sms=148):
    e = Engine.__new__(Engine)          # no device: only the planning attributes
    e.sms, e.fill, e.csk, e.csk_bn = sms, fill, csk, csk_bn
    e.ctas_per_sm = {64: 2, 128: ctas128, 256: 1}
    e.h3 = e.x3 = False
    e.shortk_bn = 64
    return e


MT_FULL, MT_CROPS = 48, 36      # 128-row tiles of the 40x152 full-image grid and of 32 crops x 12 x 12


@pytest.mark.parametrize("mt,ncols,nkb,expect", [
    (MT_FULL, 256, 72, (128, 2)),      # 3x3 256->256: long reduction -> 2-CTA cluster split-K
    (MT_CROPS, 256, 72, (128, 2)),
    (MT_FULL, 256, 32, (128, 2)),      # 1x1 1024->256
    (MT_FULL, 1024, 8, (64, 1)),       # 1x1 256->1024: epilogue-bound, 192 wide tiles would need two waves
    (MT_CROPS, 1024, 8, (256, 1)),     # ... 144 wide tiles fit one wave
    (576, 128, 72, (128, 1)),          # decoder 48x48 maps: plenty of tiles, no split
    (144, 256, 144, (256, 1)),         # decoder 24x24 maps, 512->256
    (1, 1024, 32, (64, 1)),            # FC layers: 32 rows
])
def test_plan(mt, ncols, nkb, expect):
    assert planner()._plan_tiles(mt, ncols, nkb) == expect


def test_plan_without_cluster_splitk():
    e = planner(csk=0)
    assert e._plan_tiles(MT_FULL, 256, 72) == (128, 1)
    assert e._plan_tiles(MT_CROPS, 256, 72)[1] == 1


@pytest.mark.parametrize("mt", [1, 7, 36, 48, 144, 576])
@pytest.mark.parametrize("ncols", [64, 128, 256, 512, 1024, 18432])
@pytest.mark.parametrize("nkb", [2, 8, 16, 32, 72, 576])
def test_plan_is_launchable(mt, ncols, nkb):
    """the tile width divides the column count and every K slice of a cluster gets at least one k-block"""
    bn, ks = planner()._plan_tiles(mt, ncols, nkb)
    assert ncols % bn == 0 and ks in (1, 2)
    per = -(-nkb // ks)
    assert (ks - 1) * per < nkb


def test_x3_mode_dispatch_of_forward_gemms():
    """MPB_PRECISION=x3: forward launches go to mpb_tc_gemm_x3 with no rounding, no rounded second output, no
    cluster split-K and a 128- or 64-wide tile; backward launches keep the single-pass kernel (host logic only)"""
    import torch
    from monopsr_b200.lib_net import TC_DGRAD, TC_FWD

    class Lib(object):
        def __init__(self):
            self.calls = []

        def mpb_tc_gemm(self, p, bn, st):
            q = p._obj
            self.calls.append(("tf32", bn, q.round_tf32, q.out_r, q.ksplit, q.atomic))
            return 0

        def mpb_tc_gemm_x3(self, p, bn, st):
            q = p._obj
            self.calls.append(("x3", bn, q.round_tf32, q.out_r, q.ksplit, q.atomic))
            return 0

    e = planner(csk=1)
    e.csk_fwd, e.x3, e.L, e._record, e._launch_checks = 1, True, Lib(), [], True
    e._st = lambda: None
    e._chk = lambda status, what: None
    x = torch.zeros(6080, 256)
    w, o, o2 = torch.zeros(256, 2304), torch.zeros(6080, 256), torch.zeros(6080, 256)
    e.gemm(TC_FWD, 6080, 40, 152, 3, 2, 256, 256, x, 256, w, 2304, o, 256, relu=1, round_tf32=1, out_r=o2, ldor=256)
    e.gemm(TC_FWD, 6080, 40, 152, 1, 1, 256, 64, x, 256, w, 256, o, 64)
    e.gemm(TC_FWD, 32, 1, 1, 1, 1, 256, 1024, x, 256, w, 256, o, 1024, atomic=1, ksplit=4, bn=64)
    e.gemm(TC_DGRAD, 6080, 40, 152, 3, 2, 256, 256, x, 256, w, 2304, o, 256, round_tf32=1)
    kinds = [c[0] for c in e.L.calls]
    assert kinds == ["x3", "x3", "x3", "tf32"]
    assert e.L.calls[0][1:] == (128, 0, None, 1, 0)          # planner would have asked for a 2-CTA cluster: dropped
    assert e.L.calls[1][1] == 64 and e.L.calls[2][1:] == (128, 0, None, 4, 1)      # atomic split-K is kept
    assert e.L.calls[3][2] == 1 and e.L.calls[3][4] == 2     # the backward launch is planned as before
    assert [(r[1], r[3]) for r in e._record] == [(128, "x3"), (64, "x3"), (128, "x3"), (128, "tf32")]
    e.x3 = False
    e.gemm(TC_FWD, 6080, 40, 152, 3, 2, 256, 256, x, 256, w, 2304, o, 256, relu=1, round_tf32=1, out_r=o2, ldor=256)
    assert e.L.calls[-1][0] == "tf32" and e.L.calls[-1][2] == 1 and e.L.calls[-1][3] == o2.data_ptr()
