"""Host-side tile planning of the GEMM launches (no GPU): Engine._plan_tiles picks (tile width, cluster split-K)
from the output-tile grid and the reduction depth.  The expectations are the configurations that
tools/gemm_sweep.py measured fastest on B200 for the block3 bottleneck shapes (profiles/r1_gemm_sweep.txt)."""
import pytest

from monopsr_b200.core.engine import Engine


def planner(csk=1, csk_bn=128, ctas128=2, fill=0.9, sms=148):
    e = Engine.__new__(Engine)          # no device: only the planning attributes
    e.sms, e.fill, e.csk, e.csk_bn = sms, fill, csk, csk_bn
    e.ctas_per_sm = {64: 2, 128: ctas128, 256: 1}
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
