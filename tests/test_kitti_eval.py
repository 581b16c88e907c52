"""KITTI AP evaluator (csrc/kitti_eval.cu + core/kitti_eval.py) against the reference's own evaluator: golden output of
the binary compiled from the reference source (tests/golden/make_kitti_eval_golden.py) on five synthetic detectors,
normal and low-IoU thresholds -- every printed AP line and every stats_*.txt curve -- plus, where oracle/_ref holds
the binary, a live comparison on a fresh case; and the box overlaps against closed forms."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kitti_eval_cases  # noqa: E402
from monopsr_b200.core import kitti_eval as E  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "kitti_eval_golden.json")))
REF_BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "evaluate_object_3d_offline")


def _check(res, res_dir, want):
    assert res["lines"] == want["lines"]                     # same classes / metrics, same order, same 6 decimals
    for fname, rows in want["stats"].items():
        got = [[float(v) for v in ln.split()] for ln in open(os.path.join(res_dir, fname)).read().splitlines()]
        assert np.asarray(got).shape == np.asarray(rows).shape, fname
        np.testing.assert_allclose(got, rows, rtol=0, atol=1.01e-6, err_msg=fname)
    assert sorted(f for f in os.listdir(res_dir) if f.startswith("stats_")) == sorted(want["stats"])


@pytest.mark.parametrize("low_iou", [False, True])
@pytest.mark.parametrize("case", list(kitti_eval_cases.CASES))
def test_matches_the_reference_evaluator(tmp_path, case, low_iou):
    gt_dir, res_dir = kitti_eval_cases.make_case(str(tmp_path), case)
    res = E.evaluate(gt_dir, res_dir, low_iou=low_iou, write_stats=True)
    _check(res, res_dir, GOLD[case + ("/low_iou" if low_iou else "")])
    if case == "no_alpha":
        assert not any("orientation" in ln for ln in res["lines"])         # alpha = -10 switches AOS off
    if case == "perfect":
        assert res["ap"]["pedestrian_detection"] == [100.0, 100.0, 100.0]
        assert res["ap"]["pedestrian_detection_3D"] == res["ap"]["pedestrian_heading_3D"] == [100.0, 100.0, 100.0]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref evaluator not built (bash oracle/build_ref.sh)")
def test_live_against_the_reference_binary(tmp_path):
    kitti_eval_cases.CASES["live"] = dict(seed=77, frames=40, p_detect=0.85, box_noise=0.08, pose_noise=0.2, n_false=2,
                                          alpha=True)
    try:
        gt_dir, res_dir = kitti_eval_cases.make_case(str(tmp_path), "live")
    finally:
        del kitti_eval_cases.CASES["live"]
    out = subprocess.run([REF_BIN, gt_dir + "/", res_dir], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         universal_newlines=True).stdout
    want = {"lines": [ln for ln in out.splitlines() if " AP: " in ln], "stats": {}}
    for f in os.listdir(res_dir):
        if f.startswith("stats_"):
            want["stats"][f] = [[float(v) for v in ln.split()] for ln in open(os.path.join(res_dir, f)).read().splitlines()]
            os.remove(os.path.join(res_dir, f))
    assert len(want["lines"]) >= 12
    _check(E.evaluate(gt_dir, res_dir, write_stats=True), res_dir, want)


def _box(x, z, l, w, ry, y=1.5, h=1.5):
    r = np.zeros(16)
    r[8:15] = [h, w, l, x, y, z, ry]
    return r


def test_overlaps_closed_forms():
    a, b = np.zeros(16), np.zeros(16)
    a[4:8], b[4:8] = [0, 0, 10, 10], [5, 0, 15, 10]
    assert E.overlap(a, b, E.IMAGE) == pytest.approx(50 / 150)
    assert E.overlap(a, b, E.IMAGE, 0) == pytest.approx(0.5) and E.overlap(a, b, E.IMAGE, 1) == pytest.approx(0.5)
    b[4:8] = [10, 0, 20, 10]
    assert E.overlap(a, b, E.IMAGE) == 0                                          # touching boxes do not overlap
    g = _box(0, 10, 4, 2, 0.0)
    assert E.overlap(_box(0, 10, 4, 2, 0.0), g, E.GROUND) == pytest.approx(1.0)
    assert E.overlap(_box(0, 10, 4, 2, np.pi / 2), g, E.GROUND) == pytest.approx(4 / 12)      # a cross: 2 x 2 in common
    assert E.overlap(_box(2, 10, 4, 2, 0.0), g, E.GROUND) == pytest.approx(4 / 12)
    assert E.overlap(_box(2, 10, 4, 2, np.pi), g, E.GROUND) == pytest.approx(4 / 12)          # heading flip: same footprint
    assert E.overlap(_box(50, 10, 4, 2, 0.3), g, E.GROUND) == 0
    # 45 degrees: a square of side s rotated inside an equal square shares an octagon of area 2 (sqrt(2) - 1) s^2
    sq = _box(0, 10, 2, 2, 0.0)
    oct_area = 2 * (np.sqrt(2) - 1) * 4
    assert E.overlap(_box(0, 10, 2, 2, np.pi / 4), sq, E.GROUND) == pytest.approx(oct_area / (8 - oct_area))
    assert E.overlap(_box(0, 10, 2, 2, np.pi / 4), sq, E.GROUND, 0) == pytest.approx(oct_area / 4)
    # 3-D: footprint x overlap of the vertical extents [y - h, y]
    assert E.overlap(_box(0, 10, 4, 2, 0.0, y=1.5, h=1.5), _box(0, 10, 4, 2, 0.0, y=2.25, h=1.5), E.BOX3D) == \
        pytest.approx((8 * 0.75) / (12 + 12 - 6))
    assert E.overlap(_box(0, 10, 4, 2, 0.0, y=1.5), _box(0, 10, 4, 2, 0.0, y=4.0), E.BOX3D) == 0
    # random rotated rectangles against a Monte-Carlo estimate
    rs = np.random.RandomState(0)
    for _ in range(5):
        d = _box(rs.uniform(-1, 1), 10 + rs.uniform(-1, 1), rs.uniform(2, 5), rs.uniform(1, 3), rs.uniform(-3, 3))
        g = _box(rs.uniform(-1, 1), 10 + rs.uniform(-1, 1), rs.uniform(2, 5), rs.uniform(1, 3), rs.uniform(-3, 3))
        pts = rs.uniform(-6, 6, (400000, 2)) + [0, 10]

        def inside(b):
            c, s = np.cos(b[14]), np.sin(b[14])          # corners = R * local + t with R = [[c, s], [-s, c]]
            dx, dz = pts[:, 0] - b[11], pts[:, 1] - b[13]
            lx, lz = c * dx - s * dz, s * dx + c * dz     # R^T * (p - t)
            return (np.abs(lx) <= b[10] / 2) & (np.abs(lz) <= b[9] / 2)
        i_d, i_g = inside(d), inside(g)
        mc = (i_d & i_g).sum() / max((i_d | i_g).sum(), 1)
        assert abs(E.overlap(d, g, E.GROUND) - mc) < 0.01


def test_parsing_and_argument_checks(tmp_path):
    p = tmp_path / "000003.txt"
    p.write_text("Car 0.00 1 1.5 10 20 50 70 1.5 1.6 3.9 1 2 30 0.1\r\nTram 0.1 0 0 1 2 3 4 5 6 7 8 9 10 11\nbroken line\n")
    gt = E.parse_objects(str(p), False)
    assert gt.shape == (2, 15) and gt[0, 0] == 0 and gt[1, 0] == E.OTHER and gt[0, 2] == 1
    (tmp_path / "data").mkdir()
    (tmp_path / "data" / "000003.txt").write_text("car -1 -1 0.5 10 20 50 70 1.5 1.6 3.9 1 2 30 0.1 0.9\n")
    (tmp_path / "data" / "x.txt").write_text("")
    assert E.eval_indices(str(tmp_path)) == [3]
    det = E.parse_objects(str(tmp_path / "data" / "000003.txt"), True)
    assert det.shape == (1, 16) and det[0, 0] == 0 and det[0, 15] == 0.9        # type names are case-insensitive
    with pytest.raises(ValueError):
        E.eval_class([gt], [], 0, 0, 0, 0.7)
    with pytest.raises(FileNotFoundError):
        E.evaluate(str(tmp_path / "nowhere"), str(tmp_path))
    r = E.eval_class([gt], [det], 0, 2, E.IMAGE, 0.7, compute_aos=True)
    assert r["n_gt"] == 1 and r["n_thresholds"] == 1 and r["precision"][0] == 1.0 and r["precision"][1] == 0.0
    assert r["aos"][0] == pytest.approx((1 + np.cos(1.5 - 0.5)) / 2)
    empty = E.eval_class([], [], 0, 0, 0, 0.7)
    assert empty["n_gt"] == 0 and not empty["precision"].any()
    assert E.average_precision(np.ones(41)) == 100.0


def test_command_line(tmp_path, capsys):
    gt_dir, res_dir = kitti_eval_cases.make_case(str(tmp_path), "sparse")
    assert E.main([gt_dir, res_dir]) == 0
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "results" and out[1:] == GOLD["sparse"]["lines"]
    assert E.main(["--low_iou", gt_dir, res_dir]) == 0
    assert capsys.readouterr().out.splitlines()[1:] == GOLD["sparse/low_iou"]["lines"]
    assert E.main([gt_dir]) == 1
