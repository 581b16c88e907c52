"""KITTI label / calibration readers and label-derived sample fields (monopsr_b200/datasets/kitti_formats.py) against
golden vectors produced by the reference's own functions on three samples of its KITTI test fixture
(tests/golden/make_kitti_formats_golden.py; the label / calib text files live in tests/golden/kitti)."""
import os

import numpy as np
import pytest

from monopsr_b200.datasets import kitti_formats as K

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "kitti_formats_golden.npz"))
LABELS, CALIB = os.path.join(HERE, "golden", "kitti", "label_2"), os.path.join(HERE, "golden", "kitti", "calib")
SAMPLES = ["000001", "000008", "000076"]


@pytest.mark.parametrize("s", SAMPLES)
def test_readers(s):
    labels = K.read_labels(LABELS, s)
    assert [o.type for o in labels] == G[s + "_types"].tolist()
    raw = np.asarray([[o.truncation, o.occlusion, o.alpha, o.x1, o.y1, o.x2, o.y2, o.h, o.w, o.l, o.t[0], o.t[1],
                       o.t[2], o.ry, o.score] for o in labels], np.float64)
    assert np.array_equal(raw, G[s + "_raw"])
    c = K.read_frame_calib(os.path.join(CALIB, s + ".txt"))
    assert np.array_equal(c.p2, G[s + "_p2"]) and np.array_equal(c.r0_rect, G[s + "_r0"])
    assert np.array_equal(c.velo_to_cam, G[s + "_v2c"])
    cars, mask = K.filter_labels_by_class(labels, ["Car"])
    assert mask == G[s + "_car_mask"].tolist() and len(cars) == len(G[s + "_boxes_2d"])


@pytest.mark.parametrize("s", SAMPLES)
def test_boxes_angles_and_bins(s):
    cars, _ = K.filter_labels_by_class(K.read_labels(LABELS, s), ["Car"])
    p2 = K.read_frame_calib(os.path.join(CALIB, s + ".txt")).p2
    for i, o in enumerate(cars):
        b2, b3 = K.object_label_to_box_2d(o), K.object_label_to_box_3d(o)
        assert np.array_equal(b2, G[s + "_boxes_2d"][i]) and np.array_equal(b3, G[s + "_boxes_3d"][i])
        assert K.get_viewing_angle_box_2d(b2, p2) == G[s + "_va2d"][i]
        assert K.get_viewing_angle_box_3d(b3, p2) == G[s + "_va3d"][i]
        np.testing.assert_allclose(K.get_viewing_angle_box_3d(b3, p2, version="projection"), G[s + "_va3d_proj"][i],
                                   rtol=1e-12)
        b, r, v = K.np_orientation_to_angle_bin(o.alpha, 12, 0.0)
        assert b == G[s + "_bins"][i] and np.array_equal(r, G[s + "_regs"][i]) and np.array_equal(v, G[s + "_valid"][i])


def test_angle_bins_with_overlap_and_tables():
    for a, b, r, v in zip(G["ov_angles"], G["ov_bins"], G["ov_regs"], G["ov_valid"]):
        gb, gr, gv = K.np_orientation_to_angle_bin(a, 8, 0.2)
        assert gb == b and np.array_equal(gr, r) and np.array_equal(gv, v)
    for i, c in enumerate(("Car", "Pedestrian", "Cyclist")):
        assert K.get_mean_lwh_and_std_dev(c)[0] == G["mean_lwh"][i].tolist()
        assert K.get_prop_cen_z_offset(c) == G["prop_off"][i]
    assert K.class_str_to_index("Car", ["Car"]) == 1
    for bad in (lambda: K.class_str_to_index("Van", ["Car"]), lambda: K.get_mean_lwh_and_std_dev("Van"),
                lambda: K.get_prop_cen_z_offset("Van"), lambda: K.get_viewing_angle_box_3d(np.zeros(7), None, "nope")):
        with pytest.raises(ValueError):
            bad()
    with pytest.raises(FileNotFoundError):
        K.read_labels(LABELS, "999999")


@pytest.mark.parametrize("s", SAMPLES)
def test_label_fields_are_engine_inputs(s):
    """padded to num_boxes rows under the key names of model_spec.synthetic_sample"""
    from monopsr_b200.core import model_spec as ms
    p2 = K.read_frame_calib(os.path.join(CALIB, s + ".txt")).p2
    f = K.label_fields(K.read_labels(LABELS, s), p2, (375, 1242))
    n = len(G[s + "_boxes_2d"])
    assert f["num_objs"] == n
    ref = ms.synthetic_sample(0)
    for k, v in f.items():
        if k in ref:
            assert v.shape == ref[k].shape and v.dtype == ref[k].dtype, k
    if n:
        assert np.array_equal(f["boxes_2d"][:n], G[s + "_boxes_2d"]) and np.array_equal(f["boxes_3d"][:n], G[s + "_boxes_3d"])
        np.testing.assert_allclose(f["est_view_angs"][:n], G[s + "_va2d"], rtol=1e-6)
        np.testing.assert_allclose(f["gt_view_angs"][:n], G[s + "_va3d"], rtol=1e-6)
        assert np.array_equal(f["gt_alpha_bins"][:n], G[s + "_bins"])
        np.testing.assert_allclose(f["boxes_2d_norm"][:n], G[s + "_boxes_2d"] / np.array([375, 1242, 375, 1242]), rtol=1e-6)
    assert not f["boxes_2d"][n:].any() and not f["gt_alpha_valid_bins"][n:].any()
