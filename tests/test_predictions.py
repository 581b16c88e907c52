"""Prediction formatting / scoring / writers (monopsr_b200/core/predictions.py) against golden vectors computed by the
reference's own numpy functions (tests/golden/make_predictions_golden.py) and structural checks of
format_predictions / save_predictions."""
import os

import numpy as np
import pytest

from monopsr_b200.core import predictions as P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "predictions_golden.npz"))


def test_corners_projection_and_orientation():
    for b, c in zip(G["boxes_3d"], G["corners"]):
        np.testing.assert_allclose(P.compute_box_3d_corners(b), c, rtol=1e-12, atol=1e-12)
    got = [P.np_angle_bin_to_orientation(b, r, 12) for b, r in zip(G["ang_bins"], G["ang_res"])]
    np.testing.assert_allclose(got, G["ang"], rtol=0, atol=1e-12)
    assert -np.pi <= min(got) and max(got) <= np.pi


def test_postprocess_cen_x_and_projection():
    got = [np.squeeze(P.postprocess_cen_x(b2, b3, G["cam_p"])) for b2, b3 in zip(G["boxes_2d"], G["boxes_3d"])]
    np.testing.assert_allclose(got, G["cen_x"], rtol=1e-10, atol=1e-10)
    for b, want in zip(G["boxes_3d"], G["proj_trunc"]):
        r = P.project_to_image_space(b, G["cam_p"], truncate=True, image_size=(1242, 375))
        if np.isnan(want[0]):
            assert r is None
        else:
            np.testing.assert_allclose(r, want, rtol=1e-10, atol=1e-9)
    with pytest.raises(ValueError):
        P.project_to_image_space(G["boxes_3d"][0], G["cam_p"], truncate=True)


def test_score_boxes():
    got = P.score_boxes(G["cam_p"], (375, 1242), G["boxes_2d"], G["boxes_3d"], G["scores"])
    np.testing.assert_allclose(got, G["new_scores"], rtol=1e-10, atol=1e-10)


def _sample(n_boxes=32, num_objs=5, seed=0):
    rng = np.random.RandomState(seed)
    b3 = np.zeros((n_boxes, 7), np.float32)
    b3[:24] = G["boxes_3d"]
    b3[24:] = G["boxes_3d"][:8]
    b2 = np.concatenate([G["boxes_2d"], G["boxes_2d"][:8]]).astype(np.float32)
    sample = {P.SAMPLE_NAME: "000042", P.SAMPLE_IMAGE_INPUT: np.zeros((375, 1242, 3), np.uint8),
              P.SAMPLE_NUM_OBJS: num_objs, P.SAMPLE_CAM_P: G["cam_p"].astype(np.float32),
              P.SAMPLE_LABEL_SCORES: rng.uniform(0.2, 1, n_boxes).astype(np.float32),
              P.SAMPLE_LABEL_BOXES_2D: b2, P.SAMPLE_LABEL_BOXES_3D: b3,
              P.SAMPLE_VIEWING_ANGLES_3D: np.arctan2(b3[:, 0], b3[:, 2])}
    out = {P.KEY_VALID_MASK_MAPS: rng.standard_normal((n_boxes, 48, 48, 1)).astype(np.float32),
           P.KEY_INST_XYZ_MAP_LOCAL: rng.standard_normal((n_boxes, 48, 48, 3)).astype(np.float32),
           P.KEY_LWH: b3[:, 3:6] + 0.1, P.KEY_VIEW_ANG: np.arctan2(b3[:, 0], b3[:, 2])[:, None],
           P.KEY_ALPHA_BINS: rng.standard_normal((n_boxes, 12)).astype(np.float32),
           P.KEY_ALPHA_REGS: rng.uniform(-0.2, 0.2, (n_boxes, 12)).astype(np.float32),
           P.KEY_CENTROIDS: b3[:, 0:3] - np.array([0, 0.8, 0], np.float32),
           P.SAMPLE_LABEL_CLASS_INDICES: np.ones((n_boxes, 1), np.int32)}
    return sample, out


TYPES = [P.KEY_INST_XYZ_MAP_LOCAL, P.KEY_VALID_MASK_MAPS, P.KEY_CENTROIDS, P.KEY_LWH, P.KEY_VIEW_ANG, P.KEY_ALPHA]


def test_format_predictions_structure():
    sample, out = _sample()
    cen_before = out[P.KEY_CENTROIDS].copy()
    pred = P.format_predictions(out, sample, output_types=TYPES)
    n = sample[P.SAMPLE_NUM_OBJS]
    assert pred[P.KEY_BOX_3D].shape == (n, 9) and pred[P.KEY_BOX_2D].shape == (n, 7)
    assert set(np.unique(pred[P.KEY_VALID_MASK_MAPS])) <= {0.0, 1.0}
    assert not pred[P.KEY_INST_XYZ_MAP_LOCAL][pred[P.KEY_VALID_MASK_MAPS][..., 0] == 0].any()
    b3 = pred[P.KEY_BOX_3D]
    np.testing.assert_allclose(b3[:, 3:6], out[P.KEY_LWH][:n], rtol=1e-6)
    # 'middle' centroid -> KITTI bottom-centre y: + h/2
    np.testing.assert_allclose(b3[:, 1], cen_before[:n, 1] + out[P.KEY_LWH][:n, 2] / 2, rtol=1e-5)
    np.testing.assert_array_equal(out[P.KEY_CENTROIDS], cen_before)          # the caller's array is not modified
    bins = np.argmax(out[P.KEY_ALPHA_BINS][:n], 1)
    alpha = [P.np_angle_bin_to_orientation(b, out[P.KEY_ALPHA_REGS][i, b], 12) for i, b in enumerate(bins)]
    np.testing.assert_allclose(pred[P.KEY_BOX_2D][:, 4], alpha, rtol=1e-6)
    np.testing.assert_allclose(b3[:, 6], np.asarray(alpha) + out[P.KEY_VIEW_ANG][:n, 0], rtol=1e-6)
    assert (b3[:, 8] == 0).all() and (b3[:, 7] > 0).all()                    # class index 'Car' -> 0, scores
    # x re-derived from the 2-D box (postprocess_cen_x) unless switched off
    raw = P.format_predictions(out, sample, output_types=TYPES, post_process_cen_x=False)
    np.testing.assert_allclose(raw[P.KEY_BOX_3D][:, 0], cen_before[:n, 0], rtol=1e-6)
    test = P.format_predictions(out, sample, output_types=TYPES, train_val_test="test")
    assert test[P.KEY_BOX_3D].shape == (n, 9)
    with pytest.raises(ValueError):
        P.format_predictions(out, sample, output_types=TYPES, train_val_test="bogus")


def test_save_predictions_roundtrip(tmp_path):
    sample, out = _sample()
    pred = P.format_predictions(out, sample, output_types=TYPES)
    dirs = {P.OUT_DIR_XYZ_MAP_LOCAL: str(tmp_path / "xyz"), P.OUT_DIR_BOX_2D: str(tmp_path / "b2"),
            P.OUT_DIR_BOX_3D: str(tmp_path / "b3")}
    for d in dirs.values():
        os.makedirs(d)
    P.save_predictions("000042", pred, dirs, TYPES)
    xyz = np.load(os.path.join(dirs[P.OUT_DIR_XYZ_MAP_LOCAL], "000042.npy"))
    assert xyz.dtype == np.float16 and xyz.shape == pred[P.KEY_INST_XYZ_MAP_LOCAL].shape
    b3 = np.loadtxt(os.path.join(dirs[P.OUT_DIR_BOX_3D], "000042.txt"))
    np.testing.assert_allclose(b3, pred[P.KEY_BOX_3D], atol=1e-5)
    b2 = np.loadtxt(os.path.join(dirs[P.OUT_DIR_BOX_2D], "000042.txt"))
    assert b2.shape == pred[P.KEY_BOX_2D].shape
