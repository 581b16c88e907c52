"""CPU tests of the ground-truth target synthesis oracle (oracle/targets.py) against golden vectors produced by the
reference's own TF-free numpy functions (tests/golden/make_targets_golden.py) and known answers for the two
TensorFlow primitives it restates."""
import os

import numpy as np
import pytest

from oracle import targets as T

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "targets_golden.npz"))


@pytest.mark.parametrize("i", range(6))
def test_pc_map_matches_reference_numpy_twin(i):
    """depth_patch_to_pc_map (depth_map_utils.py:52-126, pixel centres, unrounded box, no correction factors)"""
    pc = T.depth_patch_to_pc_map(G["patches"][i], G["boxes"][i], G["cam_p"], (48, 48))
    np.testing.assert_allclose(pc, G["pc_maps"][i], rtol=2e-6, atol=2e-5)


@pytest.mark.parametrize("i", range(6))
def test_view_norm_matches_reference_numpy_twin(i):
    """apply_view_norm_to_pc_map / np_get_tr_mat (instance_utils.py:512-536)"""
    pc = G["pc_maps"][i].astype(np.float32)
    tm = T.tr_mat(-G["view_angs"][i], -G["centroids"][i])
    p = np.concatenate([pc.reshape(3, -1), np.ones((1, 48 * 48), np.float32)], 0)
    loc = (tm @ p)[0:3].T.reshape(48, 48, 3) * G["valid"][i][..., None]
    np.testing.assert_allclose(loc, G["xyz_local"][i], rtol=1e-5, atol=1e-4)


def test_resize_nearest_align_corners_known_answers():
    a = np.arange(5, dtype=np.float32)[None, :].repeat(5, 0)
    assert T.resize_nearest_align_corners(a, 3, 3)[0].tolist() == [0, 2, 4]           # scale 2
    b = np.arange(4, dtype=np.float32)[None, :].repeat(4, 0)
    assert T.resize_nearest_align_corners(b, 3, 3)[0].tolist() == [0, 2, 3]           # scale 1.5: roundf(1.5) = 2
    c = np.arange(2, dtype=np.float32)[None, :].repeat(2, 0)
    assert T.resize_nearest_align_corners(c, 4, 4)[0].tolist() == [0, 0, 1, 1]        # upsampling, scale 1/3
    one = np.full((1, 1), 7, np.float32)
    assert (T.resize_nearest_align_corners(one, 48, 48) == 7).all()


def test_linspace_and_rounding():
    x = T.tf_linspace(1.0, 2.0, 5)
    assert x.dtype == np.float32 and x[0] == 1.0 and abs(x[-1] - 2.0) < 1e-6 and len(x) == 5
    assert np.rint(np.array([100.5, 101.5, -0.5, 2.4999], np.float32)).tolist() == [100, 102, -0, 2]   # half to even


def _scene(seed, n=5, H=120, W=300):
    rng = np.random.RandomState(seed)
    depth = rng.uniform(2, 60, (H, W)).astype(np.float32)
    masks = rng.rand(n, H, W) < 0.7
    h, w = rng.uniform(10, 60, n), rng.uniform(10, 120, n)
    y1, x1 = rng.uniform(0, H - h), rng.uniform(0, W - w)
    boxes = np.stack([y1, x1, y1 + h, x1 + w], 1).astype(np.float32)
    b3 = np.concatenate([rng.uniform(-10, 10, (n, 1)), rng.uniform(1, 2, (n, 1)), rng.uniform(5, 50, (n, 1)),
                         rng.uniform(1.4, 4.2, (n, 3)), rng.uniform(-3, 3, (n, 1))], 1).astype(np.float32)
    va = rng.uniform(-0.7, 0.7, n).astype(np.float32)
    return depth, masks, boxes, b3, va, G["cam_p"]


def test_gt_maps_structure():
    depth, masks, boxes, b3, va, P = _scene(0)
    loc, glo, val = T.gt_maps(boxes, b3, masks, depth, va, P, roi=48)
    assert loc.shape == glo.shape == (5, 48, 48, 3) and val.shape == (5, 48, 48, 1)
    assert set(np.unique(val)) <= {0.0, 1.0} and 0.5 < val.mean() < 0.9
    inv = val[..., 0] == 0
    assert not loc[inv].any() and not glo[inv].any()               # invalid pixels are exactly (0, 0, 0)
    assert np.all(glo[..., 2][~inv] >= 2.0)                        # global z is the depth itself
    # local = rigid transform of global: distances between valid points are preserved
    for b in range(5):
        pts_l, pts_g = loc[b][~inv[b]], glo[b][~inv[b]]
        dl = np.linalg.norm(pts_l[:50] - pts_l[50:100], axis=1)
        dg = np.linalg.norm(pts_g[:50] - pts_g[50:100], axis=1)
        np.testing.assert_allclose(dl, dg, rtol=1e-3, atol=1e-3)
    # bottom vs middle centroid: only y moves, by h/2
    loc_b, _, _ = T.gt_maps(boxes, b3, masks, depth, va, P, roi=48, centroid_type="bottom")
    dy = (loc - loc_b)[..., 1]
    for b in range(5):
        np.testing.assert_allclose(dy[b][~inv[b]], b3[b, 5] / 2, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ image inputs
def test_image_oracle_against_independent_implementations():
    """align-corners resize against torch's interpolate, crop_and_resize against the torch restatement used for the
    feature-map crops (oracle/network.py), legacy resize against hand-computed values"""
    import torch
    import torch.nn.functional as TF
    from oracle import network as onet
    rng = np.random.RandomState(1)
    img = rng.uniform(-100, 150, (37, 53, 3)).astype(np.float32)
    ac = T.resize_bilinear_ac(img, (20, 31))
    ref = TF.interpolate(torch.from_numpy(img).permute(2, 0, 1)[None].double(), size=(20, 31), mode="bilinear",
                         align_corners=True)[0].permute(1, 2, 0).numpy()
    np.testing.assert_allclose(ac, ref, rtol=1e-5, atol=1e-3)      # fp32 lerps of +-150 values vs an fp64 reference
    # (a box ending exactly at 1.0 is avoided here: whether its last row is "inside" depends on the last fp32 bit of
    # y1*(H-1) + i*scale, which the fp32 oracle reproduces as TF computes it and an fp64 reference does not)
    boxes = np.array([[0.1, 0.2, 0.6, 0.9], [0.0, 0.0, 0.98, 0.97], [-0.1, 0.5, 0.4, 1.23]], np.float32)   # last one leaves the image
    cr = T.crop_and_resize(img, boxes, 8)
    ref = onet.crop_and_resize(torch.from_numpy(img)[None].double(), torch.from_numpy(boxes).double(), 8, 8).numpy()
    np.testing.assert_allclose(cr, ref, rtol=1e-5, atol=1e-3)
    assert (cr[2][0] == 0).all()                                   # rows above the image: extrapolation value 0
    # legacy (align_corners=False, no half-pixel) bilinear: 4 -> 2 samples source pixels 0 and 2 exactly
    u8 = np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 4, 3)
    pre = T.preprocess_input(u8, (2, 2), means=np.zeros(3, np.float32))
    assert np.array_equal(pre, u8[::2, ::2].astype(np.float32))
    same = T.preprocess_input(u8, (4, 4))
    np.testing.assert_allclose(same, u8.astype(np.float32) - T.KITTI_CHANNEL_MEANS, rtol=0, atol=1e-5)


def test_against_the_reference_tf_functions_executed_on_arrays():
    """oracle.targets.instance_xyz_crop_from_depth_map against the reference's TENSORFLOW version of the same function,
    unmodified, executed through the numpy-backed TF stand-in (tests/golden/make_targets_tf_golden.py): box rounding,
    mask, crop, nearest-neighbour resize, back-projection at pixel centres, centroid adjustment, view normalisation,
    valid-pixel threshold.  The stand-in computes in float64, the oracle in float32 (as TF would): tolerance 1e-4 on
    values of order 10, exact on the validity masks."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_targets_tf_golden", os.path.join(here, "golden", "make_targets_tf_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    depth, masks, b2, b3, view, cam_p = gen.inputs()
    G = np.load(os.path.join(here, "golden", "targets_tf_golden.npz"))
    for name, view_norm, ctype, rot in (("local", True, "middle", True), ("global", False, "middle", True),
                                        ("local_bottom_norot", True, "bottom", False)):
        for i in range(4):
            xyz, valid = T.instance_xyz_crop_from_depth_map(i, b2, b3, masks, depth, (48, 48), view, cam_p, view_norm,
                                                            centroid_type=ctype, rotate_view=rot)
            assert np.array_equal(valid, G["valid_" + name][i]), (name, i)
            np.testing.assert_allclose(xyz, G["xyz_" + name][i], rtol=0, atol=2e-4, err_msg="%s %d" % (name, i))
    assert 0.3 < G["valid_local"].mean() < 0.9
