"""GPU parity of the ground-truth target synthesis kernel (csrc/targets.cu) against the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

from monopsr_b200.core import targets as mt  # noqa: E402
from oracle import targets as T  # noqa: E402
from test_targets_oracle import G, _scene  # noqa: E402


def run(scene, **kw):
    depth, masks, boxes, b3, va, P = scene
    out = mt.gt_maps_from_depth(depth, masks, boxes, b3, va, P, "cuda:0", **kw)
    torch.cuda.synchronize()
    return [out[k].cpu().numpy() for k in ("gt_inst_xyz_maps_local", "gt_inst_xyz_maps_global", "gt_valid_mask_maps")]


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("centroid_type,rotate_view", [("middle", True), ("bottom", True), ("middle", False)])
def test_kernel_matches_oracle(cuda, seed, centroid_type, rotate_view):
    scene = _scene(seed, n=7, H=150, W=400)
    loc, glo, val = run(scene, centroid_type=centroid_type, rotate_view=rotate_view)
    depth, masks, boxes, b3, va, P = scene
    eloc, eglo, eval_ = T.gt_maps(boxes, b3, masks, depth, va, P, roi=48, centroid_type=centroid_type,
                                  rotate_view=rotate_view)
    assert np.array_equal(val, eval_)                    # sampling grid + mask: exact
    assert np.array_equal(glo, eglo)                     # element-wise fp32 chain in the reference's order: exact
    # the 4x4 matmul of the view normalisation is a BLAS call in numpy (and in TF): tolerance 1e-5 relative
    np.testing.assert_allclose(loc, eloc, rtol=1e-5, atol=2e-5)


def test_kitti_size_and_edge_boxes(cuda):
    rng = np.random.RandomState(5)
    H, W, n = 375, 1242, 32
    depth = rng.uniform(1, 80, (H, W)).astype(np.float32)
    depth[rng.rand(H, W) < 0.1] = 0.05                   # below the 0.1 validity threshold
    masks = rng.rand(n, H, W) < 0.6
    masks[3] = False                                     # an instance with an empty mask: nothing valid
    boxes = np.stack([rng.uniform(0, 200, n), rng.uniform(0, 900, n), np.zeros(n), np.zeros(n)], 1).astype(np.float32)
    boxes[:, 2] = boxes[:, 0] + rng.uniform(2, 170, n)
    boxes[:, 3] = boxes[:, 1] + rng.uniform(2, 340, n)
    boxes[0] = [0, 0, 375, 1242]                         # the whole image
    boxes[1] = [100.5, 200.5, 101.5, 202.5]              # rounds (half to even) to a 2 x 2 pixel crop: 100..102, 200..202
    boxes[2] = [370.2, 1230.7, 375, 1242]                # touches the bottom-right corner
    b3 = np.concatenate([rng.uniform(-20, 20, (n, 1)), rng.uniform(1, 2, (n, 1)), rng.uniform(3, 70, (n, 1)),
                         rng.uniform(1.4, 4.2, (n, 3)), rng.uniform(-3, 3, (n, 1))], 1).astype(np.float32)
    va = rng.uniform(-0.8, 0.8, n).astype(np.float32)
    scene = (depth, masks, boxes, b3, va, G["cam_p"])
    loc, glo, val = run(scene)
    eloc, eglo, eval_ = T.gt_maps(boxes, b3, masks, depth, va, G["cam_p"], roi=48)
    assert np.array_equal(val, eval_) and np.array_equal(glo, eglo)
    np.testing.assert_allclose(loc, eloc, rtol=1e-5, atol=5e-5)
    assert not val[3].any() and not loc[3].any()


def test_rejects_bad_arguments(cuda):
    scene = _scene(0)
    with pytest.raises(ValueError):
        mt.gt_maps_from_depth(scene[0], scene[1], scene[2][:3], scene[3], scene[4], scene[5], "cuda:0")
    with pytest.raises(ValueError):
        mt.gt_maps_from_depth(scene[0], scene[1], scene[2], scene[3], scene[4], scene[5], "cuda:0", centroid_type="top")


def test_engine_accepts_raw_depth_inputs(cuda):
    """a sample with depth_map + instance_masks instead of the three target maps trains on the synthesised targets"""
    from monopsr_b200.core import model_spec as ms
    from monopsr_b200.core.engine import Engine
    S = ms.synthetic_sample(0)
    rng = np.random.RandomState(3)
    H, W = 375, 1242
    depth = rng.uniform(3, 60, (H, W)).astype(np.float32)
    masks = rng.rand(ms.NUM_BOXES, H, W) < 0.6
    raw = {k: v for k, v in S.items() if not k.startswith("gt_inst") and k != "gt_valid_mask_maps"}
    raw["depth_map"], raw["instance_masks"] = depth, masks
    eng = Engine(cuda, params=ms.init_params(0))
    eng.set_inputs(raw)
    loc, glo, val = T.gt_maps(S["boxes_2d"], S["boxes_3d"], masks, depth, S["est_view_angs"], S["cam_p"], roi=48,
                              centroid_type="middle")
    assert np.array_equal(eng.inputs["gt_valid_mask_maps"].cpu().numpy(), val)
    assert np.array_equal(eng.inputs["gt_inst_xyz_maps_global"].cpu().numpy(), glo)
    np.testing.assert_allclose(eng.inputs["gt_inst_xyz_maps_local"].cpu().numpy(), loc, rtol=1e-5, atol=5e-5)
    eng.forward(train=True)
    eng.backward()
    torch.cuda.synchronize()
    assert np.isfinite(eng.losses()["total_loss"])


@pytest.mark.parametrize("as_u8", [True, False])
def test_image_inputs_match_oracle(cuda, as_u8):
    rng = np.random.RandomState(11)
    H, W, n = 375, 1242, 32
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    src = img if as_u8 else img.astype(np.float32) + rng.rand(H, W, 3).astype(np.float32)
    boxes = np.stack([rng.uniform(0, 0.5, n), rng.uniform(0, 0.7, n), rng.uniform(0.55, 1.0, n), rng.uniform(0.75, 1.0, n)],
                     1).astype(np.float32)
    boxes[0] = [0, 0, 1, 1]
    boxes[1] = [-0.05, 0.2, 0.5, 1.1]                      # partly outside the image: extrapolation value 0
    out = mt.image_inputs(src, boxes, "cuda:0")
    torch.cuda.synchronize()
    ecrops, efull, epre = T.image_inputs(src, boxes)
    # fp32 lerps of 0..255 pixel values; the kernel contracts a + (b - a) * l into FMAs, numpy does not: a few ulps of
    # 255 per stage, two stages
    np.testing.assert_allclose(out["img_preprocessed"].cpu().numpy(), epre, rtol=1e-5, atol=5e-4)
    np.testing.assert_allclose(out["rgb_crops"].cpu().numpy(), ecrops, rtol=1e-5, atol=2e-3)
    np.testing.assert_allclose(out["full_img"].cpu().numpy(), efull, rtol=1e-5, atol=2e-3)
    assert out["rgb_crops"].shape == (n, 48, 48, 3) and out["full_img"].shape == (1, 160, 608, 3)
    assert (out["rgb_crops"][1, 0] == 0).all()


def test_image_inputs_reject_bad_arguments(cuda):
    with pytest.raises(ValueError):
        mt.image_inputs(np.zeros((10, 10, 4), np.uint8), np.zeros((1, 4), np.float32), "cuda:0")
    with pytest.raises(ValueError):
        mt.image_inputs(np.zeros((10, 10, 3), np.uint8), np.zeros((1, 4), np.float32), "cuda:0", mean_sub_type="none")
