"""The oracle's box / map geometry (rows a15-a16: oracle/network.py box_heads pieces, prop_cen_y_from_box,
train_projections) against the reference's TENSORFLOW versions of the same functions, unmodified, executed on arrays
through the numpy-backed TF stand-in (tests/golden/make_geometry_tf_golden.py): centroid-z and -y proposals, centroid
x from the viewing angle, local map -> camera frame, projection, expected pixel grid, the normalised projection error
and the global depth map including its linspace-over-rows quirk."""
import importlib.util
import os

import numpy as np
import torch

from oracle import network as onet

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "geometry_tf_golden.npz"))
SUB = (slice(None), slice(1, None, 6), slice(2, None, 6))


def _big_inputs():
    xyz = (np.random.RandomState(22).randn(32, 48, 48, 3) * [1.8, 0.7, 0.9]).astype(np.float32).astype(np.float64)
    valid = (np.random.RandomState(23).rand(32, 48, 48, 1) < 0.6).astype(np.float64)
    valid[5] = 0
    return xyz, valid


def test_generator_and_test_agree_on_the_regenerated_inputs():
    spec = importlib.util.spec_from_file_location("gen", os.path.join(HERE, "golden", "make_geometry_tf_golden.py"))
    src = open(spec.origin).read()
    assert "RandomState(22).randn(n, roi[0], roi[1], 3) * [1.8, 0.7, 0.9]" in src and "RandomState(23).rand(n, roi[0], roi[1], 1) < 0.6" in src


def test_proposals_and_centroid_x():
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    cam_p, b2 = t(G["in/cam_p"]), t(G["in/boxes_2d"])
    lwh = t(G["in/lwh"])
    prop_z = (cam_p[0, 0] * lwh[:, 2] / (b2[:, 2] - b2[:, 0]) + t(G["in/prop_cen_z_offset"])).reshape(-1, 1)    # as box_heads
    np.testing.assert_allclose(prop_z.numpy(), G["out/prop_cen_z"], rtol=1e-13)
    np.testing.assert_allclose(onet.prop_cen_y_from_box(b2, cam_p, prop_z).numpy(), G["out/prop_cen_y"], rtol=1e-12, atol=1e-13)
    x_offset = -cam_p[0, 3] / cam_p[0, 0]
    cen_x = t(G["in/cen_z"]) * torch.tan(t(G["in/est_view_angs"])) + x_offset
    np.testing.assert_allclose(cen_x.numpy(), G["out/cen_x"], rtol=1e-13)


def test_box_heads_computes_those_same_proposals():
    """the formulas above are the ones inside box_heads: run it with zero FC weights (offsets = 0, lwh = mean_lwh)"""
    from monopsr_b200.core import model_spec as ms
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    P = {n: torch.zeros(s, dtype=torch.float64) for n, s, _ in ms.param_table() if n.startswith("output/")}
    S = {"boxes_2d": t(G["in/boxes_2d"]), "cam_p": t(G["in/cam_p"]), "est_view_angs": t(G["in/est_view_angs"]).reshape(-1),
         "class_indices": torch.ones(32, 1, dtype=torch.int64), "mean_lwh": t(G["in/lwh"]),
         "prop_cen_z_offset": t(G["in/prop_cen_z_offset"])}
    out = onet.box_heads(P, S, torch.zeros(32, 6, 6, 512, dtype=torch.float64))
    np.testing.assert_allclose(out["prop_cen_z"].numpy(), G["out/prop_cen_z"], rtol=1e-13)
    np.testing.assert_allclose(out["cen_y"].numpy(), G["out/prop_cen_y"], rtol=1e-12, atol=1e-13)       # offsets are 0
    # centroid x from the PREDICTED z (= the proposal here) and the estimated viewing angle
    cam_p = G["in/cam_p"]
    want = G["out/prop_cen_z"] * np.tan(G["in/est_view_angs"]) - cam_p[0, 3] / cam_p[0, 0]
    np.testing.assert_allclose(out["cen_x"].numpy(), want, rtol=1e-12)


def test_train_projections_against_the_tf_functions():
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    xyz, valid = _big_inputs()
    g = onet.train_projections(t(xyz), t(valid), t(G["in/boxes_2d"]), t(G["in/cam_p"]), t(G["in/est_view_angs"]),
                               t(G["in/gt_view_angs"]), t(G["in/cen_y"]), t(G["in/cen_z"]))
    np.testing.assert_allclose(g["global_xyz"].numpy()[SUB], G["out/xyz_global"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(g["exp_uv"].numpy()[SUB], G["out/exp_proj_uv_map"], rtol=0, atol=1e-9)
    # the reference's debug copy of the projection is masked by the valid map
    np.testing.assert_allclose((g["proj_uv"].numpy() * valid)[SUB], G["out/proj_uv_map"], rtol=1e-11, atol=1e-8)
    np.testing.assert_allclose(g["proj_err_norm"].numpy(), G["out/proj_err_norm"], rtol=1e-10, atol=1e-12)
    assert G["out/proj_err_norm"][5] == 0                                  # no valid pixel: 0 / max(count, 1)
    np.testing.assert_allclose(g["inst_depth_map_global"].numpy()[SUB], G["out/inst_depth_map_global"], rtol=0, atol=1e-10)
