"""The geometry of the network oracle (oracle/network.py: train_projections, prop_cen_y_from_box) against the
reference's own TF-free numpy twins of the graph ops (tests/golden/make_oracle_geometry_golden.py).  The reference's
tests assert that its TF ops equal these twins (instance_utils_test.py:27-73, transform_utils_test.py:39-100), so this
pins rows a15 / a16 of the oracle -- local map -> camera frame -> image, the expected pixel-centre grid, the centroid-y
proposal -- to the reference's arithmetic.  (The conv / FC / BN part of the oracle has no such twin: unpinned.)"""
import os

import numpy as np
import torch

from oracle import network as onet

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "oracle_geometry_golden.npz"))
SUB = (slice(None), slice(1, None, 5), slice(2, None, 5))


def _run():
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    n = len(G["view"])
    cam_p, view, cen = t(G["cam_p"]), t(G["view"]).reshape(n, 1), t(G["cen"])
    x_offset = -cam_p[0, 3] / cam_p[0, 0]
    # train_projections places the cloud at (cen_z tan(view) + x_offset, cen_y, cen_z): choose cen_z / cen_y freely and
    # hand the numpy twin exactly that centroid
    cen_y, cen_z = cen[:, 1:2], cen[:, 2:3]
    centroid = torch.cat([cen_z * torch.tan(view) + x_offset, cen_y, cen_z], dim=1)
    valid = torch.ones(n, 48, 48, 1, dtype=torch.float64)
    g = onet.train_projections(t(G["xyz_local"]), valid, t(G["boxes_2d"]), cam_p, view * 0.9, view, cen_y, cen_z)
    return g, centroid.numpy()


def test_local_to_global_projection_and_expected_grid():
    g, centroid = _run()
    n = len(G["view"])
    # the golden clouds were produced with G["cen"] as centroid; redo the twin's translation for the centroid used here
    want_glob = G["glob"] - G["cen"][:, None, None, :] + centroid[:, None, None, :]
    np.testing.assert_allclose(g["global_xyz"].numpy()[SUB], want_glob, rtol=0, atol=1e-12)
    # projection of exactly those points with the reference's project_pc_to_image arithmetic (P x / w)
    P = G["cam_p"]
    pts = np.concatenate([want_glob, np.ones(want_glob.shape[:-1] + (1,))], axis=-1) @ P.T
    np.testing.assert_allclose(g["proj_uv"].numpy()[SUB], pts[..., :2] / pts[..., 2:3], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(g["exp_uv"].numpy()[SUB], G["exp_centres"], rtol=0, atol=1e-10)
    assert np.abs(G["exp_centres"] - G["exp_topleft"]).min() > 0.2        # and NOT the top-left convention
    assert g["proj_err_norm"].shape == (n,) and g["inst_depth_map_global"].shape == (n, 48, 48, 1)


def test_projection_twin_consistency():
    """the stored projections are the twin's own: P applied to the twin's global points"""
    P = G["cam_p"]
    pts = np.concatenate([G["glob"], np.ones(G["glob"].shape[:-1] + (1,))], axis=-1) @ P.T
    np.testing.assert_allclose(pts[..., :2] / pts[..., 2:3], G["proj"], rtol=1e-12, atol=1e-9)
    # np_get_tr_mat(angle, t) = R_y(angle) . T(t): translate, THEN rotate about y -- with the sign convention
    # (x' = c x + s z, z' = -s x + c z) the oracle's cos / sin placement uses
    for a, c, m in zip(G["view"], G["cen"], G["tr_mat"]):
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        np.testing.assert_allclose(m[:3, :3], R, atol=1e-15)
        np.testing.assert_allclose(m[:3, 3], R @ c, atol=1e-13)
        assert m[3].tolist() == [0, 0, 0, 1]


def test_centroid_y_proposal():
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    n = len(G["depth"])
    got = onet.prop_cen_y_from_box(t(G["boxes_2d"]), t(G["cam_p"]), t(G["depth"]).reshape(n, 1))
    np.testing.assert_allclose(got.numpy().reshape(-1), G["est_y"], rtol=0, atol=1e-13)


def test_projection_error_definition():
    """proj_err_norm = mean over valid pixels of clip((expected - projected) / box size, +-2), u and v summed
    (monopsr_output_builder.py:681-746), recomputed here from the pinned pieces"""
    g, _ = _run()
    b = G["boxes_2d"]
    exp, proj = g["exp_uv"].numpy(), g["proj_uv"].numpy()
    eu = np.clip((exp[..., 0] - proj[..., 0]) / (b[:, 3] - b[:, 1])[:, None, None], -2, 2)
    ev = np.clip((exp[..., 1] - proj[..., 1]) / (b[:, 2] - b[:, 0])[:, None, None], -2, 2)
    np.testing.assert_allclose(g["proj_err_norm"].numpy(), (eu + ev).sum((1, 2)) / (48 * 48), rtol=1e-12)
