"""The oracle's loss (oracle/network.py::loss, row a17) against the reference's OWN loss assembly executed on arrays:
MonoPSRModel.loss + loss_builder + the loss classes, unmodified, run through a numpy-backed stand-in for the TF ops
they call (tests/golden/fake_tf_numeric.py, tests/golden/make_loss_golden.py).  Which tensor meets which target under
which mask, the label smoothing, the loss weights of the yaml, the divisions by num_boxes and the total are the
reference's; the Huber / softmax-cross-entropy / SUM_BY_NONZERO_WEIGHTS primitives are restated from the TF docs."""
import os

import numpy as np
import torch

from oracle import network as onet

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "loss_golden.npz"))


def _inputs():
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    n = 32
    out = {k[4:]: t(G[k]) for k in G.files if k.startswith("out/")}
    gt = {k[3:]: G[k] for k in G.files if k.startswith("gt/")}
    # the oracle derives the regression targets from the sample: build a sample whose derived targets are the golden ones
    h = np.full((n, 1), 1.5)
    b3 = np.zeros((n, 7))
    b3[:, 3:6] = gt["lwh_offs"]                      # gt_lwh - pred_lwh with pred_lwh = 0
    b3[:, 5:6] = gt["lwh_offs"][:, 2:3]
    b3[:, 2:3] = gt["cen_z_offs"]                    # gt_cen_z - prop_cen_z with prop_cen_z = 0
    b3[:, 1:2] = gt["cen_y_offs"] + b3[:, 5:6] / 2   # 'middle' centroid: y - h/2 - prop_cen_y with prop_cen_y = 0
    out["lwh"] = torch.zeros(n, 3, dtype=torch.float64)
    out["prop_cen_z"] = torch.zeros(n, 1, dtype=torch.float64)
    out["cen_y"] = out["cen_y_offs"].clone()         # prop_cen_y = cen_y - cen_y_offs = 0
    S = {"gt_valid_mask_maps": t(gt["valid_mask_maps"]), "gt_inst_xyz_maps_local": t(gt["inst_xyz_map_local"]),
         "boxes_3d": t(b3), "gt_alpha_bins": torch.as_tensor(gt["alpha_bins"].reshape(-1)),
         "gt_alpha_regs": t(gt["alpha_regs"]), "gt_alpha_valid_bins": t(gt["alpha_valid_bins"]),
         "gt_inst_xyz_maps_global": torch.cat([torch.zeros(n, 6, 6, 2, dtype=torch.float64), t(gt["inst_depth_map_global"])], dim=-1)}
    return out, S


def test_every_loss_term_and_the_total():
    out, S = _inputs()
    L, total = onet.loss(out, S)
    want = {k[5:]: float(G[k]) for k in G.files if k.startswith("loss/")}
    assert sorted(L) == sorted(want)
    for k, v in want.items():
        assert abs(float(L[k]) - v) <= 1e-6 * max(1.0, abs(v)), (k, float(L[k]), v)       # maps were stored as float32
    assert abs(float(total) - float(G["total"])) <= 1e-6 * float(G["total"])
    assert want["inst_xyz_map_local"] > 1 and want["alpha_bins"] > 0.5                    # non-trivial values


def test_terms_react_to_their_own_inputs_only():
    out, S = _inputs()
    base, _ = onet.loss(out, S)
    for key, term in (("lwh_offs", "lwh_offs"), ("alpha_regs", "alpha_regs"), ("cen_z_offs", "cen_z_offs"),
                      ("proj_err_norm", "proj_err"), ("inst_depth_map_global", "inst_depth_map_global")):
        o2 = dict(out)
        o2[key] = out[key] + 0.37
        L2, _ = onet.loss(o2, S)
        changed = {k for k in base if abs(float(L2[k]) - float(base[k])) > 1e-12}
        assert changed == {term}, (key, changed)


def test_regression_targets():
    """gt_dict of the offset heads as MonoPSROutputBuilder.add_lwh_output / add_cen_y_output / add_cen_z_output create it
    (executed on arrays, fully_connected answering with given offsets) against oracle.regression_targets -- including
    the quirk that the 'ground-truth' dimension offsets are measured from the PREDICTED dimensions"""
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    g = {k[len("targets/"):]: G[k] for k in G.files if k.startswith("targets/")}
    np.testing.assert_allclose(g["out/lwh"], g["in/mean_lwh"] + g["pred/lwh"])             # sanity of the recording
    out = {"lwh": t(g["out/lwh"]), "lwh_offs": t(g["out/lwh_offs"]), "cen_y": t(g["out/cen_y"]),
           "cen_y_offs": t(g["out/cen_y_offs"]), "prop_cen_z": t(g["in/prop_cen_z"])}
    T = onet.regression_targets(out, {"boxes_3d": t(g["in/boxes_3d"])})
    for k in ("lwh_offs", "cen_y_offs", "cen_z_offs"):
        np.testing.assert_allclose(T[k].numpy(), g["gt/" + k], rtol=1e-12, atol=1e-13, err_msg=k)
    assert np.allclose(g["gt/lwh_offs"], g["in/boxes_3d"][:, 3:6] - g["out/lwh"])          # gt_lwh - PREDICTED lwh
