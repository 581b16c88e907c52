"""THE WIRING TEST: oracle.network.forward / loss (with oracle.targets for the inputs) against the reference's WHOLE
training graph executed end to end -- MonoPSRModel.__init__ / build / loss, net_builder, the ResNet builders and the
output builder, all unmodified -- on arrays (tests/golden/make_graph_golden.py, tests/golden/fake_tf_full.py: numeric
TF-slim layers with real scoping; the TF kernels themselves are supplied by the oracle's primitives, so what is
compared is what feeds what, in which order, with which arguments -- from the raw camera image, depth map and instance
masks to every output tensor and every loss term).  A wiring difference shows up as O(1); the tolerance covers the
float32 arithmetic of oracle.targets (as TF computes those inputs) against the float64 stand-in."""
import importlib.util
import os

import numpy as np
import torch

from monopsr_b200.core import model_spec as ms
from oracle import network as onet
from oracle import targets as T

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "graph_golden.npz"))


def _raw_sample():
    spec = importlib.util.spec_from_file_location("make_graph_golden", os.path.join(HERE, "golden", "make_graph_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen.raw_sample()


import pytest


@pytest.mark.parametrize("mode", ["train", "val"])
def test_whole_graph(mode):
    """'val': the is_training=False graph (decoder batch norm on its moving statistics) with projections and losses"""
    prefix = "" if mode == "train" else "val/"
    R = _raw_sample()
    f32 = lambda a: np.asarray(a, np.float32)
    crops, full, _ = T.image_inputs(f32(R["rgb_image"]), f32(R["boxes_2d_norm"]))
    loc, glo, val = T.gt_maps(f32(R["boxes_2d"]), f32(R["boxes_3d"]), f32(R["instance_masks"]), f32(R["depth_map"]),
                              f32(R["est_view_angs"]), f32(R["cam_p"]), roi=48, centroid_type="middle", rotate_view=True)
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    S = {"rgb_crops": t(crops), "full_img": t(full), "boxes_2d": t(R["boxes_2d"]), "boxes_2d_norm": t(R["boxes_2d_norm"]),
         "cam_p": t(R["cam_p"]), "class_indices": torch.as_tensor(R["class_indices"]), "mean_lwh": t(R["mean_lwh"]),
         "prop_cen_z_offset": t(R["prop_cen_z_offset"]), "est_view_angs": t(R["est_view_angs"]), "boxes_3d": t(R["boxes_3d"]),
         "gt_alpha_bins": torch.as_tensor(R["alpha_bins"]), "gt_alpha_regs": t(R["alpha_regs"]),
         "gt_alpha_valid_bins": t(R["alpha_valid_bins"]), "gt_view_angs": t(R["view_angs"]),
         "gt_inst_xyz_maps_local": t(loc), "gt_inst_xyz_maps_global": t(glo), "gt_valid_mask_maps": t(val)}
    P = onet.to_torch(ms.init_params(0, randomize_bn=True), torch.float64)
    with torch.no_grad():
        out, _ = onet.forward(P, S, train=(mode == "train"), projections=True)
        L, total = onet.loss(out, S)
    want = {k[len(prefix) + 4:]: G[k] for k in G.files if k.startswith(prefix + "out/")}
    assert len(want) == 16
    assert set(want) <= set(out) | {"valid_mask_maps"}
    for k, v in want.items():
        a = out[k].numpy()
        a = a[:, ::6, ::6] if a.ndim == 4 else a.reshape(v.shape)
        if k == "valid_mask_maps":
            assert np.array_equal(a, v)
            continue
        err = np.linalg.norm(a - v) / max(np.linalg.norm(v), 1e-30)
        assert err < 2e-4, (k, err)
    terms = [k for k in G.files if k.startswith(prefix + "loss/")]
    assert len(terms) == 8
    for k in terms:
        name = k[len(prefix) + 5:]
        assert abs(float(L[name]) - float(G[k])) <= 2e-4 * max(1.0, abs(float(G[k]))), (k, float(L[name]), float(G[k]))
    assert abs(float(total) - float(G[prefix + "total"])) <= 2e-4 * float(G[prefix + "total"])
    if mode == "val":
        assert abs(float(G["val/total"]) - float(G["total"])) > 1e-3        # the two graphs really differ
    # every variable the reference's graph created is in the parameter table -- except block4, which it builds and
    # never uses
    created = set(G["created"].tolist())
    table = {n for n, _, _ in ms.param_table()}
    assert table <= created and all("/block4/" in n for n in created - table)
