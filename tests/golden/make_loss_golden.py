"""Generate tests/golden/loss_golden.npz: the reference's OWN loss assembly -- MonoPSRModel.loss
(monopsr_model.py:554-958) with loss_builder.add_loss_tensor and the loss classes of object_detection/core/losses.py
and monopsr/core/losses_custom.py, unmodified, imported from /root/reference/src -- EXECUTED on arrays through the
numpy-backed TensorFlow stand-in of tests/golden/fake_tf_numeric.py, with the loss configuration of the reference's
configs/monopsr_model_000.yaml.  Inputs: seeded random output / ground-truth dictionaries (32 boxes; 6x6 maps keep the fixture small -- the loss code is size-agnostic).
Run from the repository root:  python tests/golden/make_loss_golden.py"""
import os
import sys
import types

import numpy as np
import yaml

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_tf_numeric  # noqa: E402


def main():
    fake_tf_numeric.install()
    sys.path.insert(0, "/root/reference/src")
    from monopsr.core.models.monopsr.monopsr_model import MonoPSRModel
    from monopsr.core.models.monopsr.monopsr_output_builder import MonoPSROutputBuilder

    cfg = yaml.safe_load(open("/root/reference/src/monopsr/configs/monopsr_model_000.yaml"))
    mc = cfg["model_config"]

    def obj(d):
        return types.SimpleNamespace(**{k: obj(v) for k, v in d.items()}) if isinstance(d, dict) else d
    model_config = obj(mc)
    n, nb = cfg["dataset_config"]["num_boxes"], cfg["dataset_config"]["num_alpha_bins"]
    rs = np.random.RandomState(7)
    valid = (rs.rand(n, 6, 6, 1) < 0.6).astype(np.float64)
    valid[3] = 0                                             # an instance without a single valid pixel
    out = {"inst_xyz_map_local": rs.randn(n, 6, 6, 3) * 1.5, "lwh_offs": rs.randn(n, 3) * 0.8,
           "alpha_bins": rs.randn(n, nb) * 2, "alpha_regs": rs.randn(n, nb), "cen_z_offs": rs.randn(n, 1) * 2,
           "cen_y_offs": rs.randn(n, 1) * 0.7, "proj_err_norm": rs.randn(n) * 1.2,
           "inst_depth_map_global": rs.randn(n, 6, 6, 1) * 2 + 20}
    gt = {"inst_xyz_map_local": rs.randn(n, 6, 6, 3) * 1.5, "valid_mask_maps": valid, "lwh_offs": rs.randn(n, 3) * 0.8,
          "alpha_bins": rs.randint(0, nb, (n, 1)), "alpha_regs": rs.randn(n, nb), "cen_z_offs": rs.randn(n, 1) * 2,
          "cen_y_offs": rs.randn(n, 1) * 0.7, "inst_depth_map_global": rs.randn(n, 6, 6, 1) * 2 + 20}
    valid_bins = (rs.rand(n, nb) < 0.2).astype(np.float64)
    me = types.SimpleNamespace(
        model_config=model_config, num_boxes=n, output_config=model_config.output_config,
        output_types=MonoPSROutputBuilder.get_output_types_list(model_config.output_config), num_alpha_bins=nb,
        dataset=types.SimpleNamespace(num_alpha_bins=nb), pl_gt_alpha_valid_bins=valid_bins, map_roi_size=mc["map_roi_size"])
    losses, total = MonoPSRModel.loss(me, {k: fake_tf_numeric.t(v) for k, v in out.items()},
                                      {k: fake_tf_numeric.t(v) for k, v in gt.items()})
    # ---- the regression targets (gt_dict of the offset heads): the output builder's own add_*_output methods, with
    # slim.fully_connected answering with given offsets
    from monopsr.core import constants
    preds = {"lwh": rs.randn(n, 3) * 0.3, "cen_y": rs.randn(n, 1) * 0.2, "cen_z": rs.randn(n, 1)}
    import monopsr.core.models.monopsr.monopsr_output_builder as mob
    mob.slim.fully_connected = lambda x, num_outputs, activation_fn=None, scope=None, **k: fake_tf_numeric.t(preds[scope])
    ob = MonoPSROutputBuilder(model_config.output_config, model_config, obj(cfg["dataset_config"]),
                              {constants.FEATURES_FOR_MAP: None, constants.FEATURES_FOR_BOX_3D: None}, n, mc["map_roi_size"],
                              None, "train")
    tg = {"mean_lwh": np.tile([3.892, 1.619, 1.530], (n, 1)), "boxes_3d": rs.rand(n, 7) * [20, 2, 40, 2, 1, 1, 3] + [-10, 1, 5, 3, 1.2, 1.2, -1.5],
          "prop_cen_y": rs.randn(n, 1) * 0.5 + 1.2, "prop_cen_z": rs.rand(n, 1) * 40 + 5}
    b3 = tg["boxes_3d"]
    gt_cen_y = b3[:, 1:2] - b3[:, 5:6] / 2                   # monopsr_model.py:266-270, centroid_type 'middle'
    T = fake_tf_numeric.t
    ob.add_lwh_output(features_to_use=None, est_lwh=T(tg["mean_lwh"]), gt_lwh=T(b3[:, 3:6]))
    ob.add_cen_y_output(output_key=constants.KEY_CEN_Y, features_in=None, prop_cen_y=T(tg["prop_cen_y"]), gt_cen_y=T(gt_cen_y))
    ob.add_cen_z_output(output_key=constants.KEY_CEN_Z, features_in=None, prop_cen_z=T(tg["prop_cen_z"]), gt_cen_z=T(b3[:, 2:3]))
    targets = {"in/" + k: v for k, v in tg.items()}
    targets.update({"pred/" + k: v for k, v in preds.items()})
    for k in ("lwh", "lwh_offs", "cen_y", "cen_y_offs", "cen_z", "cen_z_offs"):
        targets["out/" + k] = np.asarray(ob._output_dict[k], np.float64)
        targets["gt/" + k] = np.asarray(ob._gt_dict[k], np.float64)

    save = {"targets/" + k: v for k, v in targets.items()}
    save.update({"out/" + k: v for k, v in out.items()})
    save.update({"gt/" + k: v for k, v in gt.items()})
    save["gt/alpha_valid_bins"] = valid_bins
    save.update({"loss/" + k: np.asarray(v, np.float64) for k, v in losses.items()})
    save["total"] = np.asarray(total, np.float64)
    # only the per-instance statistics the test needs are stored for the big maps
    path = os.path.join(HERE, "loss_golden.npz")
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.ndim == 4 else v) for k, v in save.items()})
    print({k: float(v) for k, v in losses.items()}, float(total))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
