"""Generate tests/golden/metrics_golden.npz: the reference's OWN MonoPSRModel.evaluate_predictions
(monopsr_model.py:1105-1221), unmodified, executed on arrays through the numpy-backed TF stand-in.  The two custom ops
it calls are answered by plain numpy (brute-force nearest neighbours; a fixed matching with its cost), because what
is pinned here is the assembly around them: masking, per-object slicing by num_objs, division by the number of valid
pixels, and the centroid / dimension / viewing-angle error definitions.
Run from the repository root:  python tests/golden/make_metrics_golden.py"""
import os
import sys
import types

import numpy as np
import yaml

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_tf_numeric as F  # noqa: E402


def nn_distance(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
    return F.t(d.min(2)), F.t(d.argmin(2)), F.t(d.min(1)), F.t(d.argmin(1))


def approx_match(a, b):
    n = np.asarray(a).shape[1]
    return F.t(np.tile(np.eye(n)[None], (np.asarray(a).shape[0], 1, 1)))          # point i <-> point i


def match_cost(a, b, match):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.sqrt(((b[:, :, None, :] - a[:, None, :, :]) ** 2).sum(-1))             # [l, k] = |xyz2_l - xyz1_k|
    return F.t((d * np.asarray(match)).sum((1, 2)))


def main():
    F.install()
    sys.path.insert(0, "/root/reference/src")
    import monopsr.core.models.monopsr.monopsr_model as mm
    from monopsr.core.models.monopsr.monopsr_output_builder import MonoPSROutputBuilder
    mm.tf_nndistance.nn_distance = nn_distance
    mm.tf_approxmatch.approx_match, mm.tf_approxmatch.match_cost = approx_match, match_cost
    cfg = yaml.safe_load(open("/root/reference/src/monopsr/configs/monopsr_model_000.yaml"))

    def obj(d):
        return types.SimpleNamespace(**{k: obj(v) for k, v in d.items()}) if isinstance(d, dict) else d
    mc = obj(cfg["model_config"])
    n, num_objs, roi = 32, 5, 6
    rs = np.random.RandomState(31)
    valid = (rs.rand(n, roi, roi, 1) < 0.6).astype(np.float64)
    pred = {"inst_xyz_map_local": rs.randn(n, roi, roi, 3), "prop_cen_z": rs.rand(n, 1) * 40, "centroids": rs.randn(n, 3) * 5,
            "lwh_offs": rs.randn(n, 3) * 0.3, "view_ang": rs.randn(n, 1) * 0.3}
    gt = {"inst_xyz_map_local": rs.randn(n, roi, roi, 3), "valid_mask_maps": valid, "centroids": rs.randn(n, 3) * 5,
          "lwh_offs": rs.randn(n, 3) * 0.3, "view_ang": rs.randn(n, 1) * 0.3}
    me = types.SimpleNamespace(output_types=MonoPSROutputBuilder.get_output_types_list(mc.output_config), num_boxes=n,
                               pl_num_objs=num_objs)
    metrics, _ = mm.MonoPSRModel.evaluate_predictions(me, {k: F.t(v) for k, v in pred.items()}, {k: F.t(v) for k, v in gt.items()})
    save = {"pred/" + k: v for k, v in pred.items()}
    save.update({"gt/" + k: v for k, v in gt.items()})
    save.update({"metric/" + k: np.asarray(v, np.float64) for k, v in metrics.items()})
    save["num_objs"] = np.asarray(num_objs)
    # the two point-set LOSS classes (losses_custom.py:135-198) on the same arrays, same stand-in ops
    import monopsr.core.losses_custom as lc
    lc.tf_nndistance.nn_distance = nn_distance
    lc.tf_approxmatch.approx_match, lc.tf_approxmatch.match_cost = approx_match, match_cost
    save["loss/chamfer_dist"] = np.asarray(lc.ChamferDistance()(F.t(pred["inst_xyz_map_local"]), F.t(gt["inst_xyz_map_local"]),
                                                                weights=F.t(valid)), np.float64)
    save["loss/emd"] = np.asarray(lc.EarthMoversDistance()(F.t(pred["inst_xyz_map_local"]), F.t(gt["inst_xyz_map_local"]),
                                                           weights=F.t(valid)), np.float64)
    path = os.path.join(HERE, "metrics_golden.npz")
    np.savez_compressed(path, **save)
    print({k: np.asarray(v).shape for k, v in metrics.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
