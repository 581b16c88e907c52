"""core/evaluator_utils.py against files written by the reference's own evaluator_utils functions
(tests/golden/make_evaluator_utils_golden.py): the KITTI-format result files of both converters, byte for byte, and the
metrics csv files; then the whole tail of an evaluation run (convert -> native AP evaluation -> results file)."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kitti_tree  # noqa: E402
from monopsr_b200.core import evaluator_utils as U  # noqa: E402
from monopsr_b200.datasets import kitti_loader as KL  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "evaluator_utils_golden.json")))


def tree_files(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in sorted(files):
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = open(p, newline="").read()
    return out


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("evalutils"))
    dataset_dir, data_dir = kitti_tree.make_tree(root)
    cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), dataset_dir,
                                     {"data_split": "trainval"})
    ds = KL.KittiDataset(cfg, "val", data_dir=data_dir)
    return root, ds, kitti_tree.write_predictions(os.path.join(root, "pred"), ds.get_sample_names())


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("projected", {"project_3d_box": True})])
def test_box_3d_converter_writes_the_reference_files(setup, tag, kw):
    root, ds, (d3, d2, _) = setup
    base = os.path.join(root, "out_" + tag)
    U.save_predictions_box_3d_in_kitti_format(0.1, ds, base, d3, d2, 1200, log=None, **kw)
    U.save_predictions_box_3d_in_kitti_format(0.55, ds, base, d3, d2, 1200, log=None, **kw)
    got, want = tree_files(base), GOLD["box_3d/" + tag]
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k] == want[k], k
    assert any(v == "" for v in want.values()) and any("\r\n" in v for v in want.values())


def test_box_2d_converter_writes_the_reference_files(setup):
    root, ds, (_, _, d2only) = setup
    base = os.path.join(root, "out_2d")
    out_dir = U.save_predictions_box_2d_in_kitti_format(0.30000001, ds, base, d2only, 7, log=None)
    assert out_dir.endswith("/kitti_predictions_3d/trainval/0.3/7/data")
    assert tree_files(base) == GOLD["box_2d"]


def test_metrics_csv_equal_the_reference(tmp_path):
    shown = []
    for step in (100, 200):
        U.save_metrics(str(tmp_path), "val", step, GOLD["metrics_in/%d" % step],
                       metrics_to_show=[["metric_cen_z_err", "avg_abs"], ["metric_chamfer", "avg"]],
                       summary_fn=lambda tag, value, gs: shown.append((tag, gs)))
    assert tree_files(str(tmp_path)) == GOLD["metrics"]
    assert shown == [("metrics/avg_abs/metric_cen_z_err", 100), ("metrics/avg/metric_chamfer", 100),
                     ("metrics/avg_abs/metric_cen_z_err", 200), ("metrics/avg/metric_chamfer", 200)]
    with pytest.raises(ValueError):
        U.save_metrics(str(tmp_path), "val", 1, {"m": [1.0]}, metrics_to_show=[["m", "median"]])


def test_time_statistics_lines():
    lines = []
    U.print_inference_time_statistics([0.1, 0.3, 0.2], [0.01, 0.02], log=lines.append)
    assert lines == ["Feed dict time:", "Min:  0.1", "Max:  0.3", "Mean:  0.2", "Median:  0.2",
                     "Inference time:", "Min:  0.01", "Max:  0.02", "Mean:  0.015", "Median:  0.015"]


def test_convert_then_native_ap_evaluation(setup, tmp_path):
    """the tail of Evaluator.run_checkpoint_once in the reference: KITTI-format conversion, AP evaluation against
    label_2, results appended to <results>/<split>/<name>_results_<thr>.txt"""
    root, ds, _ = setup
    # predictions = the ground-truth cars themselves (perfect detector) for every sample of the split
    pred = str(tmp_path / "pred")
    d3, d2 = os.path.join(pred, "box_3d"), os.path.join(pred, "box_2d")
    os.makedirs(d3), os.makedirs(d2)
    from monopsr_b200.datasets import kitti_formats as K
    for name in ds.get_sample_names():
        cars = [o for o in K.read_labels(ds.kitti_label_dir, name) if o.type == "Car"]
        b3 = np.asarray([[o.t[0], o.t[1], o.t[2], o.l, o.w, o.h, o.ry, 0.9 - 0.05 * i, 0] for i, o in enumerate(cars)]).reshape(-1, 9)
        b2 = np.asarray([[o.y1, o.x1, o.y2, o.x2, o.alpha, 0.9 - 0.05 * i, 0] for i, o in enumerate(cars)]).reshape(-1, 7)
        np.savetxt(os.path.join(d3, name + ".txt"), b3, fmt="%0.5f")
        np.savetxt(os.path.join(d2, name + ".txt"), b2, fmt="%0.5f")
    base = str(tmp_path / "predictions")
    U.save_predictions_box_3d_in_kitti_format(0.1, ds, base, d3, d2, 42, log=None)
    lines = []
    res = U.run_kitti_native_eval("ckpt", ds.data_split, 0.1, 42, ds.kitti_label_dir, base, str(tmp_path / "offline_eval"),
                                  log=lines.append)
    for k in ("car_detection", "car_orientation", "car_detection_BEV", "car_detection_3D", "car_heading_3D"):
        c = res["curves"][k]            # a perfect detector: precision / similarity 1 at every recall point reached
        assert c[1][0] == 1.0 and c[2][0] == 1.0 and set(np.unique(np.round(c, 12))) <= {0.0, 1.0}, (k, c)
    assert res["ap"]["car_detection"] == res["ap"]["car_detection_3D"] == res["ap"]["car_heading_BEV"]
    txt = open(str(tmp_path / "offline_eval" / "results" / "trainval" / "ckpt_results_0.1.txt")).read().splitlines()
    assert txt[0] == "42" and txt[1:] == res["lines"] and lines == ["42"] + res["lines"]
    U.run_kitti_native_eval("ckpt", ds.data_split, 0.1, 42, ds.kitti_label_dir, base, str(tmp_path / "offline_eval"),
                            low_iou=True, log=None)
    assert os.path.exists(str(tmp_path / "offline_eval" / "results_low_iou" / "trainval" / "ckpt_results_0.1.txt"))
