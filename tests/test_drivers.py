"""Training / evaluation drivers (core/trainer.py, core/evaluator.py) with a stub engine: control flow, checkpoint
naming / resume / pruning and the prediction files -- the reference's trainer.py:122-209 and evaluator loop."""
import os

import numpy as np
import pytest

from monopsr_b200.core import evaluator as E
from monopsr_b200.core import predictions as P
from monopsr_b200.core import trainer as T
from monopsr_b200.core.config_utils import config_dict_to_object


class StubEngine(object):
    """records calls; save_checkpoint writes an empty '<prefix>.index' like a tensor bundle would"""

    def __init__(self):
        self.calls, self.step_count, self.loaded = [], 0, []
        self.N = 32

    def train_step(self, sample):
        self.calls.append(("train", self.step_count))
        self.step_count += 1

    def losses(self):
        return {"total_loss": 10.0 / (1 + self.step_count), "lwh_offs": 1.0}

    def save_checkpoint(self, prefix, global_step=None):
        for ext in (".index", ".data-00000-of-00001"):
            open(prefix + ext, "wb").close()
        self.calls.append(("save", global_step))

    def load_checkpoint(self, prefix, kind="monopsr", use_ema=False, resume=False):
        self.loaded.append((os.path.basename(prefix), kind, use_ema))
        return {"loaded": [], "missing": []}

    # inference side
    def set_inputs(self, sample):
        self.sample = sample

    def forward(self, train=True, compute_losses=None):
        self.calls.append(("forward", train))

    def outputs(self):
        rng = np.random.RandomState(0)
        n = self.N
        b3 = self.sample.get("boxes_3d")          # absent in 'test' mode samples
        if b3 is None:
            b3 = np.tile(np.asarray([0.0, 1.5, 20.0, 3.9, 1.6, 1.5, 0.0], np.float32), (n, 1))
        return {"inst_xyz_map_local": rng.randn(n, 48, 48, 3).astype(np.float32), "valid_mask_maps": None,
                "lwh": b3[:, 3:6], "alpha_bins": rng.randn(n, 12).astype(np.float32),
                "alpha_regs": rng.randn(n, 12).astype(np.float32) * 0.1, "view_ang": self.sample["est_view_angs"][:, None],
                "centroids": b3[:, :3].copy(), "prop_cen_z": b3[:, 2:3] + 0.5, "lwh_offs": np.zeros((n, 3), np.float32)}


def _config(tmp_path, overwrite=False, max_it=25):
    return config_dict_to_object({
        "config_name": "stub_cfg", "model_config": {"model_type": "monopsr_model"},
        "train_config": {"max_iterations": max_it, "summary_interval": 10, "checkpoint_interval": 10,
                         "max_checkpoints_to_keep": 2, "overwrite_checkpoints": overwrite,
                         "paths_config": {"checkpoint_dir": str(tmp_path / "ckpt"), "logdir": str(tmp_path / "log")}}})


def test_train_loop_checkpoints_and_resume(tmp_path):
    eng, lines = StubEngine(), []
    cfg = _config(tmp_path)
    loss = T.train(eng, cfg, lambda: {"x": 1}, log=lambda *a: lines.append(" ".join(map(str, a))))
    trains = [c for c in eng.calls if c[0] == "train"]
    assert len(trains) == 26 and [c[1] for c in eng.calls if c[0] == "save"] == [0, 10, 20]      # steps 0..25 inclusive
    assert eng.calls.index(("save", 10)) < eng.calls.index(("train", 10))                        # saved BEFORE the step
    ck = sorted(f for f in os.listdir(tmp_path / "ckpt") if f.endswith(".index"))
    assert ck == ["monopsr_model-00000010.index", "monopsr_model-00000020.index"]                # max_to_keep = 2
    assert any("Step 20: Total Loss" in l for l in lines) and any("Starting from step 0 / 25" in l for l in lines)
    assert loss == pytest.approx(10.0 / 22)
    # resume: newest checkpoint, step from its name
    eng2 = StubEngine()
    T.train(eng2, _config(tmp_path, max_it=30), lambda: {}, log=lambda *a: None)
    assert eng2.loaded == [("monopsr_model-00000020", "monopsr", False)]
    assert [c[1] for c in eng2.calls if c[0] == "train"] == list(range(20, 31))
    # overwrite_checkpoints: start from scratch, pre-trained detection weights if given
    eng3 = StubEngine()
    T.train(eng3, _config(tmp_path, overwrite=True, max_it=3), lambda: {}, pretrained_checkpoint="/x/model.ckpt",
            log=lambda *a: None)
    assert eng3.loaded == [("model.ckpt", "detection", False)] and eng3.calls[0] == ("save", 0)
    assert T.latest_checkpoint(str(tmp_path / "nope"), "m") == (None, 0)


def test_evaluator_writes_predictions(tmp_path):
    from monopsr_b200.core import model_spec as ms
    S = ms.synthetic_sample(0)
    eng = StubEngine()
    types = [P.KEY_INST_XYZ_MAP_LOCAL, P.KEY_VALID_MASK_MAPS, P.KEY_CENTROIDS, P.KEY_LWH, P.KEY_VIEW_ANG, P.KEY_ALPHA]
    dirs = {P.OUT_DIR_XYZ_MAP_LOCAL: str(tmp_path / "xyz"), P.OUT_DIR_BOX_2D: str(tmp_path / "b2"),
            P.OUT_DIR_BOX_3D: str(tmp_path / "b3")}
    ev = E.Evaluator(eng, types, dirs, train_val_test="test", log=lambda *a: None)
    sd = {P.SAMPLE_NAME: "000007", P.SAMPLE_IMAGE_INPUT: np.zeros((375, 1242, 3), np.uint8), P.SAMPLE_NUM_OBJS: 4,
          P.SAMPLE_CAM_P: S["cam_p"], P.SAMPLE_LABEL_SCORES: np.linspace(0.3, 0.9, 32).astype(np.float32),
          P.SAMPLE_LABEL_BOXES_2D: S["boxes_2d"]}
    res = ev.run_checkpoint_once("/ckpts/monopsr_model-00120000", [(S, sd)])
    # the RAW variables, as the reference's evaluator restores them (plain tf.train.Saver: core/evaluator.py:125,144)
    assert res["num_samples"] == 1 and eng.loaded == [("monopsr_model-00120000", "monopsr", False)]
    assert ("forward", False) in eng.calls
    b3 = np.loadtxt(os.path.join(dirs[P.OUT_DIR_BOX_3D], "000007.txt"))
    b2 = np.loadtxt(os.path.join(dirs[P.OUT_DIR_BOX_2D], "000007.txt"))
    assert b3.shape == (4, 9) and b2.shape == (4, 7)
    assert np.load(os.path.join(dirs[P.OUT_DIR_XYZ_MAP_LOCAL], "000007.npy")).shape == (4, 48, 48, 3)
    with pytest.raises(ValueError):
        E.Evaluator(eng, types, dirs, train_val_test="train")


def test_loader_to_ap_numbers_with_a_stub_engine(tmp_path):
    """val split of the synthetic KITTI tree -> PrefetchLoader(epochs=1) -> Evaluator (the stub engine answers with the
    ground-truth boxes) -> prediction files -> KITTI result files -> AP lines and the results file"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import kitti_tree
    from monopsr_b200.datasets import kitti_loader as KL
    dataset_dir, data_dir = kitti_tree.make_tree(str(tmp_path / "tree"))
    cfg = kitti_tree.apply_overrides(KL.DatasetBuilder.get_config_obj(KL.DatasetBuilder.KITTI_TRAIN), dataset_dir,
                                     {"data_split": "val", "use_mscnn_detections": False})
    ds = KL.KittiDataset(cfg, "val", data_dir=data_dir, rng=np.random.RandomState(0))
    eng = StubEngine()
    types = [P.KEY_CENTROIDS, P.KEY_LWH, P.KEY_VIEW_ANG, P.KEY_ALPHA]
    base = str(tmp_path / "predictions")
    dirs = {P.OUT_DIR_BOX_2D: base + "/box_2d", P.OUT_DIR_BOX_3D: base + "/box_3d"}
    lines = []
    ev = E.Evaluator(eng, types, dirs, train_val_test="val", log=lambda *a: lines.append(" ".join(map(str, a))))
    with KL.PrefetchLoader(ds, epochs=1) as loader:
        res = ev.run_checkpoint_once(None, loader)
    assert res["num_samples"] == 3 and set(res["mean_losses"]) == {"total_loss", "lwh_offs"}
    assert sorted(os.listdir(dirs[P.OUT_DIR_BOX_3D])) == ["000008.txt", "000108.txt"]
    done = str(tmp_path / "already_evaluated.txt")
    # (box scores of unscored ground-truth labels are tiny: score_boxes multiplies by the label score)
    out = ev.convert_and_evaluate(ds, base, 1200, kitti_score_threshold=0.001, checkpoint_name="stub_cfg",
                                  already_evaluated_path=done)
    assert len(out) == 1 and open(done).read().split() == ["1200"]
    ap = out[0]["ap"]
    assert {"car_detection", "car_detection_BEV", "car_detection_3D", "car_heading_3D"} <= set(ap)
    c = out[0]["curves"]["car_detection"]
    assert c[2][0] == 1.0                                   # the labelled 2-D boxes come back: precision 1
    assert out[0]["curves"]["car_detection_3D"].shape == (3, 41)      # (the stub's 3-D boxes are not meant to match)
    results = open(os.path.join(base, "offline_eval", "results", "val", "stub_cfg_results_0.001.txt")).read().splitlines()
    assert results[0] == "1200" and results[1:] == out[0]["lines"]
    assert any(l.startswith("car_detection_3D AP:") for l in lines) and "Finished evaluation" in lines[-1]
    kitti_files = sorted(os.listdir(os.path.join(base, "kitti_predictions_3d", "val", "0.001", "1200", "data")))
    assert kitti_files == ["000008.txt", "000076.txt", "000108.txt"]       # a sample without detections: empty file


def test_non_chief_rank_trains_without_writing(tmp_path):
    eng, lines = StubEngine(), []
    T.train(eng, _config(tmp_path, max_it=12), lambda: {}, log=lines.append, chief=False)
    assert len([c for c in eng.calls if c[0] == "train"]) == 13 and not [c for c in eng.calls if c[0] == "save"]
    assert lines == [] and not [f for f in os.listdir(tmp_path / "ckpt")]
    from monopsr_b200.core import experiment as X
    assert X.data_parallel_setup("cuda:3") == (0, 1, "cuda:3")


def test_validation_metrics():
    """core/metrics.py: the error metrics of MonoPSRModel.evaluate_predictions on the first num_objs rows"""
    from monopsr_b200.core import metrics as M
    rs = np.random.RandomState(0)
    n = 8
    b3 = rs.rand(n, 7).astype(np.float32) * [10, 2, 40, 4, 2, 2, 3]
    out = {"prop_cen_z": rs.rand(n, 1).astype(np.float32) * 40, "centroids": rs.rand(n, 3).astype(np.float32) * 10,
           "lwh": rs.rand(n, 3).astype(np.float32) * 4, "lwh_offs": rs.rand(n, 3).astype(np.float32),
           "view_ang": rs.rand(n, 1).astype(np.float32)}
    sample = {"boxes_3d": b3, "gt_view_angs": rs.rand(n).astype(np.float32)}
    types = [P.KEY_CENTROIDS, P.KEY_LWH, P.KEY_VIEW_ANG, P.KEY_INST_XYZ_MAP_LOCAL]
    m = M.evaluate_predictions(out, sample, 3, types, "middle",
                               point_set={"metric_emd": np.arange(8.0), "metric_chamfer": np.arange(8.0) * 2})
    cen = np.column_stack([b3[:, 0], b3[:, 1] - b3[:, 5] / 2, b3[:, 2]])
    assert m[M.METRIC_EMD].tolist() == [0, 1, 2] and m[M.METRIC_CHAMFER].tolist() == [0, 2, 4]
    np.testing.assert_allclose(m[M.METRIC_PROP_CEN_Z_ERR], cen[:3, 2:3] - out["prop_cen_z"][:3], rtol=1e-6)
    np.testing.assert_allclose(m[M.METRIC_CEN_Y_ERR], cen[:3, 1] - out["centroids"][:3, 1], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m[M.METRIC_DIM_ERR], (b3[:3, 3:6] - out["lwh"][:3]) - out["lwh_offs"][:3], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m[M.METRIC_VIEW_ANG_ERR], sample["gt_view_angs"][:3, None] - out["view_ang"][:3], rtol=1e-6)
    assert m[M.METRIC_DIM_ERR].shape == (3, 3) and m[M.METRIC_CEN_X_ERR].shape == (3,)
    bottom = M.evaluate_predictions(out, sample, 3, [P.KEY_CENTROIDS], "bottom")
    np.testing.assert_allclose(bottom[M.METRIC_CEN_Y_ERR], b3[:3, 1] - out["centroids"][:3, 1], rtol=1e-5, atol=1e-6)
    assert set(bottom) == {M.METRIC_PROP_CEN_Z_ERR, M.METRIC_CEN_X_ERR, M.METRIC_CEN_Y_ERR, M.METRIC_CEN_Z_ERR}
    with pytest.raises(ValueError):
        M.gt_centroids(b3, "top")
    lists = M.accumulate({}, m)
    lists = M.accumulate(lists, {M.METRIC_EMD: np.array([np.nan, 1.0]), M.METRIC_CHAMFER: np.array([5.0])})
    assert lists[M.METRIC_EMD] == [0, 1, 2] and lists[M.METRIC_CHAMFER] == [0, 2, 4, 5] and len(lists[M.METRIC_DIM_ERR]) == 9


def test_metrics_equal_the_reference_evaluate_predictions():
    """core/metrics.py + losses_custom.point_set_metrics against MonoPSRModel.evaluate_predictions executed on arrays
    (tests/golden/make_metrics_golden.py); the two custom ops are answered by the same plain stand-ins on both sides"""
    import torch
    from monopsr_b200.core import losses_custom
    from monopsr_b200.core import metrics as M
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.npz"))
    pred = {k[5:]: G[k] for k in G.files if k.startswith("pred/")}
    gt = {k[3:]: G[k] for k in G.files if k.startswith("gt/")}
    n_obj = int(G["num_objs"])

    def nn_distance(a, b):
        d = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
        return d.min(2).values, d.argmin(2), d.min(1).values, d.argmin(1)

    def approx_match(a, b):
        return torch.eye(a.shape[1], dtype=a.dtype).expand(a.shape[0], -1, -1)

    def match_cost(a, b, match):
        return (torch.sqrt(((b[:, :, None, :] - a[:, None, :, :]) ** 2).sum(-1)) * match).sum((1, 2))
    t = lambda a: torch.as_tensor(np.asarray(a, np.float64))
    ps = losses_custom.point_set_metrics(t(pred["inst_xyz_map_local"]), t(gt["inst_xyz_map_local"]), t(gt["valid_mask_maps"]),
                                         n_obj, ops=(approx_match, match_cost, nn_distance))
    # a sample whose derived ground truth is the golden gt_dict: pred lwh = 0 so that gt offsets = box dimensions
    b3 = np.zeros((32, 7), np.float32)
    b3[:, 3:6] = gt["lwh_offs"]
    b3[:, 0], b3[:, 2] = gt["centroids"][:, 0], gt["centroids"][:, 2]
    b3[:, 1] = gt["centroids"][:, 1] + b3[:, 5] / 2
    outputs = {"prop_cen_z": pred["prop_cen_z"], "centroids": pred["centroids"], "lwh": np.zeros((32, 3)),
               "lwh_offs": pred["lwh_offs"], "view_ang": pred["view_ang"]}
    sample = {"boxes_3d": b3, "gt_view_angs": gt["view_ang"][:, 0]}
    types_ = [P.KEY_INST_XYZ_MAP_LOCAL, P.KEY_CENTROIDS, P.KEY_LWH, P.KEY_VIEW_ANG]
    m = M.evaluate_predictions(outputs, sample, n_obj, types_, "middle", point_set={k: v.numpy() for k, v in ps.items()})
    want = {k[7:]: G[k] for k in G.files if k.startswith("metric/")}
    assert sorted(m) == sorted(want)
    for k, v in want.items():
        assert np.asarray(m[k]).shape == v.shape, (k, np.asarray(m[k]).shape, v.shape)
        np.testing.assert_allclose(m[k], v, rtol=2e-6, atol=2e-6, err_msg=k)          # (boxes_3d round-trips through float32)
    # the two point-set loss classes: mask both clouds, flatten to (B, h*w, 3), op, sum / B
    cd = losses_custom.ChamferDistance(ops=(nn_distance,))(t(pred["inst_xyz_map_local"]), t(gt["inst_xyz_map_local"]),
                                                            weights=t(gt["valid_mask_maps"]))
    emd = losses_custom.EarthMoversDistance(ops=(approx_match, match_cost))(
        t(pred["inst_xyz_map_local"]), t(gt["inst_xyz_map_local"]), weights=t(gt["valid_mask_maps"]))
    np.testing.assert_allclose(float(cd), float(G["loss/chamfer_dist"]), rtol=1e-12)
    np.testing.assert_allclose(float(emd), float(G["loss/emd"]), rtol=1e-12)
