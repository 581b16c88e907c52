"""Entry points (monopsr_b200/experiments/run_*.py, core/experiment.py) end to end on the synthetic KITTI tree with a
stub engine: flags and config handling of the reference scripts, checkpoint discovery, the per-step output folders,
the evaluated-checkpoints list, polling, and the AP results file."""
import os
import shutil
import sys

import numpy as np
import pytest
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import kitti_tree  # noqa: E402
from test_drivers import StubEngine  # noqa: E402
from monopsr_b200.core import experiment as X  # noqa: E402
from monopsr_b200.core import predictions as P  # noqa: E402
from monopsr_b200.experiments import run_evaluation, run_inference, run_training  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def exp(tmp_path):
    """a config file = the reference yaml with the dataset pointed at the synthetic tree and a short schedule"""
    dataset_dir, data_dir = kitti_tree.make_tree(str(tmp_path))
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "monopsr_model_000.yaml")))
    cfg["dataset_config"]["dataset_dir"] = dataset_dir
    cfg["train_config"].update(max_iterations=4, checkpoint_interval=2, summary_interval=2)
    path = str(tmp_path / "monopsr_model_000.yaml")
    yaml.safe_dump(cfg, open(path, "w"))
    return path, data_dir


def test_training_entry_point(exp):
    path, data_dir = exp
    made = []

    def factory(device):
        made.append(StubEngine())
        return made[-1]
    lines = []
    np.random.seed(0)
    run_training.main(["--config_path", path, "--data_dir", data_dir, "--device", "0"], engine_factory=factory,
                      log=lambda *a: lines.append(" ".join(map(str, a))))
    eng = made[0]
    assert [c[1] for c in eng.calls if c[0] == "save"] == [0, 2, 4] and len([c for c in eng.calls if c[0] == "train"]) == 5
    out_dir = os.path.join(data_dir, "outputs", "monopsr_model_000")
    assert os.path.exists(os.path.join(out_dir, "monopsr_model_000.yaml"))               # config copy
    assert sorted(f for f in os.listdir(os.path.join(out_dir, "checkpoints")) if f.endswith(".index")) == \
        ["monopsr-00000000.index", "monopsr-00000002.index", "monopsr-00000004.index"]
    assert any("Step 4: Total Loss" in l for l in lines)
    # a changed config is backed up next to the copy
    cfg = yaml.safe_load(open(path))
    cfg["train_config"]["max_iterations"] = 5
    yaml.safe_dump(cfg, open(path, "w"))
    run_training.main(["--config_path", path, "--data_dir", data_dir], engine_factory=factory, log=lambda *a: None)
    assert len([f for f in os.listdir(out_dir) if f.startswith("monopsr_model_000.yaml.")]) == 1
    assert made[1].loaded == [("monopsr-00000004", "monopsr", False)]                      # resumed from the newest


def _train_first(exp):
    path, data_dir = exp
    run_training.main(["--config_path", path, "--data_dir", data_dir], engine_factory=lambda d: StubEngine(), log=lambda *a: None)
    return path, data_dir


def test_evaluation_entry_point(exp):
    path, data_dir = _train_first(exp)
    lines = []
    res = run_evaluation.main(["--config_path", path, "--data_dir", data_dir, "--data_split", "val", "--max_polls", "1"],
                              engine_factory=lambda d: StubEngine(), log=lambda *a: lines.append(" ".join(map(str, a))))
    # (the dataset object -- and with it the position inside the split -- is shared by the passes, as in the reference)
    assert [r["global_step"] for r in res] == [0, 2, 4] and all(2 <= r["num_samples"] <= 3 for r in res)
    pred = os.path.join(data_dir, "outputs", "monopsr_model_000", "predictions")
    assert open(os.path.join(pred, "evaluated_val.txt")).read().split() == ["0", "2", "4"]
    for step in (0, 2, 4):
        assert os.path.isdir(os.path.join(pred, "predictions_box_3d", "val", str(step)))
        assert os.path.isdir(os.path.join(pred, "kitti_predictions_3d", "val", "0.1", str(step), "data"))
    results = os.path.join(pred, "offline_eval", "results", "val", "monopsr_model_000_results_0.1.txt")
    assert open(results).read().splitlines()[0] == "0"
    assert any(l.startswith("All checkpoints evaluated") for l in lines)
    mdir = os.path.join(pred, "offline_eval", "metrics", "monopsr_model_000", "val")          # metrics csv per step
    rows = open(os.path.join(mdir, "metrics_avg_abs_val.csv")).read().splitlines()
    assert len(rows) == 4 and rows[0].split(",")[0].strip() == "step" and "cen_z_err" in rows[0]
    assert [r.split(",")[0].strip() for r in rows[1:]] == ["0", "2", "4"]
    # a second run finds everything evaluated already
    again = run_evaluation.main(["--config_path", path, "--data_dir", data_dir, "--max_polls", "1"],
                                engine_factory=lambda d: StubEngine(), log=lambda *a: None)
    assert again == []


def test_inference_entry_point(exp):
    path, data_dir = _train_first(exp)
    res = run_inference.main(["--checkpoint_name", "monopsr_model_000", "--data_split", "test", "--ckpt_num", "2",
                              "--data_dir", data_dir], engine_factory=lambda d: StubEngine(), log=lambda *a: None)
    assert [r["global_step"] for r in res] == [2] and res[0]["num_samples"] == 2 and "kitti" not in res[0]
    pred = os.path.join(data_dir, "outputs", "monopsr_model_000", "predictions")
    assert sorted(os.listdir(os.path.join(pred, "predictions_box_3d", "test", "2"))) == ["000008.txt", "000108.txt"]
    assert os.path.isdir(os.path.join(pred, "predictions_" + P.KEY_INST_XYZ_MAP_LOCAL, "test", "2"))
    newest = run_inference.main(["--ckpt_num", "-1", "--data_split", "test", "--data_dir", data_dir],
                                engine_factory=lambda d: StubEngine(), log=lambda *a: None)
    assert [r["global_step"] for r in newest] == [4]
    with pytest.raises(KeyError):
        run_inference.main(["--ckpt_num", "3", "--data_split", "test", "--data_dir", data_dir],
                           engine_factory=lambda d: StubEngine(), log=lambda *a: None)


def test_evaluator_argument_checks(exp, tmp_path):
    path, data_dir = exp
    from monopsr_b200.core import config_utils
    cfg = config_utils.parse_yaml_config(path, data_dir=data_dir)
    with pytest.raises(ValueError):
        X.ExperimentEvaluator(StubEngine(), None, cfg, eval_mode="val")              # no checkpoint directory yet
    os.makedirs(cfg.train_config.paths_config.checkpoint_dir)
    with pytest.raises(ValueError):
        X.ExperimentEvaluator(StubEngine(), None, cfg, eval_mode="train")
    with pytest.raises(ValueError):
        X.ExperimentEvaluator(StubEngine(), None, cfg, eval_mode="test", do_kitti_native_eval=True)
    ev = X.ExperimentEvaluator(StubEngine(), None, cfg, eval_mode="test", do_kitti_native_eval=False)
    assert list(ev.get_evaluated_ckpts()) == [] and X.checkpoints_in(ev.checkpoint_dir, "monopsr") == []
    assert ev.output_types == sorted(ev.output_types) and P.KEY_CENTROIDS in ev.output_types
    with pytest.raises(ValueError):
        ev.run_latest_checkpoints([-1])
