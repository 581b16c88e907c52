"""Generate tests/golden/evaluator_utils_golden.json with the reference's OWN monopsr.core.evaluator_utils functions
(imported from /root/reference/src; tensorflow / pypng / distutils are stubbed by an import hook -- the functions used
do not touch them) on the synthetic KITTI tree of tests/kitti_tree.py and seeded synthetic prediction files
(tests/kitti_tree.write_predictions): the KITTI-format result files of both converters and the metrics csv files.
Run from the repository root:  python tests/golden/make_evaluator_utils_golden.py"""
import importlib.abc
import importlib.machinery
import json
import os
import sys
import tempfile
import types

import numpy as np
import yaml

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import kitti_tree  # noqa: E402


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return self


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in ("tensorflow", "png", "distutils"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def tree_files(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in sorted(files):
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = open(p, newline="").read()
    return out


def main():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, "/root/reference/src")
    _load = yaml.load
    yaml.load = lambda s, Loader=yaml.SafeLoader: _load(s, Loader=Loader)
    import monopsr
    from monopsr.builders.dataset_builder import DatasetBuilder
    from monopsr.core import evaluator_utils

    out = {}
    with tempfile.TemporaryDirectory() as root:
        dataset_dir, data_dir = kitti_tree.make_tree(root)
        monopsr.data_dir = lambda: data_dir
        monopsr.scripts_dir = lambda: os.path.join(root, "scripts")
        cfg = kitti_tree.apply_overrides(DatasetBuilder.get_config_obj(DatasetBuilder.KITTI_TRAIN), dataset_dir,
                                         {"data_split": "trainval"})
        ds = DatasetBuilder.build_kitti_dataset(cfg, "val")
        d3, d2, d2only = kitti_tree.write_predictions(os.path.join(root, "pred"), ds.get_sample_names())
        for tag, kw in (("plain", {}), ("projected", {"project_3d_box": True})):
            base = os.path.join(root, "out_" + tag)
            evaluator_utils.save_predictions_box_3d_in_kitti_format(0.1, ds, base, d3, d2, 1200, **kw)
            evaluator_utils.save_predictions_box_3d_in_kitti_format(0.55, ds, base, d3, d2, 1200, **kw)
            out["box_3d/" + tag] = tree_files(base)
        base = os.path.join(root, "out_2d")
        evaluator_utils.save_predictions_box_2d_in_kitti_format(0.30000001, ds, base, d2only, 7)
        out["box_2d"] = tree_files(base)
        rs = np.random.RandomState(3)
        mcfg = types.SimpleNamespace(metrics_to_show=[["metric_cen_z_err", "avg_abs"], ["metric_chamfer", "avg"]])
        for step in (100, 200):
            metrics = {"metric_cen_z_err": list(rs.randn(17)), "metric_chamfer": list(rs.rand(17)), "other": list(rs.randn(5))}
            out["metrics_in/%d" % step] = metrics
            evaluator_utils.save_metrics("ckpt_name", "val", step, metrics, mcfg,
                                         types.SimpleNamespace(add_summary=lambda summary, global_step: None))
        out["metrics"] = tree_files(os.path.join(root, "scripts", "offline_eval", "metrics", "ckpt_name", "val"))
    path = os.path.join(HERE, "evaluator_utils_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
