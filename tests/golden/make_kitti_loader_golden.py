"""Generate tests/golden/kitti_loader_golden.npz by running the reference's OWN loader -- the unmodified
monopsr.datasets.kitti.kitti_dataset.KittiDataset, kitti_aug and obj_utils imported from /root/reference/src -- on the
synthetic KITTI tree of tests/kitti_tree.py, with numpy's global generator seeded per case.

The reference modules import tensorflow and pypng, neither installed here and neither touched by the loader: both are
satisfied by an import hook that hands out empty stub modules.  yaml.load is given its (since then mandatory) Loader.
Run from the repository root:  python tests/golden/make_kitti_loader_golden.py
"""
import importlib.abc
import importlib.machinery
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
        if name.split(".")[0] in ("tensorflow", "png"):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, "/root/reference/src")
    _load = yaml.load
    yaml.load = lambda s, Loader=yaml.SafeLoader: _load(s, Loader=Loader)
    import monopsr
    from monopsr.builders.dataset_builder import DatasetBuilder
    from monopsr.core import evaluation
    from monopsr.datasets.kitti import evaluation as kitti_evaluation
    from monopsr.datasets.kitti import kitti_aug, obj_utils

    out = {}
    with tempfile.TemporaryDirectory() as root:
        dataset_dir, data_dir = kitti_tree.make_tree(root)
        monopsr.data_dir = lambda: data_dir
        for case, (mode, overrides) in kitti_tree.CASES.items():
            cfg = kitti_tree.apply_overrides(DatasetBuilder.get_config_obj(DatasetBuilder.KITTI_TRAIN), dataset_dir, overrides)
            np.random.seed(1234)
            ds = DatasetBuilder.build_kitti_dataset(cfg, mode)
            samples = ds.get_sample_dict(np.arange(ds.num_samples))
            out[case + "/n"] = np.asarray(len(samples))
            for i, s in enumerate(samples):
                for k, v in kitti_tree.summarize(s).items():
                    out["%s/%d/%s" % (case, i, k)] = v
            # epoch bookkeeping: names served and the counters after each call
            np.random.seed(99)
            ds = DatasetBuilder.build_kitti_dataset(cfg, mode)
            trace = []
            for bs in (3, 3, 2, 5, ds.num_samples, 1):
                bs = min(bs, ds.num_samples)        # (the reference indexes past the list for larger batches)
                batch = ds.next_batch(bs, shuffle=(mode == "train"))
                trace.append([bs, len(batch), ds._index_in_epoch, ds.epochs_completed])
                out["%s/batch%d/names" % (case, len(trace) - 1)] = np.asarray(
                    ["" if s is None else s["sample_name"] for s in batch])
            out[case + "/trace"] = np.asarray(trace)

        # stand-alone pieces
        labels = obj_utils.read_labels(os.path.join(dataset_dir, "training", "label_2"), "000008")
        for name, kw in (("hard", dict(difficulty=2)), ("easy", dict(difficulty=0)), ("moderate_h60", dict(difficulty=1, box_2d_height=60)),
                         ("occ2", dict(occlusion=2)), ("trunc", dict(truncation=0.3)), ("depth", dict(depth_range=[5, 15])),
                         ("cars_all", dict(classes=["Car"], difficulty=3))):
            out["filter/" + name] = np.asarray(obj_utils.filter_labels(labels, **kw)[1])
        rs = np.random.RandomState(5)
        boxes = rs.uniform(0, 100, (40, 2))
        boxes = np.hstack([boxes, boxes + rs.uniform(1, 60, (40, 2))]).astype(np.float32)
        out["iou/boxes"] = boxes
        out["iou/values"] = np.asarray([evaluation.two_d_iou(b, boxes) for b in boxes])
        out["iou/values_rounded"] = np.asarray([kitti_evaluation.two_d_iou(b, boxes) for b in boxes])
        np.random.seed(7)
        jit = kitti_aug.jitter_obj_boxes_2d(labels, 0.7, (kitti_tree.H, kitti_tree.W))
        out["jitter/boxes"] = np.asarray([[o.x1, o.y1, o.x2, o.y2] for o in jit], np.float64)
        img = (np.arange(20 * 30 * 3) % 251).astype(np.uint8).reshape(20, 30, 3)
        out["noise/image"] = img
        for seed in range(12):
            np.random.seed(seed)
            out["noise/%d" % seed] = kitti_aug.apply_image_noise(img)
        kitti = obj_utils.read_labels(os.path.join(dataset_dir, "training", "label_2"), "000008")
        det = obj_utils.read_labels(os.path.join(data_dir, "detections/mscnn/kitti_fmt/val/merged_0.2_0.2_0.2/data"), "000008")
        for st in ("distance", "max", "min"):
            merged = obj_utils.merge_kitti_and_mscnn_obj_labels(kitti, det, 0.7, default_score_type=st)
            out["merge/" + st] = np.asarray([[o.x1, o.y1, o.x2, o.y2, o.score] for o in merged], np.float64)
        b3 = np.asarray([[1, 2, 3, 4, 5, 6, 0.3], [-2, 1, 9, 4, 2, 1, -2.0], [0, 0, 5, 1, 1, 1, 0.0]])
        out["flip/boxes_3d_in"], out["flip/boxes_3d"] = b3, kitti_aug.flip_boxes_3d(b3)
        out["flip/boxes_3d_noflip"] = kitti_aug.flip_boxes_3d(b3, flip_ry=False)
        p2 = np.arange(12, dtype=np.float64).reshape(3, 4) + 0.5
        out["flip/p2_in"], out["flip/p2"] = p2, kitti_aug.flip_stereo_calib_p2(p2, (375, 1242))
        fl = kitti_aug.flip_label_in_3d_only(labels[1])
        out["flip/label"] = np.asarray([fl.ry, fl.t[0], fl.t[1], fl.t[2], fl.x1], np.float64)
    path = os.path.join(HERE, "kitti_loader_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
