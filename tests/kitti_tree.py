"""A small synthetic KITTI object tree for the loader tests, built deterministically from the three label / calibration
files of the reference's test fixture that are committed under tests/golden/kitti (the fixture itself carries no depth
or instance images, so the reference's own loader test cannot run on it).  Used by BOTH the golden generator
(tests/golden/make_kitti_loader_golden.py, which runs the reference's KittiDataset on this tree) and the tests (which
run monopsr_b200.datasets.kitti_loader on an identical tree)."""
import os
import shutil

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "golden", "kitti")
H, W = 375, 1242
# new sample name -> committed label / calib it is a copy of
TRAINING = {"000001": "000001", "000008": "000008", "000076": "000076", "000108": "000008", "000176": "000076",
            "000208": "000008", "000301": "000001"}
SPLITS = {"train": ["000001", "000008", "000076", "000108", "000176", "000208", "000301"],
          "val": ["000008", "000108", "000076"],
          "trainval": sorted(TRAINING),
          "test": ["000008", "000108"]}
MSCNN_THR = "0.2_0.2_0.2"


def _rows(path):
    return [r.split(" ") for r in open(path).read().splitlines() if r]


def _image(rng):
    coarse = rng.randint(0, 256, (H // 25 + 1, W // 27 + 1, 3)).astype(np.uint8)
    return np.ascontiguousarray(np.repeat(np.repeat(coarse, 25, axis=0), 27, axis=1)[:H, :W])      # numpy only:
    # the golden file holds checksums of these images, so their content must not depend on a library's resampler


def _depth(rng):
    coarse = rng.uniform(0.0, 60.0, (H // 15 + 1, W // 23 + 1))
    coarse[rng.rand(*coarse.shape) < 0.15] = 0.04          # below the 10 cm cut -> "no depth"
    d = np.repeat(np.repeat(coarse, 15, axis=0), 23, axis=1)[:H, :W]
    return np.ascontiguousarray((d * 256.0).astype(np.uint16))


def _instances(rows, rng):
    inst = np.full((H, W), 255, np.uint8)
    for i, r in enumerate(rows):
        x1, y1, x2, y2 = [float(v) for v in r[4:8]]
        mx, my = 0.1 * (x2 - x1), 0.1 * (y2 - y1)
        xs, xe = int(round(x1 + mx)), int(round(x2 - mx))
        ys, ye = int(round(y1 + my)), int(round(y2 - my))
        block = inst[ys:ye, xs:xe]
        block[rng.rand(*block.shape) < 0.8] = i
    return inst


def _detections(rows, rng, classes=("Car", "Pedestrian", "Cyclist")):
    """MS-CNN style results (16 columns): most labelled objects re-detected with a slightly moved box and a score,
    every third one moved far enough to miss the IoU gate, plus one false positive"""
    out = []
    k = 0
    for r in rows:
        if r[0] not in classes:
            continue
        x1, y1, x2, y2 = [float(v) for v in r[4:8]]
        w, h = x2 - x1, y2 - y1
        far = (k % 3 == 2)
        k += 1
        sx, sy = (0.5 * w, 0.4 * h) if far else (rng.uniform(-0.04, 0.04) * w, rng.uniform(-0.04, 0.04) * h)
        box = [max(0.0, x1 + sx), max(0.0, y1 + sy), min(W - 1.0, x2 + sx), min(H - 1.0, y2 + sy)]
        out.append([r[0], "-1", "-1", r[3]] + ["%.2f" % v for v in box] + r[8:15] + ["%.4f" % rng.uniform(0.3, 1.0)])
    out.append(["Car", "-1", "-1", "0.50", "20.00", "200.00", "80.00", "240.00", "1.50", "1.60", "3.90", "-20.00", "1.70",
                "30.00", "0.10", "0.2500"])
    return out


def make_tree(root, seed=0):
    """-> (dataset_dir, data_dir).  Layout: <dataset_dir>/{train,val,trainval,test}.txt, training/ and testing/ with
    image_2, calib, label_2 (training only), depth_2_multiscale, instance_2_depth_2_multiscale;
    <data_dir>/detections/mscnn/kitti_fmt/<split>/merged_0.2_0.2_0.2/data."""
    rng = np.random.RandomState(seed)
    dataset_dir, data_dir = os.path.join(root, "Kitti", "object"), os.path.join(root, "data")
    for split, names in SPLITS.items():
        os.makedirs(dataset_dir, exist_ok=True)
        with open(os.path.join(dataset_dir, split + ".txt"), "w") as f:
            f.write("\n".join(names) + "\n")
    with open(os.path.join(dataset_dir, "readme.txt"), "w") as f:
        f.write("synthetic\n")
    for sub, names in (("training", sorted(TRAINING)), ("testing", SPLITS["test"])):
        base = os.path.join(dataset_dir, sub)
        for d in ("image_2", "calib", "label_2", "depth_2_multiscale", "instance_2_depth_2_multiscale"):
            os.makedirs(os.path.join(base, d), exist_ok=True)
        for name in names:
            src = TRAINING[name]
            rows = _rows(os.path.join(SRC, "label_2", src + ".txt"))
            shutil.copy(os.path.join(SRC, "calib", src + ".txt"), os.path.join(base, "calib", name + ".txt"))
            if sub == "training":
                shutil.copy(os.path.join(SRC, "label_2", src + ".txt"), os.path.join(base, "label_2", name + ".txt"))
            cv2.imwrite(os.path.join(base, "image_2", name + ".png"), _image(rng))
            cv2.imwrite(os.path.join(base, "depth_2_multiscale", name + ".png"), _depth(rng))
            cv2.imwrite(os.path.join(base, "instance_2_depth_2_multiscale", name + ".png"), _instances(rows, rng))
    for split, names in SPLITS.items():
        det_dir = os.path.join(data_dir, "detections", "mscnn", "kitti_fmt", split, "merged_" + MSCNN_THR, "data")
        os.makedirs(det_dir, exist_ok=True)
        for name in names:
            rows = _rows(os.path.join(SRC, "label_2", TRAINING[name] + ".txt"))
            with open(os.path.join(det_dir, name + ".txt"), "w") as f:
                f.write("\n".join(" ".join(r) for r in _detections(rows, rng)) + "\n")
    return dataset_dir, data_dir


# the loader configurations the golden file covers: name -> (mode, config overrides)
CASES = {
    "train_default": ("train", {}),
    "train_all_noise": ("train", {"aug_config.box_jitter_type": "all", "aug_config.use_image_aug": True}),
    "train_gt_jitter": ("train", {"aug_config.box_jitter_type": "oversample_gt"}),
    "train_plain": ("train", {"aug_config.box_jitter_type": None, "oversample": False, "use_mscnn_detections": False}),
    "train_ped": ("train", {"classes": ["Pedestrian"], "num_alpha_bins": 8, "alpha_bin_overlap": 0.1}),
    "val_mscnn": ("val", {"data_split": "val"}),
    "val_kitti": ("val", {"data_split": "val", "use_mscnn_detections": False}),
    "test": ("test", {"data_split": "test", "data_split_dir": "testing", "has_kitti_labels": False}),
}
BASE_CONFIG = dict(num_boxes=32, num_alpha_bins=12, alpha_bin_overlap=0.0)      # monopsr_model_000.yaml values


def apply_overrides(cfg, dataset_dir, overrides):
    for k, v in dict(BASE_CONFIG, dataset_dir=dataset_dir, **overrides).items():
        obj, parts = cfg, k.split(".")
        for p in parts[:-1]:
            obj = getattr(obj, p)
        setattr(obj, parts[-1], v)
    return cfg


def summarize(sample_dict):
    """sample_dict -> flat dict of small arrays (images / depth / masks as crc32 + shape), None -> {'none': 1}"""
    import zlib
    if sample_dict is None:
        return {"none": np.asarray(1)}
    out = {}
    for k, v in sample_dict.items():
        a = np.asarray(v)
        if a.size > 4096:
            a = np.ascontiguousarray(a)
            out[k + "__crc"] = np.asarray([zlib.crc32(a.tobytes())] + list(a.shape), np.int64)
            out[k + "__dtype"] = np.asarray(str(a.dtype))
        else:
            out[k] = a
    return out


def write_predictions(root, sample_names, seed=11):
    """'%0.5f' prediction files as core/predictions.save_predictions writes them, for the evaluator_utils converters:
    <root>/box_3d (x y z l w h ry score class), <root>/box_2d (y1 x1 y2 x2 alpha score class), <root>/box_2d_only
    (y1 x1 y2 x2 score class).  One sample gets no file, one an empty file, one only low scores."""
    rng = np.random.RandomState(seed)
    d3, d2, d2o = (os.path.join(root, d) for d in ("box_3d", "box_2d", "box_2d_only"))
    for d in (d3, d2, d2o):
        os.makedirs(d, exist_ok=True)
    for i, name in enumerate(sample_names):
        if i == 1 and len(sample_names) > 3:
            continue                                  # no prediction file at all
        n = 0 if i == 2 else int(rng.randint(1, 7))
        x = rng.uniform(-25, 25, n)
        z = rng.uniform(4, 60, n)
        b3 = np.column_stack([x, rng.uniform(1.2, 2.0, n), z, rng.uniform(3, 5, n), rng.uniform(1.4, 2, n),
                              rng.uniform(1.3, 1.8, n), rng.uniform(-3.1, 3.1, n),
                              rng.uniform(0.0, 0.09, n) if i == 0 else rng.uniform(0, 1, n), np.zeros(n)])
        y1, x1 = rng.uniform(100, 250, n), rng.uniform(0, 1000, n)
        b2 = np.column_stack([y1, x1, y1 + rng.uniform(20, 120, n), x1 + rng.uniform(20, 200, n),
                              rng.uniform(-3.1, 3.1, n), b3[:, 7], np.zeros(n)])
        np.savetxt(os.path.join(d3, name + ".txt"), b3, fmt="%0.5f")
        np.savetxt(os.path.join(d2, name + ".txt"), b2, fmt="%0.5f")
        np.savetxt(os.path.join(d2o, name + ".txt"), b2[:, [0, 1, 2, 3, 5, 6]], fmt="%0.5f")
    return d3, d2, d2o
