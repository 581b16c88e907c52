"""Synthetic evaluation cases for the AP evaluator tests: a ground-truth label directory and a results directory, built
deterministically from the three KITTI label files committed under tests/golden/kitti.  Shared by the golden generator
(tests/golden/make_kitti_eval_golden.py: runs the reference's evaluator binary of oracle/_ref on them) and the tests."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "golden", "kitti", "label_2")
H, W = 375, 1242

# name -> knobs of the synthetic detector
CASES = {
    "good": dict(seed=1, frames=90, p_detect=0.9, box_noise=0.03, pose_noise=0.05, n_false=1, alpha=True),
    "noisy": dict(seed=2, frames=120, p_detect=0.75, box_noise=0.12, pose_noise=0.35, n_false=3, alpha=True),
    "no_alpha": dict(seed=3, frames=60, p_detect=0.8, box_noise=0.06, pose_noise=0.15, n_false=2, alpha=False),
    "sparse": dict(seed=4, frames=45, p_detect=0.3, box_noise=0.05, pose_noise=0.1, n_false=0, alpha=True),
    "perfect": dict(seed=5, frames=60, p_detect=1.0, box_noise=0.0, pose_noise=0.0, n_false=0, alpha=True),
}


def _rows(name):
    return [r.split(" ") for r in open(os.path.join(SRC, name + ".txt")).read().splitlines() if r]


def _fmt(row):
    return " ".join([row[0]] + ["%.2f" % float(v) for v in row[1:]])


def make_case(root, case):
    """-> (gt_dir, result_dir): ground truth = shifted / rescaled copies of the committed label files (objects of other
    classes, Vans, sitting persons and DontCare areas included), results = a noisy re-detection of them"""
    k = CASES[case]
    rng = np.random.RandomState(k["seed"])
    gt_dir, res_dir = os.path.join(root, case, "label_2"), os.path.join(root, case, "results")
    os.makedirs(gt_dir, exist_ok=True)
    os.makedirs(os.path.join(res_dir, "data"), exist_ok=True)
    base = [_rows(n) for n in ("000001", "000008", "000076")]
    for f in range(k["frames"]):
        rows = [list(r) for r in base[f % 3]]
        dx, dz, sc = rng.uniform(-40, 40), rng.uniform(0, 12), rng.uniform(0.7, 1.3)
        gt, det = [], []
        for r in rows:
            r = list(r)
            if r[0] != "DontCare":
                cx = (float(r[4]) + float(r[6])) / 2 + dx
                cy = (float(r[5]) + float(r[7])) / 2
                hw, hh = (float(r[6]) - float(r[4])) / 2 * sc, (float(r[7]) - float(r[5])) / 2 * sc
                r[4], r[5] = max(0.0, cx - hw), max(0.0, cy - hh)
                r[6], r[7] = min(W - 1.0, cx + hw), min(H - 1.0, cy + hh)
                r[13] = float(r[13]) + dz
                if rng.rand() < 0.12:                 # some neighbouring-class objects
                    r[0] = {"Car": "Van", "Pedestrian": "Person_sitting"}.get(r[0], r[0])
                r[2] = int(rng.choice([0, 0, 1, 2, 3])) if rng.rand() < 0.3 else int(float(r[2]))
            gt.append(_fmt(r[:2]) + " %d " % int(float(r[2])) + " ".join("%.2f" % float(v) for v in r[3:15]))
            if r[0] in ("Car", "Pedestrian", "Cyclist", "Van") and rng.rand() < k["p_detect"]:
                w_, h_ = float(r[6]) - float(r[4]), float(r[7]) - float(r[5])
                n = k["box_noise"]
                box = [float(r[4]) + rng.randn() * n * w_, float(r[5]) + rng.randn() * n * h_,
                       float(r[6]) + rng.randn() * n * w_, float(r[7]) + rng.randn() * n * h_]
                p = k["pose_noise"]
                hwl = [float(v) * (1 + rng.randn() * p * 0.2) for v in r[8:11]]
                t = [float(r[11]) + rng.randn() * p, float(r[12]) + rng.randn() * p * 0.2, float(r[13]) + rng.randn() * p * 2]
                ry = float(r[14]) + rng.randn() * p
                alpha = float(r[3]) + rng.randn() * p if k["alpha"] else -10.0
                name = "Car" if r[0] == "Van" and rng.rand() < 0.5 else r[0]
                det.append([name, -1, -1, alpha] + box + hwl + t + [ry, rng.uniform(0.05, 1.0)])
        for _ in range(k["n_false"]):
            x1, y1 = rng.uniform(0, W - 120), rng.uniform(100, H - 90)
            name = ["Car", "Pedestrian", "Cyclist"][rng.randint(3)]
            det.append([name, -1, -1, rng.uniform(-3, 3) if k["alpha"] else -10.0, x1, y1, x1 + rng.uniform(15, 110),
                        y1 + rng.uniform(15, 80), 1.5, 1.6, 3.9, rng.uniform(-20, 20), 1.6, rng.uniform(5, 60),
                        rng.uniform(-3, 3), rng.uniform(0.05, 0.9)])
        with open(os.path.join(gt_dir, "%06d.txt" % f), "w") as fh:
            fh.write("\n".join(gt) + "\n")
        with open(os.path.join(res_dir, "data", "%06d.txt" % f), "w") as fh:
            for d in det:
                fh.write(d[0] + " -1 -1 " + " ".join("%.3f" % float(v) for v in d[3:]) + "\r\n")
    return gt_dir, res_dir
