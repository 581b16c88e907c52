"""Generate tests/golden/kitti_eval_golden.json: output of the reference's OWN KITTI evaluator -- compiled from its
source by oracle/build_ref.sh into oracle/_ref/evaluate_object_3d_offline(_low_iou); boost::geometry comes from
oracle/boost_shim -- on the synthetic cases of tests/kitti_eval_cases.py: the "AP" lines it prints and the precision /
orientation curves it writes (stats_*.txt).  Run from the repository root after `bash oracle/build_ref.sh`."""
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
import kitti_eval_cases  # noqa: E402


def run_reference(binary, gt_dir, res_dir):
    out = subprocess.run([binary, gt_dir + "/", res_dir], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         universal_newlines=True).stdout
    lines = [ln for ln in out.splitlines() if " AP: " in ln]
    stats = {}
    for f in sorted(os.listdir(res_dir)):
        if f.startswith("stats_"):
            stats[f] = [[float(v) for v in ln.split()] for ln in open(os.path.join(res_dir, f)).read().splitlines()]
    return {"lines": lines, "stats": stats}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as root:
        for case in kitti_eval_cases.CASES:
            for low in (False, True):
                binary = os.path.join(ROOT, "oracle", "_ref", "evaluate_object_3d_offline" + ("_low_iou" if low else ""))
                gt_dir, res_dir = kitti_eval_cases.make_case(os.path.join(root, "low" if low else "std"), case)
                out[case + ("/low_iou" if low else "")] = run_reference(binary, gt_dir, res_dir)
    path = os.path.join(HERE, "kitti_eval_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    for k, v in out.items():
        print(k, *v["lines"], sep="\n   ")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
