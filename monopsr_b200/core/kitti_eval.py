"""KITTI object-detection AP evaluation (the step after the prediction writers).

Host side of the native evaluator (csrc/kitti_eval.cu, C ABI include/monopsr_b200_eval.h) that replaces the binary the
reference compiles and shells out to (scripts/offline_eval/kitti_native_eval/evaluate_object_3d_offline.cpp through
run_eval.sh, src/monopsr/core/evaluator_utils.py:457-535).  This module does what that program's `eval` / `main` do
around `eval_class` (:846-1006): find the evaluated frames (getEvalIndices :830-844), parse ground truth and results
(loadGroundtruth :177-205, loadDetections :130-175 -- which also decides per class whether 2-D, bird's-eye-view and 3-D
boxes are evaluated, and switches orientation scoring off when any detection carries alpha = -10), run the three
difficulties per class and metric, and report "<class>_<what> AP: easy moderate hard" lines (:746-755) plus the
stats_*.txt precision curves (:207-223).  The gnuplot / mail parts are not reproduced.
"""
import ctypes
import os

import numpy as np

from .. import lib as _lib

CLASS_NAMES = ("car", "pedestrian", "cyclist")
TYPE_CODES = {"car": 0, "pedestrian": 1, "cyclist": 2, "van": 3, "person_sitting": 4, "dontcare": 5}
OTHER = 6
MIN_OVERLAP = {False: (0.7, 0.5, 0.5), True: (0.5, 0.25, 0.25)}       # per class; True = the "low IoU" evaluator
N_SAMPLE_PTS = 41
IMAGE, GROUND, BOX3D = 0, 1, 2
_c_d = ctypes.POINTER(ctypes.c_double)
_c_i = ctypes.POINTER(ctypes.c_int)


def _lib_eval():
    L = _lib.load()
    if not getattr(L, "_mpb_eval_ready", False):
        L.mpb_kitti_eval_class.restype = ctypes.c_int
        L.mpb_kitti_eval_class.argtypes = [ctypes.c_int, _c_i, _c_d, _c_i, _c_d, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_int, ctypes.c_int, _c_d, _c_d, _c_d, _c_i, _c_i]
        L.mpb_kitti_overlap.restype = ctypes.c_double
        L.mpb_kitti_overlap.argtypes = [_c_d, _c_d, ctypes.c_int, ctypes.c_int]
        L._mpb_eval_ready = True
    return L


def parse_objects(path, with_score):
    """rows of 15 (ground truth) or 16 (results) doubles, the object type replaced by its code.  Like the reference's
    fscanf loop this reads whitespace-separated tokens (line breaks do not matter) and stops at the first record that
    does not parse."""
    with open(path, "r") as f:
        tok = f.read().split()
    n = 16 if with_score else 15
    rows = []
    for i in range(0, len(tok) - n + 1, n):
        try:
            vals = [float(x) for x in tok[i + 1:i + n]]
        except ValueError:
            break
        if not with_score:
            vals[1] = float(int(vals[1]))        # occlusion is read with %d
        rows.append([float(TYPE_CODES.get(tok[i].lower(), OTHER))] + vals)
    return np.asarray(rows, np.float64).reshape(-1, n)


def eval_indices(result_dir):
    """frame numbers of the files in <result_dir>/data (the last 10 characters of a name are '<6 digits>.txt')"""
    out = []
    for name in os.listdir(os.path.join(result_dir, "data")):
        if len(name) >= 10:
            digits = ""
            for ch in name[-10:]:               # atoi: leading digits only
                if not ch.isdigit():
                    break
                digits += ch
            out.append(int(digits) if digits else 0)
    return sorted(out)


def overlap(det_row, gt_row, metric, criterion=-1):
    d = np.ascontiguousarray(det_row, np.float64)
    g = np.ascontiguousarray(gt_row, np.float64)
    if d.size < 15 or g.size < 15:
        raise ValueError("rows of at least 15 values expected")
    return float(_lib_eval().mpb_kitti_overlap(d.ctypes.data_as(_c_d), g.ctypes.data_as(_c_d), metric, criterion))


def eval_class(gt_list, det_list, cls, difficulty, metric, min_overlap, compute_aos=False, compute_aos_ground=False):
    """one (class, difficulty, metric): -> dict(precision (41,), aos, aos_ground, n_thresholds, n_gt)"""
    if len(gt_list) != len(det_list):
        raise ValueError("one ground-truth and one detection array per image expected")
    n = len(gt_list)
    gt_off = np.zeros(n + 1, np.int32)
    det_off = np.zeros(n + 1, np.int32)
    for i in range(n):
        gt_off[i + 1] = gt_off[i] + len(gt_list[i])
        det_off[i + 1] = det_off[i] + len(det_list[i])
    gt = np.ascontiguousarray(np.concatenate([np.reshape(g, (-1, 15)) for g in gt_list] + [np.zeros((0, 15))]), np.float64)
    det = np.ascontiguousarray(np.concatenate([np.reshape(d, (-1, 16)) for d in det_list] + [np.zeros((0, 16))]), np.float64)
    prec, aos, aosg = (np.zeros(N_SAMPLE_PTS) for _ in range(3))
    nt, ngt = ctypes.c_int(0), ctypes.c_int(0)
    st = _lib_eval().mpb_kitti_eval_class(
        n, gt_off.ctypes.data_as(_c_i), gt.ctypes.data_as(_c_d), det_off.ctypes.data_as(_c_i), det.ctypes.data_as(_c_d),
        cls, difficulty, metric, float(min_overlap), int(compute_aos), int(compute_aos_ground),
        prec.ctypes.data_as(_c_d), aos.ctypes.data_as(_c_d), aosg.ctypes.data_as(_c_d), ctypes.byref(nt), ctypes.byref(ngt))
    _lib.check(st, "mpb_kitti_eval_class")
    return {"precision": prec, "aos": aos if compute_aos else None, "aos_ground": aosg if compute_aos_ground else None,
            "n_thresholds": nt.value, "n_gt": ngt.value}


def average_precision(curve):
    """the reference's 11-point figure: every 4th of the 41 samples, accumulated in single precision, in percent"""
    s = np.float32(0)
    for i in range(0, len(curve), 4):
        s = np.float32(np.float64(s) + np.float64(curve[i]))      # float += double
    return float(np.float32(s / np.float32(11) * np.float32(100)))


def evaluate(gt_dir, result_dir, low_iou=False, write_stats=False, log=None):
    """-> dict(ap={'car_detection': [easy, moderate, hard], ...}, curves={name: (3, 41) array}, lines=[...]).
    gt_dir holds KITTI label files, <result_dir>/data the result files of the frames to evaluate."""
    gts, dets = [], []
    for idx in eval_indices(result_dir):
        name = "%06d.txt" % idx
        gt_path, det_path = os.path.join(gt_dir, name), os.path.join(result_dir, "data", name)
        if not os.path.exists(gt_path):
            raise FileNotFoundError("ERROR: Couldn't read: %s of ground truth." % name)
        gts.append(parse_objects(gt_path, False))
        dets.append(parse_objects(det_path, True))
    all_det = np.concatenate(dets + [np.zeros((0, 16))])
    compute_aos = not bool(np.any(all_det[:, 3] == -10))
    eval_image, eval_ground, eval_3d = [], [], []
    for c in range(3):
        d = all_det[all_det[:, 0] == c]
        eval_image.append(bool(np.any(d[:, 4] >= 0)))
        ok_g = (d[:, 11] != -1000) & (d[:, 13] != -1000) & (d[:, 9] > 0) & (d[:, 10] > 0)
        eval_ground.append(bool(np.any(ok_g)))
        eval_3d.append(bool(np.any(ok_g & (d[:, 12] != -1000) & (d[:, 8] > 0))))
    ap, curves, lines = {}, {}, []
    suffix = "_low_iou" if low_iou else ""

    def report(name, vals):
        curves[name] = np.asarray(vals)
        ap[name] = [average_precision(v) for v in vals]
        lines.append("%s AP: %f %f %f" % ((name,) + tuple(ap[name])))

    def stats(fname, rows):
        if write_stats:
            with open(os.path.join(result_dir, fname), "w") as f:
                for r in rows:
                    f.write("".join("%f " % x for x in r) + "\n")

    plans = ((IMAGE, eval_image, compute_aos, False), (GROUND, eval_ground, False, True), (BOX3D, eval_3d, False, True))
    for metric, enabled, aos_on, aosg_on in plans:
        for c in range(3):
            if not enabled[c]:
                continue
            res = [eval_class(gts, dets, c, diff, metric, MIN_OVERLAP[low_iou][c], aos_on, aosg_on) for diff in range(3)]
            prec = [r["precision"] for r in res]
            if metric == IMAGE:
                stats("stats_%s_detection%s.txt" % (CLASS_NAMES[c], suffix), prec)
                report(CLASS_NAMES[c] + "_detection", prec)
                if aos_on:
                    stats("stats_%s_orientation%s.txt" % (CLASS_NAMES[c], suffix), [r["aos"] for r in res])
                    report(CLASS_NAMES[c] + "_orientation", [r["aos"] for r in res])
            else:
                tag = "BEV" if metric == GROUND else "3D"
                stats("stats_%s_detection_ground.txt" % CLASS_NAMES[c], prec)      # (sic) one name for both metrics
                report(CLASS_NAMES[c] + "_detection_" + tag, prec)
                report(CLASS_NAMES[c] + "_heading_" + tag, [r["aos_ground"] for r in res])
    if log is not None:
        for ln in lines:
            log(ln)
    return {"ap": ap, "curves": curves, "lines": lines}


def main(argv=None):
    """python -m monopsr_b200.core.kitti_eval [--low_iou] gt_dir result_dir   (the reference binary's command line,
    evaluate_object_3d_offline.cpp:971-1004: prints the result folder's name, then the AP lines)"""
    import sys
    args = list(sys.argv[1:] if argv is None else argv)
    low = "--low_iou" in args
    args = [a for a in args if a != "--low_iou"]
    if len(args) != 2:
        print("Usage: python -m monopsr_b200.core.kitti_eval [--low_iou] gt_dir result_dir")
        return 1
    print(os.path.basename(os.path.normpath(args[1])))
    evaluate(args[0], args[1], low_iou=low, write_stats=True, log=print)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
