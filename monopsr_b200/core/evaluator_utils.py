"""After the forward pass of an evaluation run -- host-side mirror of
src/monopsr/core/evaluator_utils.py:

  save_predictions_box_2d_in_kitti_format     :17-111    '%0.5f' box_2d files -> KITTI result files (score filter,
                                                         (x1,y1,x2,y2) order, alpha placeholder -10, 3 decimals, CRLF)
  save_predictions_box_3d_in_kitti_format     :114-277   box_3d + box_2d files -> KITTI result files, optionally with
                                                         the 2-D box re-projected from the 3-D one
  save_metrics                                :280-403   avg / std / avg|.| / std|.| csv rows per evaluated checkpoint
  print_inference_time_statistics             :437-454
  run_kitti_native_eval                       :457-535   the reference compiles and shells out to the KITTI C++ evaluator
                                                         and tees its output into results/<split>/<name>_results_<thr>.txt;
                                                         here core/kitti_eval.py (native: csrc/kitti_eval.cu) is called
                                                         in process and the same "AP" lines are appended to that file.
`dataset` is a datasets.kitti_loader.KittiDataset (data_split, num_samples, sample_list, classes, calib_dir,
get_rgb_image_path).  Golden files from the reference's own functions: tests/golden/make_evaluator_utils_golden.py."""
import csv
import os
import sys

import numpy as np

from . import kitti_eval
from . import predictions as P
from ..datasets import kitti_formats as K


def _kitti_dir(predictions_base_dir, data_split, score_threshold, global_step):
    d = predictions_base_dir + "/kitti_predictions_3d/{}/{}/{}/data".format(data_split, score_threshold, global_step)
    os.makedirs(d, exist_ok=True)
    return d


def _progress(i, n, log):
    if log is not None:
        log("\rConverting {} / {}".format(i + 1, n))


def _stdout(s):
    sys.stdout.write(s)
    sys.stdout.flush()


def save_predictions_box_2d_in_kitti_format(score_threshold, dataset, predictions_base_dir, predictions_box_2d_dir,
                                            global_step, log=_stdout):
    """rows of (y1, x1, y2, x2, score, class) -> '<type> -1000 ... -10 x1 y1 x2 y2 -1000 ... score'.  (As in the
    reference the files land in the kitti_predictions_3d tree.)  Returns the output directory."""
    score_threshold = round(score_threshold, 3)
    out_dir = _kitti_dir(predictions_base_dir, dataset.data_split, score_threshold, global_step)
    n_valid = 0
    for idx in range(dataset.num_samples):
        _progress(idx, dataset.num_samples, log)
        name = dataset.sample_list[idx].name
        out_path, in_path = out_dir + "/" + name + ".txt", predictions_box_2d_dir + "/" + name + ".txt"
        if not os.path.exists(in_path):
            np.savetxt(out_path, [])
            continue
        pred = np.loadtxt(in_path).reshape(-1, 6)
        pred[:, 0:4] = pred[:, [1, 0, 3, 2]]
        pred = pred[pred[:, 4] >= score_threshold]
        if len(pred) == 0:
            np.savetxt(out_path, [])
            continue
        n_valid += 1
        kitti = np.full([pred.shape[0], 16], -1000.0)
        kitti[:, 3] = -10.0
        kitti[:, 4:8] = pred[:, 0:4]
        kitti[:, 15] = pred[:, 4]
        kitti = np.round(kitti, 3)
        types = [dataset.classes[c] for c in pred[:, 5].astype(np.int32)]
        np.savetxt(out_path, np.column_stack([types, kitti[:, 1:16]]), newline="\r\n", fmt="%s")
    if log is not None:
        log("\nNum valid: {}\nNum samples: {}\n".format(n_valid, dataset.num_samples))
    return out_dir


def save_predictions_box_3d_in_kitti_format(score_threshold, dataset, predictions_base_dir, predictions_box_3d_dir,
                                            predictions_box_2d_dir, global_step, project_3d_box=False, log=_stdout):
    """rows of (x, y, z, l, w, h, ry, score, class) + the 2-D rows (y1, x1, y2, x2, alpha, ...) of the same objects ->
    '<type> -1 -1 alpha x1 y1 x2 y2 h w l x y z ry score'.  Returns the output directory."""
    score_threshold = round(score_threshold, 3)
    out_dir = _kitti_dir(predictions_base_dir, dataset.data_split, score_threshold, global_step)
    n_valid = 0
    for idx in range(dataset.num_samples):
        _progress(idx, dataset.num_samples, log)
        name = dataset.sample_list[idx].name
        out_path = out_dir + "/" + name + ".txt"
        p3_path, p2_path = predictions_box_3d_dir + "/" + name + ".txt", predictions_box_2d_dir + "/" + name + ".txt"
        if not os.path.exists(p3_path):
            np.savetxt(out_path, [])
            continue
        p3 = np.loadtxt(p3_path)
        if len(p3) == 0:
            np.savetxt(out_path, [])
            continue
        p3 = p3.reshape(-1, 9)
        p2 = np.loadtxt(p2_path).reshape(-1, 7)
        keep = p3[:, 7] >= score_threshold
        p3, p2 = p3[keep], p2[keep]
        if len(p3) == 0:
            np.savetxt(out_path, [])
            continue
        if project_3d_box:
            from PIL import Image
            image_size = Image.open(dataset.get_rgb_image_path(name)).size
            cam_p = K.read_frame_calib(dataset.calib_dir + "/{}.txt".format(name)).p2
            boxes, inside = [], []
            for row in p3:
                b = P.project_to_image_space(row[0:7], cam_p, truncate=True, image_size=image_size)
                inside.append(b is not None)
                if b is not None:
                    boxes.append(b)
            boxes_2d = np.asarray(boxes)
            p3, p2 = p3[inside], p2[inside]
        else:
            boxes_2d = p2[:, [1, 0, 3, 2]]
        if len(p3) == 0:
            np.savetxt(out_path, [])
            continue
        n_valid += 1
        kitti = np.zeros([len(p3), 16])
        kitti[:, 3] = p2[:, 4]
        kitti[:, 4:8] = boxes_2d
        kitti[:, 8], kitti[:, 9], kitti[:, 10] = p3[:, 5], p3[:, 4], p3[:, 3]        # h, w, l
        kitti[:, 11:14] = p3[:, 0:3]
        kitti[:, 14:16] = p3[:, 6:8]
        kitti = np.round(kitti, 3)
        types = [dataset.classes[c] for c in p3[:, 8].astype(np.int32)]
        empty = -1 * np.ones((len(kitti), 2), dtype=np.int32)
        np.savetxt(out_path, np.column_stack([types, empty, kitti[:, 3:16]]), newline="\r\n", fmt="%s")
    if log is not None:
        log("\nNum valid: {}\nNum samples: {}\n".format(n_valid, dataset.num_samples))
    return out_dir


def save_metrics(metrics_dir, data_split, global_step, metrics_dict, metrics_to_show=(), summary_fn=None):
    """append one row per statistic to metrics_{avg,std,avg_abs,std_abs}_<split>.csv in `metrics_dir` (the reference:
    <scripts>/offline_eval/metrics/<checkpoint_name>/<split>/); metrics_dict: name -> list of per-object values;
    metrics_to_show: [[name, 'avg'|'std'|'avg_abs'|'std_abs'], ...] (model_config.metrics_to_show) and
    summary_fn(tag, value, global_step) stand in for the tensorboard scalars."""
    os.makedirs(metrics_dir, exist_ok=True)
    kinds = ("avg", "std", "avg_abs", "std_abs")
    names = sorted(metrics_dict.keys())
    header = ["step".rjust(8)] + [(n[7:] if n.startswith("metric") else n).rjust(12) for n in names]
    step = "{}".format(global_step).rjust(8)
    rows = {k: [step] for k in kinds}
    show = np.asarray(metrics_to_show) if len(metrics_to_show) else np.zeros((0, 2), dtype=str)
    for key in names:
        v = metrics_dict[key]
        stat = {"avg": np.mean(v), "std": np.std(v), "avg_abs": np.mean(np.abs(v)), "std_abs": np.std(np.abs(v))}
        for k in kinds:
            rows[k].append("{:.5f}".format(stat[k]).rjust(12))
        for cfg in show[show[:, 0] == key] if len(show) else []:
            if cfg[1] not in stat:
                raise ValueError("Invalid show_metric_type", cfg[1])
            if summary_fn is not None:
                summary_fn("metrics/{}/".format(cfg[1]) + key, stat[cfg[1]], global_step)
    for k in kinds:
        path = os.path.join(metrics_dir, "metrics_{}_{}.csv".format(k, data_split))
        with open(path, "a") as f:
            w = csv.writer(f, delimiter=",")
            if os.stat(path).st_size == 0:
                w.writerow(header)
            w.writerow(rows[k])


def print_inference_time_statistics(total_feed_dict_time, total_inference_time, log=print):
    for title, t in (("Feed dict time:", total_feed_dict_time), ("Inference time:", total_inference_time)):
        t = np.asarray(t)
        log(title)
        log("Min:  {}".format(np.round(np.min(t), 5)))
        log("Max:  {}".format(np.round(np.max(t), 5)))
        log("Mean:  {}".format(np.round(np.mean(t), 5)))
        log("Median:  {}".format(np.round(np.median(t), 5)))


def run_kitti_native_eval(checkpoint_name, data_split, kitti_score_threshold, global_step, label_dir, predictions_base_dir,
                          results_root, low_iou=False, log=print):
    """evaluate <predictions_base_dir>/kitti_predictions_3d/<split>/<thr>/<step> against `label_dir` and append the
    step and the AP lines to <results_root>/results[_low_iou]/<split>/<checkpoint_name>_results_<thr>.txt (what
    run_eval.sh / run_eval_low_iou.sh tee).  Returns core.kitti_eval.evaluate's dict."""
    kitti_score_threshold = round(kitti_score_threshold, 3)
    pred_dir = predictions_base_dir + "/kitti_predictions_3d/{}/{}/{}".format(data_split, kitti_score_threshold, global_step)
    results_dir = os.path.join(results_root, "results_low_iou" if low_iou else "results", data_split)
    os.makedirs(results_dir, exist_ok=True)
    res = kitti_eval.evaluate(label_dir, pred_dir, low_iou=low_iou, write_stats=True)
    with open(os.path.join(results_dir, "{}_results_{}.txt".format(checkpoint_name, kitti_score_threshold)), "a") as f:
        f.write("{}\n".format(global_step))
        for ln in res["lines"]:
            f.write(ln + "\n")
    if log is not None:
        log(str(global_step))
        for ln in res["lines"]:
            log(ln)
    return res
