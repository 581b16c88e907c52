"""KITTI label / calibration formats and the label-derived part of a training sample (SURVEY.md section 8f rank 4,
partial: the data formats on the input side of the path; the loader, oversampling and augmentation are not built).

Host-side numpy mirrors of
  ObjectLabel, read_labels, filter_labels_by_class      src/monopsr/datasets/kitti/obj_utils.py:85-206
  read_frame_calib                                      src/monopsr/datasets/kitti/calib_utils.py:49-103
  object_label_to_box_2d / object_label_to_box_3d       src/monopsr/core/box_3d_encoder.py:38-72
  get_viewing_angle_box_2d / get_viewing_angle_box_3d   src/monopsr/datasets/kitti/obj_utils.py:913-983
  np_orientation_to_angle_bin                           src/monopsr/core/orientation_encoder.py:11-80
  get_mean_lwh_and_std_dev, class_str_to_index          src/monopsr/datasets/kitti/obj_utils.py:986-1127
  get_prop_cen_z_offset                                 src/monopsr/datasets/kitti/instance_utils.py:972-985
and of the label-derived fields of KittiDataset's sample dict (datasets/kitti/kitti_dataset.py:345-395,450-487),
returned under the key names Engine.set_inputs / model_spec.synthetic_sample use.
Golden vectors from the reference's own functions on its KITTI test fixture: tests/golden/kitti_formats_golden.npz.
"""
import csv
import os

import numpy as np


class ObjectLabel(object):
    def __init__(self):
        self.type = None
        self.truncation = 0.0
        self.occlusion = 0
        self.alpha = 0.0
        self.x1 = self.y1 = self.x2 = self.y2 = 0.0
        self.h = self.w = self.l = 0.0
        self.t = (0.0, 0.0, 0.0)
        self.ry = 0.0
        self.score = 0.0

    def __eq__(self, other):
        return isinstance(other, ObjectLabel) and all(
            np.array_equal(v, other.__dict__[k]) for k, v in self.__dict__.items())


def read_labels(label_dir, sample_name):
    """list of ObjectLabel from <label_dir>/<sample_name>.txt (15 columns, 16 for detection results)"""
    path = os.path.join(label_dir, "{}.txt".format(sample_name))
    if not os.path.exists(path):
        raise FileNotFoundError("Label file could not be found:", path)
    if os.stat(path).st_size == 0:
        return []
    labels = np.loadtxt(path, delimiter=" ", dtype=str, ndmin=2)
    if labels.shape[1] not in (15, 16):
        raise ValueError("Invalid label format")
    out = []
    for row in labels:
        o = ObjectLabel()
        o.type = row[0]
        o.truncation, o.occlusion, o.alpha = float(row[1]), float(row[2]), float(row[3])
        o.x1, o.y1, o.x2, o.y2 = row[4:8].astype(np.float32)
        o.h, o.w, o.l = row[8:11].astype(np.float32)
        o.t = row[11:14].astype(np.float32)
        o.ry = float(row[14])
        o.score = float(row[15]) if labels.shape[1] == 16 else 0.0
        out.append(o)
    return out


def filter_labels_by_class(obj_labels, classes):
    """-> (kept labels, keep mask), as the reference"""
    mask = [(o.type in classes) for o in obj_labels]
    return [o for o, m in zip(obj_labels, mask) if m], mask


class FrameCalib(object):
    def __init__(self):
        self.p0 = self.p1 = self.p2 = self.p3 = self.r0_rect = self.velo_to_cam = None


def read_frame_calib(calib_file_path):
    with open(calib_file_path, "r") as f:
        data = [row for row in csv.reader(f, delimiter=" ")]
    c = FrameCalib()
    ps = [np.reshape([float(v) for v in data[i][1:]], (3, 4)) for i in range(4)]
    c.p0, c.p1, c.p2, c.p3 = ps
    c.r0_rect = np.reshape([float(v) for v in data[4][1:]], (3, 3))
    c.velo_to_cam = np.reshape([float(v) for v in data[5][1:]], (3, 4))
    return c


def object_label_to_box_2d(o):
    return np.asarray([o.y1, o.x1, o.y2, o.x2], np.float32)


def object_label_to_box_3d(o):
    b = np.zeros(7, dtype=np.float32)
    b[0:3] = o.t
    b[3:6] = o.l, o.w, o.h
    b[6] = o.ry
    return b


def get_viewing_angle_box_2d(box_2d, cam_p):
    centre_x = np.mean(box_2d[[1, 3]])
    return np.arctan2((centre_x - cam_p[0, 2]) / cam_p[0, 0], 1.0)


def get_viewing_angle_box_3d(box_3d, cam_p=None, version="x_offset"):
    if version == "cam_0":
        return np.arctan2(box_3d[0], box_3d[2])
    if version == "x_offset":
        x_offset = -cam_p[0, 3] / cam_p[0, 0]
        return np.arctan2(box_3d[0] - x_offset, box_3d[2])
    if version == "projection":
        p = np.dot(cam_p, np.append(box_3d[0:3], 1.0))
        return np.arctan2((p[0] / p[2] - cam_p[0, 2]) / cam_p[0, 0], 1.0)
    raise ValueError("Invalid version", version)


def np_orientation_to_angle_bin(orientation, num_bins, overlap):
    """-> (best bin, residuals to ALL bin centres, one-hot of the valid bins)"""
    two_pi = 2 * np.pi
    wrapped = orientation % two_pi
    per_bin = two_pi / num_bins
    shifted = (wrapped + per_bin / 2) % two_pi
    best = int(shifted / per_bin)
    best_residual = shifted - (best * per_bin + per_bin / 2)
    centres = np.asarray([per_bin * i for i in range(num_bins)])
    residuals = np.arctan2(np.sin(wrapped - centres), np.cos(wrapped - centres))
    valid = [best]
    if overlap != 0.0:
        centre = best * per_bin
        actual = best * per_bin + best_residual
        if np.abs(centre + 0.5 * per_bin - actual) < overlap:
            valid.append(0 if best + 1 == num_bins else best + 1)
        elif np.abs(centre - 0.5 * per_bin - actual) < overlap:
            if best - 1 < 0:                  # (sic) the reference only appends the wrapped-around lower neighbour
                valid.append(num_bins - 1)
    one_hot = np.zeros(num_bins)
    one_hot[np.asarray(valid)] = 1
    return best, residuals, one_hot


def get_mean_lwh_and_std_dev(class_str):
    table = {"Car": ([3.892, 1.619, 1.530], [0.440, 0.106, 0.138]),
             "Pedestrian": ([0.818, 0.628, 1.768], [0.245, 0.122, 0.130]),
             "Cyclist": ([1.771, 0.570, 1.723], [0.153, 0.143, 0.104])}
    if class_str not in table:
        raise ValueError("Invalid class_str", class_str)
    return table[class_str]


def get_prop_cen_z_offset(class_str):
    table = {"Car": 2.17799973487854, "Pedestrian": 0.351921409368515, "Cyclist": 0.8944902420043945}
    if class_str not in table:
        raise ValueError("Invalid class_str", class_str)
    return table[class_str]


def class_str_to_index(class_str, classes):
    if class_str in classes:
        return list(classes).index(class_str) + 1
    raise ValueError("Invalid class string {}, not in {}".format(class_str, classes))


def label_fields(obj_labels, cam_p, image_shape, classes=("Car",), num_alpha_bins=12, alpha_bin_overlap=0.0,
                 num_boxes=32):
    """The label-derived fields of a training sample (kitti_dataset.py:345-395,450-487), padded with zeros to
    `num_boxes` rows as the placeholders require; `num_objs` = rows that are real."""
    labels = filter_labels_by_class(obj_labels, classes)[0][:num_boxes]
    n = len(labels)

    def pad(a, shape, dtype=np.float32):
        out = np.zeros((num_boxes,) + tuple(shape), dtype)
        if n:
            out[:n] = np.asarray(a, dtype).reshape((n,) + tuple(shape))
        return out

    boxes_2d = [object_label_to_box_2d(o) for o in labels]
    boxes_3d = [object_label_to_box_3d(o) for o in labels]
    bins = [np_orientation_to_angle_bin(o.alpha, num_alpha_bins, alpha_bin_overlap) for o in labels]
    strs = [o.type for o in labels]
    b2 = pad(boxes_2d, (4,))
    return {
        "num_objs": n,
        "boxes_2d": b2,
        "boxes_2d_norm": (b2 / np.tile(np.asarray(image_shape, np.float32), 2)).astype(np.float32),
        "boxes_3d": pad(boxes_3d, (7,)),
        "cam_p": np.asarray(cam_p, np.float32),
        "class_indices": pad([class_str_to_index(s, classes) for s in strs], (1,), np.int32),
        "mean_lwh": pad([get_mean_lwh_and_std_dev(s)[0] for s in strs], (3,)),
        "prop_cen_z_offset": pad([get_prop_cen_z_offset(s) for s in strs], ()),
        "est_view_angs": pad([get_viewing_angle_box_2d(b, cam_p) for b in boxes_2d], ()),
        "gt_view_angs": pad([get_viewing_angle_box_3d(b, cam_p) for b in boxes_3d], ()),
        "gt_alphas": pad([o.alpha for o in labels], ()),
        "gt_alpha_bins": pad([b[0] for b in bins], (), np.int32),
        "gt_alpha_regs": pad([b[1] for b in bins], (num_alpha_bins,)),
        "gt_alpha_valid_bins": pad([b[2] for b in bins], (num_alpha_bins,)),
        "label_scores": pad([o.score for o in labels], ()),
    }
