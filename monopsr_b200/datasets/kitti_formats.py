"""KITTI label / calibration / depth / instance formats, object filters and the label-derived part of a training sample
(the data formats on the input side of the path; the loader itself, oversampling and
augmentation are in datasets/kitti_loader.py and datasets/augment.py).

Host-side numpy mirrors of
  ObjectLabel, read_labels, filter_labels_by_class      src/monopsr/datasets/kitti/obj_utils.py:85-206
  read_frame_calib                                      src/monopsr/datasets/kitti/calib_utils.py:49-103
  object_label_to_box_2d / object_label_to_box_3d       src/monopsr/core/box_3d_encoder.py:38-72
  get_viewing_angle_box_2d / get_viewing_angle_box_3d   src/monopsr/datasets/kitti/obj_utils.py:913-983
  np_orientation_to_angle_bin                           src/monopsr/core/orientation_encoder.py:11-80
  get_mean_lwh_and_std_dev, class_str_to_index          src/monopsr/datasets/kitti/obj_utils.py:986-1127
  get_prop_cen_z_offset                                 src/monopsr/datasets/kitti/instance_utils.py:972-985
and of the label-derived fields of KittiDataset's sample dict (datasets/kitti/kitti_dataset.py:345-395,450-487),
returned under the key names Engine.set_inputs / model_spec.synthetic_sample use, plus
  Difficulty, ObjectFilter, filter_labels*, apply_obj_filter   src/monopsr/datasets/kitti/obj_utils.py:12-15,25-82,193-368
  two_d_iou                                             src/monopsr/core/evaluation.py:23-61 and
                                                        src/monopsr/datasets/kitti/evaluation.py:6-44 (rounded)
  merge_kitti_and_mscnn_obj_labels                      src/monopsr/datasets/kitti/obj_utils.py:1037-1089
  read_depth_map / read_instance_image / get_instance_mask_list
                                                        depth_map_utils.py:9-17, instance_utils.py:10-44
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

    def __deepcopy__(self, memo):
        # every field is an immutable scalar / string except the position `t` (the box jitter copies ~30 labels per sample)
        new = ObjectLabel.__new__(ObjectLabel)
        new.__dict__.update(self.__dict__)
        new.t = np.array(self.t, copy=True) if isinstance(self.t, np.ndarray) else tuple(self.t)
        return new


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


# ---------------------------------------------------------------------------------------------- object filters
class Difficulty(object):
    EASY, MODERATE, HARD, ALL = 0, 1, 2, 3
    _NAMES = ("easy", "moderate", "hard", "all")

    @staticmethod
    def to_string(difficulty):
        return Difficulty._NAMES[difficulty]

    @staticmethod
    def from_string(difficulty_str):
        if difficulty_str not in Difficulty._NAMES:
            raise KeyError(difficulty_str)
        return Difficulty._NAMES.index(difficulty_str)


# KITTI difficulty thresholds, indexed by Difficulty (easy, moderate, hard)
_MIN_HEIGHT, _MAX_OCCLUSION, _MAX_TRUNCATION = (40, 25, 25), (0, 1, 2), (0.15, 0.3, 0.5)


class ObjectFilter(object):
    """classes / difficulty / min 2-D box height / max truncation / max occlusion / depth range; None = not applied"""

    def __init__(self, config):
        self.classes = config.classes
        self.difficulty = Difficulty.from_string(config.difficulty_str)
        self.box_2d_height = config.box_2d_height
        self.truncation = config.truncation
        self.occlusion = config.occlusion
        self.depth_range = config.depth_range

    @staticmethod
    def create_obj_filter(classes, difficulty, occlusion, truncation, box_2d_height, depth_range):
        import types
        return ObjectFilter(types.SimpleNamespace(
            classes=classes, difficulty_str=Difficulty.to_string(difficulty), occlusion=occlusion, truncation=truncation,
            box_2d_height=box_2d_height, depth_range=depth_range))


def _as_label_array(obj_labels):
    a = np.empty(len(obj_labels), dtype=object)
    for i, o in enumerate(obj_labels):
        a[i] = o
    return a


def filter_labels(obj_labels, classes=None, difficulty=None, box_2d_height=None, occlusion=None, truncation=None,
                  depth_range=None):
    """-> (kept labels as an object array, boolean keep mask); the comparisons are the reference's: difficulty uses
    <= / >= on the KITTI thresholds, the explicit limits are strict (height >, truncation <, occlusion <, depth open)."""
    keep = np.full(len(obj_labels), True)
    for i, o in enumerate(obj_labels):
        ok = True
        if classes is not None:
            ok &= o.type in classes
        if difficulty is not None and difficulty != Difficulty.ALL:
            ok &= bool(o.occlusion <= _MAX_OCCLUSION[difficulty] and o.truncation <= _MAX_TRUNCATION[difficulty]
                       and (o.y2 - o.y1) >= _MIN_HEIGHT[difficulty])
        if box_2d_height is not None:
            ok &= bool((o.y2 - o.y1) > box_2d_height)
        if occlusion is not None:
            ok &= bool(o.occlusion < occlusion)
        if truncation is not None:
            ok &= bool(o.truncation < truncation)
        if depth_range is not None:
            ok &= bool(depth_range[0] < o.t[2] < depth_range[1])
        keep[i] = ok
    return _as_label_array(obj_labels)[keep], keep


def apply_obj_filter(obj_labels, obj_filter):
    return filter_labels(obj_labels, classes=obj_filter.classes, difficulty=obj_filter.difficulty,
                         box_2d_height=obj_filter.box_2d_height, occlusion=obj_filter.occlusion,
                         truncation=obj_filter.truncation, depth_range=obj_filter.depth_range)


def boxes_2d_from_obj_labels(obj_labels):
    return np.asarray([object_label_to_box_2d(o) for o in obj_labels], np.float32)


def boxes_3d_from_obj_labels(obj_labels):
    return np.asarray([object_label_to_box_3d(o) for o in obj_labels], np.float32)


def two_d_iou(box, boxes, decimals=None):
    """IoU of one box against (N,4) boxes, any consistent corner order.  The reference has two copies of this function:
    core/evaluation.py:23-61 returns the quotient as is (used by the box jitter's acceptance test), while
    datasets/kitti/evaluation.py:6-44 rounds it to 3 decimals (used when matching MS-CNN detections): decimals=3."""
    boxes = np.asarray(boxes)
    lo = np.maximum(box[:2], boxes[:, :2])
    hi = np.minimum(box[2:4], boxes[:, 2:4])
    wh = hi - lo
    hit = (wh[:, 0] > 0) & (wh[:, 1] > 0)
    iou = np.zeros(len(boxes), np.float64)
    if hit.any():
        inter = wh[hit, 0] * wh[hit, 1]
        area = (box[2] - box[0]) * (box[3] - box[1])
        areas = (boxes[hit, 2] - boxes[hit, 0]) * (boxes[hit, 3] - boxes[hit, 1])
        iou[hit] = inter / (area + areas - inter)
    return iou if decimals is None else iou.round(decimals)


def merge_kitti_and_mscnn_obj_labels(kitti_obj_labels, mscnn_obj_labels, min_iou, default_score_type="distance"):
    """KITTI labels whose 2-D box and score are replaced by the best-overlapping MS-CNN detection (IoU >= min_iou);
    labels left without a score get one from their depth ('distance'), 1 ('max') or 0 ('min')."""
    import copy
    merged = copy.deepcopy(kitti_obj_labels)
    kitti_boxes = boxes_2d_from_obj_labels(kitti_obj_labels)
    for det in mscnn_obj_labels:
        det_box = object_label_to_box_2d(det)
        ious = two_d_iou(det_box, kitti_boxes, decimals=3)
        best = int(np.argmax(ious))
        if ious[best] >= min_iou:
            m = merged[best]
            m.y1, m.x1, m.y2, m.x2 = det_box
            m.score = det.score
    for m in merged:
        if m.score == 0:
            if default_score_type == "distance":
                m.score = np.clip(1.0 - (m.t[2] / 45.0), 0.1, 1.0)
            elif default_score_type in ("max", "min"):
                m.score = 1.0 if default_score_type == "max" else 0.0
            else:
                raise ValueError("Invalid default score type", default_score_type)
    return merged


# ---------------------------------------------------------------------------------------------- image-like inputs
def _imread(path, flag):
    import cv2
    img = cv2.imread(path, flag)
    if img is None:
        raise FileNotFoundError("Image could not be read:", path)
    return img


def read_rgb_image(path):
    """uint8 (H,W,3) RGB (cv2 decodes BGR; kitti_dataset.py:252-253)"""
    import cv2
    return _imread(path, cv2.IMREAD_COLOR)[..., ::-1]


def read_depth_map(depth_map_path):
    """uint16 png, metres x 256 -> float32 metres; depths under 10 cm are 'no depth' (0)"""
    import cv2
    depth = _imread(depth_map_path, cv2.IMREAD_ANYDEPTH) / 256.0
    depth[depth < 0.1] = 0.0
    return depth.astype(np.float32)


def write_depth_map(depth_map_path, depth_map, png_compression=3):
    import cv2
    cv2.imwrite(depth_map_path, (np.asarray(depth_map) * 256.0).astype(np.uint16),
                [cv2.IMWRITE_PNG_COMPRESSION, png_compression])


def read_instance_image(instance_image_path):
    """uint8 (H,W): pixel = index of the object label it belongs to, 255 = none"""
    import cv2
    return _imread(instance_image_path, cv2.IMREAD_GRAYSCALE)


def get_instance_mask_list(instance_img, num_instances=None):
    """(k,H,W) boolean masks, one per label index; without `num_instances`, k = highest index present + 1"""
    if num_instances is None:
        present = instance_img[instance_img != 255]
        if len(present) == 0:
            return []
        num_instances = int(np.max(present)) + 1
    return np.asarray([instance_img == i for i in range(num_instances)])
