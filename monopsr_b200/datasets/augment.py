"""Training-time augmentation of a KITTI sample -- host-side numpy mirrors of
src/monopsr/datasets/kitti/kitti_aug.py:

  flip_image / flip_points / flip_point_cloud / flip_label_in_3d_only / flip_boxes_3d / flip_ground_plane /
  flip_stereo_calib_p2                                                    kitti_aug.py:11-125
  apply_image_noise                                                       kitti_aug.py:128-171
  jitter_obj_boxes_2d                                                     kitti_aug.py:174-254

The random draws are made in the reference's order from `rng` (default: numpy's global generator, which is what the
reference uses), so a seeded run reproduces the reference's augmented boxes / images value for value
(tests/golden/kitti_loader_golden.npz).
"""
import copy

import numpy as np

from . import kitti_formats as K

AUG_FLIPPING = "flipping"
AUG_PCA_JITTER = "pca_jitter"


def _negated(array, index):
    """copy of `array` with array[index] negated"""
    out = np.array(array, copy=True)
    out[index] = -np.asarray(array)[index]
    return out


def flip_image(image):
    return image[:, ::-1]


def flip_points(points):
    """(N,3) points mirrored in x"""
    return _negated(points, (slice(None), 0))


def flip_point_cloud(point_cloud):
    """(3,N) point cloud mirrored in x"""
    return _negated(point_cloud, 0)


def flip_ground_plane(ground_plane):
    """plane a x + b y + c z + d = 0 mirrored in x"""
    return _negated(ground_plane, 0)


def _mirror_heading(ry):
    """heading of the mirrored object: pi - ry, kept in (-pi, pi] (-pi - ry for negative headings)"""
    ry = np.asarray(ry, dtype=np.float64)
    return np.where(ry >= 0, np.pi, -np.pi) - ry


def flip_label_in_3d_only(obj_label):
    """mirrored copy of a label: ry and t.x only (the 2-D box is left as is, as in the reference)"""
    out = copy.deepcopy(obj_label)
    out.ry = float(_mirror_heading(obj_label.ry))
    out.t = (-out.t[0], out.t[1], out.t[2])
    return out


def flip_boxes_3d(boxes_3d, flip_ry=True):
    out = _negated(boxes_3d, (slice(None), 0))
    if flip_ry:
        out[:, 6] = _mirror_heading(np.asarray(boxes_3d)[:, 6])
    return out


def flip_stereo_calib_p2(calib_p2, image_shape):
    """P2 of the mirrored image: principal point reflected about the image width, baseline term negated"""
    out = _negated(calib_p2, (0, 3))
    out[0, 2] = image_shape[1] - calib_p2[0, 2]
    return out


def apply_image_noise(image_rgb, rng=np.random):
    """One of: G/B channel swap (p 0.1), per-pixel gaussian (sigma 10), per-channel gaussian (sigma 8), brightness
    (sigma 15), uniform noise of random amplitude < 10 (p 0.4 each).  As in the reference every effect is applied to
    the ORIGINAL image, so when several fire only the last one survives; the draws are made in the reference's order
    (five gate values first, then the parameters of each effect that fires)."""
    src = np.asarray(image_rgb, dtype=np.uint8)
    shape = src.shape
    gates = rng.rand(5)

    def swap_gb():
        # (sic) the reference assigns two VIEWS of the same array to each other: channel 1 receives channel 2 and
        # channel 2 then receives the already overwritten channel 1, so both end up holding the old channel 2
        img = src.copy()
        img[:, :, 1] = img[:, :, 2]
        return img

    def additive(draw):
        return lambda: np.uint8(np.clip(src + draw(), 0.0, 255.0))

    def uniform_noise():
        amplitude = rng.uniform(0, 10)
        return rng.uniform(-amplitude, amplitude, shape)

    effects = ((0.10, swap_gb),
               (0.40, additive(lambda: rng.randn(*shape) * 10.0)),
               (0.40, additive(lambda: rng.randn(3) * 8.0)),
               (0.40, additive(lambda: rng.randn(1) * 15.0)),
               (0.40, additive(uniform_noise)))
    out = src
    for gate, (p, effect) in zip(gates, effects):
        if gate < p:
            out = effect()
    return out


def jitter_obj_boxes_2d(obj_labels, iou_threshold_min, image_shape, rng=np.random):
    """Copies of the labels with randomly shifted / rescaled 2-D boxes: rejection sampling until the new box (clamped to
    the image) overlaps the original with IoU >= iou_threshold_min; boxes under 10 px in either side are left alone.
    The acceptance test is core/evaluation.two_d_iou of the reference (unrounded) for one box against one box, written
    out on scalars -- the loop runs ~100 times per sample -- with the same operand types: the candidate in double
    precision, the original box and its area in the precision the label carries (float32 when read from a file)."""
    img_h, img_w = image_shape[0], image_shape[1]
    out = np.empty(len(obj_labels), dtype=object)
    for i, o in enumerate(obj_labels):
        new = copy.deepcopy(o)
        out[i] = new
        half_w, half_h = (o.x2 - o.x1) / 2, (o.y2 - o.y1) / 2
        if o.x2 - o.x1 < 10 or o.y2 - o.y1 < 10:
            continue
        cx, cy = (o.x2 + o.x1) / 2, (o.y2 + o.y1) / 2
        original = np.asarray([[o.x1, o.y1, o.x2, o.y2]])
        ox1, oy1, ox2, oy2 = (float(v) for v in original[0])
        o_area = float((original[0, 2] - original[0, 0]) * (original[0, 3] - original[0, 1]))     # in the label's precision
        iou = 0.0
        while iou < iou_threshold_min:
            ncx, ncy = rng.normal(cx, half_w / 3), rng.normal(cy, half_h / 3)
            nhw, nhh = rng.normal(half_w, half_w / 6), rng.normal(half_h, half_h / 6)
            x1, x2 = max(0.0, ncx - nhw), min(float(img_w - 1), ncx + nhw)
            y1, y2 = max(0.0, ncy - nhh), min(float(img_h - 1), ncy + nhh)
            w, h = min(x2, ox2) - max(x1, ox1), min(y2, oy2) - max(y1, oy1)
            if w > 0 and h > 0:
                inter = w * h
                iou = inter / ((x2 - x1) * (y2 - y1) + o_area - inter)
            else:
                iou = 0.0
        new.x1, new.y1, new.x2, new.y2 = np.float64(x1), np.float64(y1), np.float64(x2), np.float64(y2)
    return out
