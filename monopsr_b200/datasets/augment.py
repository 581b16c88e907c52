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


def flip_image(image):
    return np.fliplr(image)


def flip_points(points):
    """(N,3) points mirrored in x"""
    out = np.copy(points)
    out[:, 0] = -points[:, 0]
    return out


def flip_point_cloud(point_cloud):
    """(3,N) point cloud mirrored in x"""
    out = np.copy(point_cloud)
    out[0] = -point_cloud[0]
    return out


def _flip_ry(ry):
    return np.pi - ry if ry >= 0 else -np.pi - ry


def flip_label_in_3d_only(obj_label):
    """mirrored copy of a label: ry and t.x only (the 2-D box is left as is, as in the reference)"""
    out = copy.deepcopy(obj_label)
    out.ry = _flip_ry(obj_label.ry)
    out.t = (-out.t[0], out.t[1], out.t[2])
    return out


def flip_boxes_3d(boxes_3d, flip_ry=True):
    out = np.copy(boxes_3d)
    if flip_ry:
        ry = boxes_3d[:, 6]
        out[:, 6] = np.where(ry >= 0, np.pi - ry, -np.pi - ry)
    out[:, 0] = -boxes_3d[:, 0]
    return out


def flip_ground_plane(ground_plane):
    out = np.copy(ground_plane)
    out[0] = -ground_plane[0]
    return out


def flip_stereo_calib_p2(calib_p2, image_shape):
    """P2 of the mirrored image: principal point reflected about the image width, baseline term negated"""
    out = np.copy(calib_p2)
    out[0, 2] = image_shape[1] - calib_p2[0, 2]
    out[0, 3] = -calib_p2[0, 3]
    return out


def apply_image_noise(image_rgb, rng=np.random):
    """One of: G/B channel swap (p 0.1), per-pixel gaussian (sigma 10), per-channel gaussian (sigma 8), brightness
    (sigma 15), uniform noise of random amplitude < 10 (p 0.4 each).  As in the reference every effect is applied to
    the ORIGINAL image, so when several fire only the last one survives."""
    image_rgb = np.asarray(image_rgb, dtype=np.uint8)
    out = image_rgb
    fire = rng.rand(5)

    def clipped(noise):
        return np.uint8(np.clip(image_rgb + noise, 0.0, 255.0))

    if fire[0] < 0.10:
        out = np.copy(image_rgb)
        # (sic) the reference's tuple assignment of two VIEWS copies channel 2 into 1 and then 1 (already
        # overwritten) back into 2: both end up holding the old channel 2
        out[:, :, 1], out[:, :, 2] = out[:, :, 2], out[:, :, 1]
    if fire[1] < 0.40:
        out = clipped(rng.randn(*image_rgb.shape) * 10.0)
    if fire[2] < 0.40:
        out = clipped(rng.randn(3) * 8.0)
    if fire[3] < 0.40:
        out = clipped(rng.randn(1) * 15.0)
    if fire[4] < 0.40:
        amount = rng.uniform(0, 10)
        out = clipped(rng.uniform(-amount, amount, image_rgb.shape))
    return out


def jitter_obj_boxes_2d(obj_labels, iou_threshold_min, image_shape, rng=np.random):
    """Copies of the labels with randomly shifted / rescaled 2-D boxes: rejection sampling until the new box (clamped to
    the image) overlaps the original with IoU >= iou_threshold_min; boxes under 10 px in either side are left alone."""
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
        iou = 0
        while iou < iou_threshold_min:
            ncx, ncy = rng.normal(cx, half_w / 3), rng.normal(cy, half_h / 3)
            nhw, nhh = rng.normal(half_w, half_w / 6), rng.normal(half_h, half_h / 6)
            x1, x2 = np.maximum(0, ncx - nhw), np.minimum(img_w - 1, ncx + nhw)
            y1, y2 = np.maximum(0, ncy - nhh), np.minimum(img_h - 1, ncy + nhh)
            iou = K.two_d_iou(np.asarray([x1, y1, x2, y2]), original)
        new.x1, new.y1, new.x2, new.y2 = x1, y1, x2, y2
    return out
