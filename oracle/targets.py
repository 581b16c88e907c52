"""ORACLE (test infrastructure, never on the product path): CPU restatement of MonoPSR's ground-truth target
synthesis -- SURVEY.md section 8(f) rank 2, the step immediately BEFORE the network path.

Restates, in numpy float32 with the reference's operation order:
  tf_instance_xyz_crop_from_depth_map   src/monopsr/datasets/kitti/instance_utils.py:395-481
  tf_depth_patch_to_pc_map              src/monopsr/datasets/kitti/depth_map_utils.py:161-236
  tf_get_tr_mat                         src/monopsr/core/transform_utils.py:36-66
as called by MonoPSRModel.build (core/models/monopsr/monopsr_model.py:153-203): once with view_norm=True (local maps
+ valid mask) and once with view_norm=False (global maps) per box.

TensorFlow 1.8 primitives used by the reference, restated from their documented / published kernel semantics (TF is
not installable here, so these two are UNPINNED against TF itself):
  tf.round                       round half to even                           -> np.rint
  tf.image.resize_nearest_neighbor(align_corners=True)
                                 in = min(roundf(out * (in_size-1)/(out_size-1)), in_size-1), roundf = half away from 0
PINNED pieces (tests/test_targets_oracle.py, tests/golden/targets_golden.npz made by tests/golden/make_targets_golden.py
from the reference's own TF-free numpy twins): the pixel-centre back-projection against depth_patch_to_pc_map
(depth_map_utils.py:52-126, use_corr_factors=False, round_box_2d=False) and the view normalisation against
apply_view_norm_to_pc_map / np_get_tr_mat (instance_utils.py:512-536, transform_utils.py:6-33).
"""
import numpy as np

F = np.float32


def tf_linspace(start, stop, num):
    """tf.linspace in float32: start + i * (stop - start) / (num - 1)"""
    start, stop = F(start), F(stop)
    if num == 1:
        return np.array([start], F)
    step = F((stop - start) / F(num - 1))
    return (start + np.arange(num, dtype=F) * step).astype(F)


def roundf_half_away(x):
    return np.where(x >= 0, np.floor(x + F(0.5)), np.ceil(x - F(0.5)))


def resize_nearest_align_corners(src, out_h, out_w):
    """(h, w) -> (out_h, out_w), TF 1.8 ResizeNearestNeighbor with align_corners=True"""
    in_h, in_w = src.shape
    sh = F((in_h - 1) / (out_h - 1)) if out_h > 1 else F(0)
    sw = F((in_w - 1) / (out_w - 1)) if out_w > 1 else F(0)
    ry = np.minimum(roundf_half_away(np.arange(out_h, dtype=F) * sh).astype(np.int64), in_h - 1)
    rx = np.minimum(roundf_half_away(np.arange(out_w, dtype=F) * sw).astype(np.int64), in_w - 1)
    return src[np.ix_(ry, rx)]


def depth_patch_to_pc_map(depth_patch, box_2d, cam_p, roi_size):
    """depth_map_utils.py:161-236 with use_pixel_centres=True: (3, roi_h, roi_w) camera-frame points"""
    y1, x1, y2, x2 = [F(v) for v in box_2d]
    nx, ny = roi_size[0], roi_size[1]                      # sic: x count from roi_size[0] (:200-201)
    pw, ph = F((x2 - x1) / F(nx)), F((y2 - y1) / F(ny))
    hw, hh = F(pw / F(2.0)), F(ph / F(2.0))
    xx, yy = np.meshgrid(tf_linspace(x1 + hw, x2 - hw, nx), tf_linspace(y1 + hh, y2 - hh, ny))
    f, cu, cv = F(cam_p[0, 0]), F(cam_p[0, 2]), F(cam_p[1, 2])
    ratio = (depth_patch / f).astype(F)
    return np.stack([((xx - cu) * ratio).astype(F), ((yy - cv) * ratio).astype(F), depth_patch.astype(F)], 0)


def tr_mat(ry, t):
    """transform_utils.py:36-66: rot_y(ry) @ translate(t), float32"""
    c, s = F(np.cos(F(ry))), F(np.sin(F(ry)))
    rot = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], F)
    tm = np.array([[1, 0, 0, t[0]], [0, 1, 0, t[1]], [0, 0, 1, t[2]], [0, 0, 0, 1]], F)
    return (rot @ tm).astype(F)


def instance_xyz_crop_from_depth_map(box_idx, boxes_2d, boxes_3d, instance_masks, depth_map, roi_size, viewing_angles,
                                     cam_p, view_norm, centroid_type="bottom", rotate_view=True):
    """instance_utils.py:395-481 for one box -> (xyz (roi_h, roi_w, 3), valid (roi_h, roi_w, 1))"""
    box_2d = boxes_2d[box_idx].astype(F)
    b = np.rint(box_2d).astype(np.int32)                                        # :422 tf.round -> half to even
    masked = (depth_map.astype(F) * instance_masks[box_idx].astype(F)).astype(F)  # :425
    crop = masked[b[0]:b[2], b[1]:b[3]]                                          # :426-428
    d = resize_nearest_align_corners(crop, roi_size[0], roi_size[1])             # :429-430
    pc = depth_patch_to_pc_map(d, box_2d, cam_p, roi_size)                       # :433-434 (unrounded box)
    valid = (np.abs(d) >= F(0.1)).astype(F)[..., None]                           # :437-438
    if view_norm:
        x_off = F(-F(cam_p[0, 3]) / F(cam_p[0, 0]))                              # :446
        cen = boxes_3d[box_idx, 0:3].astype(F) - np.array([x_off, 0, 0], F)
        if centroid_type == "middle":
            cen = cen - np.array([0, F(boxes_3d[box_idx, 5]) / F(2.0), 0], F)    # :449-452
        if rotate_view:
            tm = tr_mat(-F(viewing_angles[box_idx]), -cen)                       # :455
        else:
            tm = np.eye(4, dtype=F)
            tm[:3, 3] = -cen
        p = np.concatenate([pc.reshape(3, -1), np.ones((1, pc[0].size), F)], 0)  # :466-467
        xyz = (tm @ p)[0:3].T.reshape(roi_size[0], roi_size[1], 3).astype(F)     # :470-471
    else:
        xyz = pc.reshape(3, -1).T.reshape(roi_size[0], roi_size[1], 3)           # :478-479
    return (xyz * valid).astype(F), valid


def gt_maps(boxes_2d, boxes_3d, instance_masks, depth_map, viewing_angles, cam_p, roi=48, centroid_type="middle",
            rotate_view=True):
    """monopsr_model.py:165-203: (local (N,roi,roi,3), global (N,roi,roi,3), valid (N,roi,roi,1))"""
    loc, glo, val = [], [], []
    for i in range(len(boxes_2d)):
        a, v = instance_xyz_crop_from_depth_map(i, boxes_2d, boxes_3d, instance_masks, depth_map, (roi, roi),
                                                viewing_angles, cam_p, True, centroid_type, rotate_view)
        g, _ = instance_xyz_crop_from_depth_map(i, boxes_2d, boxes_3d, instance_masks, depth_map, (roi, roi),
                                                viewing_angles, cam_p, False, centroid_type, rotate_view)
        loc.append(a); glo.append(g); val.append(v)
    return np.stack(loc), np.stack(glo), np.stack(val)


# ------------------------------------------------------------------------------------------------ image inputs
# ImgPreprocessor.preprocess_input (core/img_preprocessor.py:12-35) + the two consumers in MonoPSRModel.build
# (monopsr_model.py:222-233).  TF image-op semantics restated (UNPINNED against TF itself; the crop_and_resize
# restatement is the one oracle/network.py uses for the feature-map crops).
KITTI_CHANNEL_MEANS = np.array([92.8403, 97.7996, 93.5843], F)


def _bilinear(img, sy, sx):
    """img (H,W,C); sy (oh,), sx (ow,) source coordinates -> (oh, ow, C); lerp order as TF: x first, then y"""
    H, W, _ = img.shape
    y0 = np.floor(sy).astype(np.int64); x0 = np.floor(sx).astype(np.int64)
    y1 = np.minimum(y0 + 1, H - 1); x1 = np.minimum(x0 + 1, W - 1)
    ly = (sy - y0).astype(F)[:, None, None]; lx = (sx - x0).astype(F)[None, :, None]
    tl, tr = img[np.ix_(y0, x0)], img[np.ix_(y0, x1)]
    bl, br = img[np.ix_(y1, x0)], img[np.ix_(y1, x1)]
    t = tl + (tr - tl) * lx
    b = bl + (br - bl) * lx
    return (t + (b - t) * ly).astype(F)


def preprocess_input(rgb_image, out_hw, means=KITTI_CHANNEL_MEANS):
    """float(img) - means, then tf.image.resize_images (bilinear, align_corners=False, TF1 legacy mapping)"""
    img = rgb_image.astype(F) - means.astype(F)
    H, W, _ = img.shape
    oh, ow = out_hw
    sy = np.arange(oh, dtype=F) * (F(H) / F(oh))
    sx = np.arange(ow, dtype=F) * (F(W) / F(ow))
    return _bilinear(img, sy, sx)


def crop_and_resize(img, boxes_norm, crop):
    """tf.image.crop_and_resize (bilinear, extrapolation 0), one image (H,W,C) -> (N,crop,crop,C)"""
    H, W, C = img.shape
    out = np.zeros((len(boxes_norm), crop, crop, C), F)
    for n, (y1, x1, y2, x2) in enumerate(boxes_norm.astype(F)):
        hs, ws = (y2 - y1) * F(H - 1) / F(crop - 1), (x2 - x1) * F(W - 1) / F(crop - 1)
        in_y = y1 * F(H - 1) + np.arange(crop, dtype=F) * hs
        in_x = x1 * F(W - 1) + np.arange(crop, dtype=F) * ws
        vy, vx = (in_y >= 0) & (in_y <= H - 1), (in_x >= 0) & (in_x <= W - 1)
        t = np.clip(np.floor(in_y), 0, H - 1).astype(np.int64); b = np.clip(np.ceil(in_y), 0, H - 1).astype(np.int64)
        l = np.clip(np.floor(in_x), 0, W - 1).astype(np.int64); r = np.clip(np.ceil(in_x), 0, W - 1).astype(np.int64)
        ly = (in_y - np.floor(in_y)).astype(F)[:, None, None]; lx = (in_x - np.floor(in_x)).astype(F)[None, :, None]
        tt = img[np.ix_(t, l)] + (img[np.ix_(t, r)] - img[np.ix_(t, l)]) * lx
        bb = img[np.ix_(b, l)] + (img[np.ix_(b, r)] - img[np.ix_(b, l)]) * lx
        o = tt + (bb - tt) * ly
        out[n] = np.where((vy[:, None] & vx[None, :])[..., None], o, 0).astype(F)
    return out


def resize_bilinear_ac(img, out_hw):
    """tf.image.resize_bilinear(align_corners=True), (H,W,C) -> (oh,ow,C)"""
    H, W, _ = img.shape
    oh, ow = out_hw
    sy = np.arange(oh, dtype=F) * (F(H - 1) / F(oh - 1))
    sx = np.arange(ow, dtype=F) * (F(W - 1) / F(ow - 1))
    return _bilinear(img.astype(F), sy, sx)


def image_inputs(rgb_image, boxes_norm, image_input_shape=(320, 1216), roi=48, full_shape=(160, 608)):
    pre = preprocess_input(rgb_image, image_input_shape)
    return crop_and_resize(pre, boxes_norm, roi), resize_bilinear_ac(pre, full_shape)[None], pre
