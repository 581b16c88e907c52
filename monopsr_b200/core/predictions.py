"""Prediction formatting and writers (the step immediately AFTER the network path).

Host-side numpy, as in the reference: turns the network's output_dict (Engine.outputs(), keys of core/constants.py
KEY_*) plus the sample_dict into KITTI-style 3-D / 2-D detections and writes them.  Mirrors, function by function:
  np_angle_bin_to_orientation   src/monopsr/core/orientation_encoder.py:83-107
  compute_box_3d_corners        src/monopsr/datasets/kitti/obj_utils.py:835-864 (= compute_obj_label_corners_3d :623-654)
  project_pc_to_image           src/monopsr/datasets/kitti/calib_utils.py:245-260
  postprocess_cen_x             src/monopsr/datasets/kitti/instance_utils.py:988-1032
  project_to_image_space        src/monopsr/core/box_3d_projector.py:14-88
  score_boxes                   src/monopsr/core/models/monopsr/monopsr_output_builder.py:805-860
  format_predictions            src/monopsr/core/models/monopsr/monopsr_model.py:960-1073
  save_predictions              src/monopsr/core/models/monopsr/monopsr_model.py:1075-1102
Differences from the reference are confined to plumbing: the camera matrix is passed in (the reference re-reads the
calibration file per box, monopsr_output_builder.py:823) and configuration comes as keyword arguments instead of
attributes of the TF model object.  Golden vectors from the reference's own functions: tests/golden/predictions_golden.npz.
"""
import os

import numpy as np

# core/constants.py keys used here
KEY_VALID_MASK_MAPS = "valid_mask_maps"
KEY_INST_XYZ_MAP_LOCAL = "inst_xyz_map_local"
KEY_BOX_2D, KEY_BOX_3D = "box_2d", "box_3d"
KEY_VIEW_ANG, KEY_LWH, KEY_ALPHA = "view_ang", "lwh", "alpha"
KEY_ALPHA_BINS, KEY_ALPHA_REGS, KEY_CENTROIDS = "alpha_bins", "alpha_regs", "centroids"
SAMPLE_NAME, SAMPLE_IMAGE_INPUT, SAMPLE_NUM_OBJS = "sample_name", "sample_image_input", "sample_num_objs"
SAMPLE_CAM_P, SAMPLE_LABEL_SCORES = "sample_cam_p", "sample_label_scores"
SAMPLE_LABEL_BOXES_2D, SAMPLE_LABEL_BOXES_3D = "sample_label_boxes_2d", "sample_label_boxes_3d"
SAMPLE_VIEWING_ANGLES_3D, SAMPLE_LABEL_CLASS_INDICES = "sample_viewing_angles_3d", "sample_label_class_indices"
OUT_DIR_XYZ_MAP_LOCAL, OUT_DIR_BOX_2D, OUT_DIR_BOX_3D = "output_xyz_map_dir", "output_box_2d_dir", "output_box_3d_dir"


def np_angle_bin_to_orientation(angle_bin, residual, num_bins):
    two_pi = 2 * np.pi
    angle = angle_bin * (two_pi / num_bins) + residual
    if angle < -np.pi:
        angle = angle + two_pi
    if angle > np.pi:
        angle = angle - two_pi
    return angle


def compute_box_3d_corners(box_3d):
    """(3, 8) corners of [x, y, z, l, w, h, ry] (y is the BOTTOM face, KITTI convention)"""
    tx, ty, tz, l, w, h, ry = box_3d
    rot = np.array([[+np.cos(ry), 0, +np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, +np.cos(ry)]])
    x = np.array([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2])
    y = np.array([0, 0, 0, 0, -h, -h, -h, -h])
    z = np.array([w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2])
    c = np.dot(rot, np.array([x, y, z]))
    c[0, :] += tx
    c[1, :] += ty
    c[2, :] += tz
    return c


def project_pc_to_image(point_cloud, cam_p):
    pc_padded = np.append(point_cloud, np.ones((1, point_cloud.shape[1])), axis=0)
    pts = np.dot(cam_p, pc_padded)
    pts[0:2] = pts[0:2] / pts[2]
    return pts[0:2]


def postprocess_cen_x(pred_box_2d, pred_box_3d, cam_p):
    """centroid x from the u-ratio of the projected centroid inside the projected 3-D box, applied to the 2-D box"""
    focal_length, centre_u = cam_p[0, 0], cam_p[0, 2]
    corners_uv = project_pc_to_image(compute_box_3d_corners(pred_box_3d), cam_p)
    cen_uv = project_pc_to_image(pred_box_3d[0:3, np.newaxis], cam_p)
    min_u, max_u = np.amin(corners_uv[0]), np.amax(corners_uv[0])
    ratio_u = (cen_uv[0] - min_u) / (max_u - min_u)
    u = pred_box_2d[1] + ratio_u * (pred_box_2d[3] - pred_box_2d[1])
    return (u - centre_u) * (pred_box_3d[2] / focal_length)


def project_to_image_space(box_3d, calib_p2, truncate=False, image_size=None, discard=True,
                           discard_before_truncation=True):
    """[x1, y1, x2, y2] of the projected 3-D box, or None when it is outside / too large"""
    projected = project_pc_to_image(compute_box_3d_corners(box_3d), calib_p2)
    img_box = np.array([np.amin(projected[0]), np.amin(projected[1]), np.amax(projected[0]), np.amax(projected[1])])
    if truncate:
        if not image_size:
            raise ValueError("Image size must be provided")
        image_w, image_h = image_size[0], image_size[1]
        if img_box[0] > image_w or img_box[1] > image_h or img_box[2] < 0 or img_box[3] < 0:
            return None
        if discard and discard_before_truncation:
            if (img_box[2] - img_box[0]) > image_w * 0.8 or (img_box[3] - img_box[1]) > image_h * 0.8:
                return None
        img_box[0] = max(img_box[0], 0)
        img_box[1] = max(img_box[1], 0)
        img_box[2] = min(img_box[2], image_w)
        img_box[3] = min(img_box[3], image_h)
        if discard and not discard_before_truncation:
            if (img_box[2] - img_box[0]) > image_w * 0.8 and (img_box[3] - img_box[1]) > image_h * 0.8:
                return None
    return img_box


def score_boxes(cam_p, img_shape, boxes_2d, boxes_3d, valid_scores, max_depth=45.0):
    """0.95 * detector score + 0.05 * mean(depth score, fit between the projected 3-D box and the 2-D detection)"""
    all_new_scores = np.zeros_like(valid_scores)
    for idx, (box_2d, box_3d) in enumerate(zip(boxes_2d, boxes_3d)):
        proj = project_to_image_space(box_3d, cam_p, truncate=True, image_size=(img_shape[1], img_shape[0]))
        b = np.asarray(box_2d)[[1, 0, 3, 2]]          # box_3d_encoder.boxes_2d_to_iou_fmt: [x1, y1, x2, y2]
        if proj is None:
            fit = 0.1
        else:
            height, width = b[3] - b[1], b[2] - b[0]
            fit = 1.0 - (np.abs((b[0] - proj[0]) / width) + np.abs((b[2] - proj[2]) / width) +
                         np.abs((b[1] - proj[1]) / height) + np.abs((b[3] - proj[3]) / height))
        depth_score = np.clip(1.0 - (box_3d[2] / max_depth), 0.1, 1.0)
        all_new_scores[idx] = 0.95 * valid_scores[idx] + 0.05 * ((depth_score + fit) / 2.0)
    return all_new_scores


def format_predictions(output_dict, sample_dict, *, output_types, train_val_test="val", num_boxes=32, num_alpha_bins=12,
                       centroid_type="middle", post_process_cen_x=True, alpha_type="dc"):
    """monopsr_model.py:960-1073.  output_dict values are numpy arrays (Engine.outputs() moved to the host)."""
    img = sample_dict[SAMPLE_IMAGE_INPUT]
    num_objs = sample_dict[SAMPLE_NUM_OBJS]
    cam_p = sample_dict[SAMPLE_CAM_P]
    valid_scores = np.expand_dims(sample_dict[SAMPLE_LABEL_SCORES][0:num_objs], 1)
    pred = {}
    valid_mask_maps = (output_dict[KEY_VALID_MASK_MAPS][0:num_objs] > 0.0).astype(np.float32)
    pred[KEY_VALID_MASK_MAPS] = valid_mask_maps
    if KEY_INST_XYZ_MAP_LOCAL in output_types:
        pred[KEY_INST_XYZ_MAP_LOCAL] = output_dict[KEY_INST_XYZ_MAP_LOCAL][0:num_objs] * valid_mask_maps
    if KEY_CENTROIDS in output_types:
        new_boxes_2d = np.copy(sample_dict[SAMPLE_LABEL_BOXES_2D])
        if train_val_test in ("train", "val"):
            new_boxes_3d = np.copy(sample_dict[SAMPLE_LABEL_BOXES_3D])
        elif train_val_test == "test":
            new_boxes_3d = np.zeros([num_boxes, 7], dtype=np.float32)
        else:
            raise ValueError("Invalid run mode", train_val_test)
        if KEY_LWH in output_types:
            new_boxes_3d[:, 3:6] = output_dict[KEY_LWH]
        if KEY_VIEW_ANG in output_types:
            view_angs = output_dict[KEY_VIEW_ANG]
        else:
            view_angs = sample_dict[SAMPLE_VIEWING_ANGLES_3D]
        if KEY_ALPHA in output_types:
            if alpha_type in ("dc", "dc_rotation", "gt"):
                bins = np.argmax(output_dict[KEY_ALPHA_BINS], axis=1)
                regs = [output_dict[KEY_ALPHA_REGS][i, b] for i, b in enumerate(bins)]
                pred_alphas = [np_angle_bin_to_orientation(b, r, num_alpha_bins) for b, r in zip(bins, regs)]
            elif alpha_type == "prob":
                pred_alphas = np.squeeze(output_dict[KEY_ALPHA])
            else:
                raise ValueError("Invalid alpha_type", alpha_type)
            pred_alphas = np.asarray(pred_alphas)
            new_boxes_3d[:, 6] = pred_alphas + np.squeeze(view_angs)
        else:
            pred_alphas = new_boxes_3d[:, 6] - np.squeeze(view_angs)
        pred_centroids = np.copy(output_dict[KEY_CENTROIDS])
        if centroid_type == "middle":
            pred_centroids[:, 1:2] = pred_centroids[:, 1:2] + new_boxes_3d[:, 5:6] / 2
        new_boxes_3d[:, 0:3] = pred_centroids
        if post_process_cen_x:
            new_boxes_3d[:, 0] = np.squeeze(np.asarray(
                [postprocess_cen_x(b2, b3, cam_p) for b2, b3 in zip(new_boxes_2d, new_boxes_3d)]))
        valid_boxes_3d, valid_boxes_2d = new_boxes_3d[0:num_objs], new_boxes_2d[0:num_objs]
        new_scores = score_boxes(cam_p, img.shape, valid_boxes_2d, valid_boxes_3d, valid_scores)
        classes = output_dict[SAMPLE_LABEL_CLASS_INDICES][0:num_objs] - 1
        pred[KEY_BOX_3D] = np.hstack([valid_boxes_3d, new_scores, classes])
        pred[KEY_BOX_2D] = np.hstack([valid_boxes_2d, np.expand_dims(pred_alphas[0:num_objs], 1), new_scores, classes])
    return pred


def save_predictions(sample_name, predictions, output_dirs, output_types):
    """monopsr_model.py:1075-1102 on an already formatted prediction dict: float16 .npy of the local maps, '%0.5f'
    text files of the 3-D and 2-D detections (the per-instance mask PNGs of :1086-1091 need cv2 and are written only
    when it is importable)."""
    if KEY_INST_XYZ_MAP_LOCAL in output_types:
        d = output_dirs[OUT_DIR_XYZ_MAP_LOCAL]
        np.save(os.path.join(d, "{}.npy".format(sample_name)), predictions[KEY_INST_XYZ_MAP_LOCAL].astype(np.float16))
        if KEY_VALID_MASK_MAPS in output_types:
            try:
                import cv2
            except ImportError:
                cv2 = None
            if cv2 is not None:
                masks = predictions[KEY_VALID_MASK_MAPS].astype(np.uint8) * 255
                for i, m in enumerate(masks):
                    cv2.imwrite(os.path.join(d, "{}_{}.png".format(sample_name, i)), m)
    if KEY_CENTROIDS in output_types:
        np.savetxt(os.path.join(output_dirs[OUT_DIR_BOX_3D], "{}.txt".format(sample_name)), predictions[KEY_BOX_3D],
                   fmt="%0.5f")
        np.savetxt(os.path.join(output_dirs[OUT_DIR_BOX_2D], "{}.txt".format(sample_name)), predictions[KEY_BOX_2D],
                   fmt="%0.5f")
