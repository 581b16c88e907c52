"""Ground-truth target synthesis on the GPU.

Host mirror of the reference's per-box TF sub-graphs (monopsr_model.py:165-203 calling
instance_utils.tf_instance_xyz_crop_from_depth_map, instance_utils.py:395-481): from the raw training inputs -- depth
map, instance masks, 2-D / 3-D boxes, estimated viewing angles, camera matrix -- to the three target tensors the
losses consume (`gt_inst_xyz_maps_local`, `gt_inst_xyz_maps_global`, `gt_valid_mask_maps`).  One CUDA launch
(csrc/targets.cu); there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from .. import lib as _lib


def _dev(x, dev, dtype):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x))
    return t.to(device=dev, dtype=dtype).contiguous()


def gt_maps_from_depth(depth_map, instance_masks, boxes_2d, boxes_3d, view_angs, cam_p, device, roi=48,
                       centroid_type="middle", rotate_view=True):
    """-> dict(gt_inst_xyz_maps_local (N,roi,roi,3), gt_inst_xyz_maps_global (N,roi,roi,3),
    gt_valid_mask_maps (N,roi,roi,1)) as CUDA tensors; argument meaning as in the reference function."""
    if centroid_type not in ("bottom", "middle"):
        raise ValueError("Invalid centroid_type %r" % (centroid_type,))
    dev = torch.device(device)
    L = _lib.load()
    d = _dev(depth_map, dev, torch.float32)
    m = _dev(instance_masks, dev, torch.uint8)
    b2, b3 = _dev(boxes_2d, dev, torch.float32), _dev(boxes_3d, dev, torch.float32)
    va, P = _dev(view_angs, dev, torch.float32), _dev(cam_p, dev, torch.float32)
    if d.dim() != 2 or m.dim() != 3 or tuple(m.shape[1:]) != tuple(d.shape) or b2.shape != (m.shape[0], 4) or \
            b3.dim() != 2 or b3.shape[0] != m.shape[0] or b3.shape[1] < 6 or va.shape != (m.shape[0],) or P.shape != (3, 4):
        raise ValueError("gt_maps_from_depth: inconsistent shapes")
    n, (H, W) = m.shape[0], d.shape
    loc = torch.empty(n, roi, roi, 3, device=dev)
    glo = torch.empty(n, roi, roi, 3, device=dev)
    val = torch.empty(n, roi, roi, 1, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = L.mpb_gt_xyz_from_depth(n, H, W, roi, p(d), p(m), p(b2), p(b3), b3.shape[1], p(va), p(P),
                                 1 if centroid_type == "middle" else 0, 1 if rotate_view else 0, p(loc), p(glo), p(val),
                                 ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(st, "mpb_gt_xyz_from_depth")
    return {"gt_inst_xyz_maps_local": loc, "gt_inst_xyz_maps_global": glo, "gt_valid_mask_maps": val}


KITTI_CHANNEL_MEANS = (92.8403, 97.7996, 93.5843)        # core/img_preprocessor.py:7
IMAGENET_CHANNEL_MEANS = (123.68, 116.78, 103.94)        # core/img_preprocessor.py:10


def image_inputs(rgb_image, boxes_2d_norm, device, image_input_shape=(320, 1216), img_roi_size=48,
                 resized_full_img_shape=(160, 608), mean_sub_type="kitti"):
    """Raw camera image (H,W,3 uint8 or float) -> dict(rgb_crops (N,roi,roi,3), full_img (1,FH,FW,3)) on the GPU:
    ImgPreprocessor.preprocess_input + tf.image.crop_and_resize + resize_bilinear(align_corners=True)
    (core/img_preprocessor.py:12-35, monopsr_model.py:128-133,222-233; yaml keys of the same names)."""
    if mean_sub_type == "kitti":
        means = KITTI_CHANNEL_MEANS
    elif mean_sub_type == "imagenet":
        means = IMAGENET_CHANNEL_MEANS
    else:
        raise ValueError("Invalid mean subtraction type {}".format(mean_sub_type))
    dev = torch.device(device)
    L = _lib.load()
    img = rgb_image if isinstance(rgb_image, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(rgb_image))
    if img.dim() != 3 or img.shape[2] != 3:
        raise ValueError("rgb_image must be (H, W, 3)")
    is_u8 = img.dtype == torch.uint8
    img = img.to(dev).contiguous() if is_u8 else img.to(device=dev, dtype=torch.float32).contiguous()
    boxes = _dev(boxes_2d_norm, dev, torch.float32)
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes_2d_norm must be (N, 4)")
    H, W = img.shape[:2]
    PH, PW = image_input_shape
    FH, FW = resized_full_img_shape
    n = boxes.shape[0]
    pre = torch.empty(PH, PW, 3, device=dev)
    crops = torch.empty(n, img_roi_size, img_roi_size, 3, device=dev)
    full = torch.empty(1, FH, FW, 3, device=dev)
    cm = (ctypes.c_float * 3)(*means)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = L.mpb_image_inputs(H, W, p(img), 1 if is_u8 else 0, ctypes.cast(cm, ctypes.c_void_p), PH, PW, p(pre), n, p(boxes),
                            img_roi_size, p(crops), FH, FW, p(full),
                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(st, "mpb_image_inputs")
    return {"rgb_crops": crops, "full_img": full, "img_preprocessed": pre}
