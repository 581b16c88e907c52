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
