"""Point-set losses -- host-side mirror of src/monopsr/core/losses_custom.py:135-198.

``ChamferDistance`` and ``EarthMoversDistance`` keep the reference's ``Loss.__call__`` contract
(object_detection/core/losses.py:40-90): ``loss(prediction_tensor, target_tensor, weights=...)``
with (B,h,w,3) maps and a (B,h,w,1) valid mask.  Both clouds are multiplied by the mask (so
masked pixels become the point (0,0,0) on both sides -- quirk Q8), reshaped to (B, h*w, 3) and
handed to the sm_100a ops; gradients flow through the ops' autograd hooks.
"""
import torch

from ..tf_ops.approxmatch import tf_approxmatch
from ..tf_ops.nn_distance import tf_nndistance


class Loss(object):
    def __call__(self, prediction_tensor, target_tensor, ignore_nan_targets=False, scope=None, **params):
        if ignore_nan_targets:
            target_tensor = torch.where(torch.isnan(target_tensor), prediction_tensor, target_tensor)
        return self._compute_loss(prediction_tensor, target_tensor, **params)

    def _compute_loss(self, prediction_tensor, target_tensor, **params):
        raise NotImplementedError


def _valid_points(prediction_tensor, target_tensor, weights):
    batch_size = prediction_tensor.shape[0]
    p = (prediction_tensor * weights).reshape(batch_size, -1, 3).contiguous()
    t = (target_tensor * weights).reshape(batch_size, -1, 3).contiguous()
    return p, t, batch_size


class EarthMoversDistance(Loss):
    """losses_custom.py:135-166.  `ops` = (approx_match, match_cost) replaces the sm_100a ops (CPU tests of the assembly)."""

    def __init__(self, ops=None):
        self._approx_match, self._match_cost = ops or (tf_approxmatch.approx_match, tf_approxmatch.match_cost)

    def _compute_loss(self, prediction_tensor, target_tensor, weights):
        p, t, b = _valid_points(prediction_tensor, target_tensor, weights)
        match = self._approx_match(p, t)
        distances = self._match_cost(p, t, match)
        return distances.sum() / float(b)


class ChamferDistance(Loss):
    """losses_custom.py:169-198.  `ops` = (nn_distance,) replaces the sm_100a op (CPU tests of the assembly)."""

    def __init__(self, ops=None):
        (self._nn_distance,) = ops or (tf_nndistance.nn_distance,)

    def _compute_loss(self, prediction_tensor, target_tensor, weights):
        p, t, b = _valid_points(prediction_tensor, target_tensor, weights)
        dist1, idx1, dist2, idx2 = self._nn_distance(p, t)
        return (dist1.sum() + dist2.sum()) / float(b)


def point_set_metrics(pred_xyz_maps, gt_xyz_maps, valid_mask_maps, num_objs, ops=None):
    """metric_emd / metric_chamfer of monopsr_model.py:1112-1170: per-object distances divided by
    the object's number of valid pixels -> two (num_objs,) tensors.  `ops` = (approx_match, match_cost, nn_distance)
    replaces the sm_100a ops (the assembly around them is tested on the CPU with stand-ins)."""
    approx_match, match_cost, nn_distance = ops or (tf_approxmatch.approx_match, tf_approxmatch.match_cost,
                                                     tf_nndistance.nn_distance)
    n = pred_xyz_maps.shape[0]
    p = (pred_xyz_maps * valid_mask_maps).reshape(n, -1, 3).contiguous()
    t = (gt_xyz_maps * valid_mask_maps).reshape(n, -1, 3).contiguous()
    nvalid = valid_mask_maps[:num_objs].sum(dim=(1, 2, 3))
    match = approx_match(p, t)
    emd = match_cost(p, t, match)[:num_objs] / nvalid
    d1, _, d2, _ = nn_distance(p, t)
    chamfer = (d1.sum(1) + d2.sum(1))[:num_objs] / nvalid
    return {"metric_emd": emd, "metric_chamfer": chamfer}
