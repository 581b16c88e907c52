"""Chamfer nearest-neighbour op -- host-side mirror of the reference's op module.

Reference: src/tf_ops/nn_distance/tf_nndistance.py:15-40 (``nn_distance`` + the
``NnDistance`` gradient registration) and the op shell's shape validation
(src/tf_ops/nn_distance/tf_nndistance.cpp:51-58,94-105).  Same names, argument meaning,
output order and error behaviour; tensors are torch CUDA tensors instead of TF graph
tensors, and the backward hook is a ``torch.autograd.Function`` instead of
``ops.RegisterGradient``.  Compute is the sm_100a kernel behind ``mpb_nn_distance``.
"""
import torch

from ... import lib as _lib


def _check_clouds(xyz1, xyz2, op):
    # OP_REQUIRES checks of the reference op shell, same messages
    if xyz1.dim() != 3:
        raise ValueError("%s requires xyz1 be of shape (batch,#points,3)" % op)
    if xyz1.shape[2] != 3:
        raise ValueError("%s only accepts 3d point set xyz1" % op)
    if xyz2.dim() != 3:
        raise ValueError("%s requires xyz2 be of shape (batch,#points,3)" % op)
    if xyz2.shape[2] != 3:
        raise ValueError("%s only accepts 3d point set xyz2" % op)
    if xyz2.shape[0] != xyz1.shape[0]:
        raise ValueError("%s expects xyz1 and xyz2 have same batch size" % op)


def _nn_distance_raw(xyz1, xyz2):
    xyz1 = _lib.require_cuda(xyz1, "xyz1", torch.float32)
    xyz2 = _lib.require_cuda(xyz2, "xyz2", torch.float32)
    _check_clouds(xyz1, xyz2, "NnDistance")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = torch.empty((b, n), dtype=torch.float32, device=xyz1.device)
    idx1 = torch.empty((b, n), dtype=torch.int32, device=xyz1.device)
    dist2 = torch.empty((b, m), dtype=torch.float32, device=xyz1.device)
    idx2 = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        st = _lib.load().mpb_nn_distance(b, n, xyz1.data_ptr(), m, xyz2.data_ptr(),
                                         dist1.data_ptr(), idx1.data_ptr(),
                                         dist2.data_ptr(), idx2.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "mpb_nn_distance")
    return dist1, idx1, dist2, idx2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """``NnDistanceGrad`` op (tf_nndistance.cpp:10-18): -> grad_xyz1 (b,n,3), grad_xyz2 (b,m,3)."""
    xyz1 = _lib.require_cuda(xyz1, "xyz1", torch.float32)
    xyz2 = _lib.require_cuda(xyz2, "xyz2", torch.float32)
    _check_clouds(xyz1, xyz2, "NnDistanceGrad")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    grad_dist1 = _lib.require_cuda(grad_dist1, "grad_dist1", torch.float32)
    grad_dist2 = _lib.require_cuda(grad_dist2, "grad_dist2", torch.float32)
    idx1 = _lib.require_cuda(idx1, "idx1", torch.int32)
    idx2 = _lib.require_cuda(idx2, "idx2", torch.int32)
    if tuple(grad_dist1.shape) != (b, n):
        raise ValueError("NnDistanceGrad requires grad_dist1 be of shape(batch,#points)")
    if tuple(idx1.shape) != (b, n):
        raise ValueError("NnDistanceGrad requires idx1 be of shape(batch,#points)")
    if tuple(grad_dist2.shape) != (b, m):
        raise ValueError("NnDistanceGrad requires grad_dist2 be of shape(batch,#points)")
    if tuple(idx2.shape) != (b, m):
        raise ValueError("NnDistanceGrad requires idx2 be of shape(batch,#points)")
    g1 = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    g2 = torch.empty((b, m, 3), dtype=torch.float32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        st = _lib.load().mpb_nn_distance_grad(b, n, xyz1.data_ptr(), m, xyz2.data_ptr(),
                                              grad_dist1.data_ptr(), idx1.data_ptr(),
                                              grad_dist2.data_ptr(), idx2.data_ptr(),
                                              g1.data_ptr(), g2.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "mpb_nn_distance_grad")
    return g1, g2


class _NnDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, idx1, dist2, idx2 = _nn_distance_raw(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, idx1, dist2, idx2

    @staticmethod
    def backward(ctx, grad_dist1, grad_idx1, grad_dist2, grad_idx2):
        # tf_nndistance.py:34-40: grad_idx* are ignored
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        if grad_dist1 is None:
            grad_dist1 = torch.zeros(idx1.shape, dtype=torch.float32, device=xyz1.device)
        if grad_dist2 is None:
            grad_dist2 = torch.zeros(idx2.shape, dtype=torch.float32, device=xyz1.device)
        return nn_distance_grad(xyz1, xyz2, grad_dist1.contiguous(), idx1,
                                grad_dist2.contiguous(), idx2)


def nn_distance(xyz1, xyz2):
    '''
Computes the distance of nearest neighbors for a pair of point clouds
input: xyz1: (batch_size,#points_1,3)  the first point cloud
input: xyz2: (batch_size,#points_2,3)  the second point cloud
output: dist1: (batch_size,#point_1)   distance from first to second
output: idx1:  (batch_size,#point_1)   nearest neighbor from first to second
output: dist2: (batch_size,#point_2)   distance from second to first
output: idx2:  (batch_size,#point_2)   nearest neighbor from second to first
    '''
    return _NnDistance.apply(xyz1, xyz2)
