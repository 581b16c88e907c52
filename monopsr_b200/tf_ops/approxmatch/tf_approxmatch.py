"""Approximate Earth Mover's Distance ops -- host-side mirror of the reference's op module.

Reference: src/tf_ops/approxmatch/tf_approxmatch.py:15-71 (``approx_match`` is
non-differentiable, ``match_cost`` has gradients w.r.t. xyz1/xyz2 only, scaled by
grad_cost[:,None,None]) and the op shells' validation
(src/tf_ops/approxmatch/tf_approxmatch.cpp:145-260).  Compute is sm_100a CUDA behind
``mpb_approxmatch`` / ``mpb_matchcost`` / ``mpb_matchcostgrad``.
"""
import torch

from ... import lib as _lib


def _check(xyz1, xyz2, op):
    if xyz1.dim() != 3 or xyz1.shape[2] != 3:
        raise ValueError("%s expects (batch_size,num_points,3) xyz1 shape" % op)
    if xyz2.dim() != 3 or xyz2.shape[2] != 3 or xyz2.shape[0] != xyz1.shape[0]:
        raise ValueError("%s expects (batch_size,num_points,3) xyz2 shape, and batch_size must match" % op)


def approx_match(xyz1, xyz2):
    '''
input:
    xyz1 : batch_size * #dataset_points * 3
    xyz2 : batch_size * #query_points * 3
returns:
    match : batch_size * #query_points * #dataset_points
    '''
    xyz1 = _lib.require_cuda(xyz1.detach(), "xyz1", torch.float32)
    xyz2 = _lib.require_cuda(xyz2.detach(), "xyz2", torch.float32)
    _check(xyz1, xyz2, "ApproxMatch")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = torch.empty((b, m, n), dtype=torch.float32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        st = _lib.load().mpb_approxmatch(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(),
                                         match.data_ptr(), None, _lib.stream_ptr())
    _lib.check(st, "mpb_approxmatch")
    return match


def _match_cost_raw(xyz1, xyz2, match):
    xyz1 = _lib.require_cuda(xyz1, "xyz1", torch.float32)
    xyz2 = _lib.require_cuda(xyz2, "xyz2", torch.float32)
    match = _lib.require_cuda(match, "match", torch.float32)
    _check(xyz1, xyz2, "MatchCost")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if tuple(match.shape) != (b, m, n):
        raise ValueError("MatchCost expects (batch_size,#query,#dataset) match shape")
    cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        st = _lib.load().mpb_matchcost(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(),
                                       match.data_ptr(), cost.data_ptr(), _lib.stream_ptr())
    _lib.check(st, "mpb_matchcost")
    return cost


def match_cost_grad(xyz1, xyz2, match):
    """``MatchCostGrad`` op (tf_approxmatch.cpp:16-21): -> grad1 (b,n,3), grad2 (b,m,3)."""
    xyz1 = _lib.require_cuda(xyz1, "xyz1", torch.float32)
    xyz2 = _lib.require_cuda(xyz2, "xyz2", torch.float32)
    match = _lib.require_cuda(match, "match", torch.float32)
    _check(xyz1, xyz2, "MatchCostGrad")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if tuple(match.shape) != (b, m, n):
        raise ValueError("MatchCost expects (batch_size,#query,#dataset) match shape")
    g1 = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    g2 = torch.empty((b, m, 3), dtype=torch.float32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        st = _lib.load().mpb_matchcostgrad(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(),
                                           match.data_ptr(), g1.data_ptr(), g2.data_ptr(),
                                           _lib.stream_ptr())
    _lib.check(st, "mpb_matchcostgrad")
    return g1, g2


class _MatchCost(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, match):
        ctx.save_for_backward(xyz1, xyz2, match)
        return _match_cost_raw(xyz1, xyz2, match)

    @staticmethod
    def backward(ctx, grad_cost):
        # tf_approxmatch.py:52-71
        xyz1, xyz2, match = ctx.saved_tensors
        g1, g2 = match_cost_grad(xyz1, xyz2, match)
        gc = grad_cost.reshape(-1, 1, 1)
        return g1 * gc, g2 * gc, None


def match_cost(xyz1, xyz2, match):
    '''
input:
    xyz1 : batch_size * #dataset_points * 3
    xyz2 : batch_size * #query_points * 3
    match : batch_size * #query_points * #dataset_points
returns:
    cost : batch_size
    '''
    return _MatchCost.apply(xyz1, xyz2, match)
