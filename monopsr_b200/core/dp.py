"""Data parallelism of the training step: the single collective of the hot path.

One process per GPU, one sample (num_boxes crops + one image) per rank, no activation exchange
(decoder batch-norm statistics stay per sample, as in the reference).  After the backward pass
every rank holds the gradient of its own sample in one flat fp32 arena; the arenas are summed
with ONE all-reduce over NVLink/NVSwitch (NCCL) and the mean is taken inside the fused train-op
(``grad_scale = 1/world``), BEFORE the per-variable clip -- i.e. DP-N equals the reference's
gradient averaged over N samples.
"""
import torch
import torch.distributed as dist


def allreduce_flat(flat, group=None):
    """sum `flat` (a 1-D tensor) over the process group in place; returns the scale that turns the
    sum into the mean.  Works with nccl (CUDA tensors) and gloo (CPU tensors, used by the tests)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def shard_samples(num_samples, rank, world):
    """weak scaling: sample indices handled by `rank` (round-robin over the global batch)"""
    return list(range(rank, num_samples, world))
