"""Loss registry -- host-side mirror of src/monopsr/builders/loss_builder.py:19-84.

Same strings as the reference.  ``chamfer_dist`` and ``emd`` are callable losses backed by the
sm_100a point-set ops.  The per-box / per-pixel losses of ``monopsr_model_000`` (smooth_l1,
smooth_l1_nonzero, softmax) are not separate graph nodes here: they are fused with their
gradients into the heads kernel (csrc/heads.cu) of the engine, so the registry returns a
descriptor for them that the engine's config validation understands.
"""
from ..core import losses_custom

FUSED = ("smooth_l1", "smooth_l1_nonzero", "softmax")
KNOWN = ("berHu", "chamfer_dist", "emd", "smooth_l1", "smooth_l1_nonzero", "softmax", "focal", "softmax_temp",
         "sigmoid_ce")


class FusedLoss(object):
    """A loss that the engine evaluates inside mpb_heads_final together with its gradient."""

    def __init__(self, loss_type):
        self.loss_type = loss_type

    def __call__(self, *a, **k):
        raise NotImplementedError(
            "loss type %r is fused into the engine step (monopsr_b200.core.engine.Engine); it is not a "
            "stand-alone op in the B200 build" % self.loss_type)


def build_loss(loss_type):
    if loss_type == "chamfer_dist":
        return losses_custom.ChamferDistance()
    if loss_type == "emd":
        return losses_custom.EarthMoversDistance()
    if loss_type in FUSED:
        return FusedLoss(loss_type)
    if loss_type in KNOWN:
        raise NotImplementedError("loss type %r is not used by monopsr_model_000 and has no sm_100a kernel" % loss_type)
    raise ValueError("Invalid loss type", loss_type)


def get_loss_type_and_weight(loss_config, output_type):
    entry = getattr(loss_config, output_type, None)
    if entry is None:
        return None, None
    return entry[0], entry[1]


def add_loss_tensor(loss_config, output_type, pred_tensor, gt_tensor, mask):
    """loss_builder.py:60-84 for the stand-alone (point-set) loss types."""
    loss_type, loss_weight = get_loss_type_and_weight(loss_config, output_type)
    if loss_type is None:
        return pred_tensor.new_zeros(())
    return build_loss(loss_type)(pred_tensor, gt_tensor, weights=mask) * loss_weight
