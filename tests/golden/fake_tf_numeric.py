"""TEST INFRASTRUCTURE: a NUMPY-backed stand-in for the TensorFlow ops the reference's loss assembly uses
(monopsr/core/models/monopsr/monopsr_model.py:554-958 `MonoPSRModel.loss`, monopsr/builders/loss_builder.py,
object_detection/core/losses.py:40-157,283-317, monopsr/core/losses_custom.py:93-132), so that this code -- unmodified --
can be EXECUTED on arrays: which tensors enter which loss, with which mask, expand_dims / reshape gymnastics, loss
weights from the yaml, divisions by num_boxes and the summation into the total are then the reference's own.
The primitives are restated from the TF 1.x documentation (they are the part that lives in TensorFlow, not in the
reference): tf.losses.huber_loss + compute_weighted_loss reductions NONE / SUM_BY_NONZERO_WEIGHTS (weights broadcast
to the losses' shape before counting), tf.nn.softmax_cross_entropy_with_logits, one_hot with on / off values."""
import contextlib
import sys
import types

import numpy as np


class T(np.ndarray):
    def get_shape(self):
        return _Shape(self.shape)


class _Shape(list):
    def as_list(self):
        return list(self)


def t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


class _Any(types.ModuleType):
    """module / namespace whose unknown attributes are inert stubs (import-time references to TF symbols that the loss
    code never executes)"""
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:      # used as a decorator
            return a[0]
        return self


def huber_loss(labels, predictions, weights=1.0, delta=1.0, scope=None, loss_collection=None, reduction="weighted_sum_by_nonzero_weights"):
    error = np.asarray(predictions, np.float64) - np.asarray(labels, np.float64)
    abs_error = np.abs(error)
    quadratic = np.minimum(abs_error, delta)
    linear = abs_error - quadratic
    losses = 0.5 * quadratic ** 2 + delta * linear
    w = np.broadcast_to(np.asarray(weights, np.float64), losses.shape)
    weighted = losses * w
    if reduction == "none":
        return t(weighted)
    if reduction == "weighted_sum_by_nonzero_weights":
        present = np.count_nonzero(w)
        return t(weighted.sum() / present if present > 0 else 0.0)
    raise NotImplementedError(reduction)


def softmax_cross_entropy_with_logits(labels=None, logits=None, **kw):
    z = np.asarray(logits, np.float64)
    z = z - z.max(axis=-1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(axis=-1, keepdims=True))
    return t(-(np.asarray(labels, np.float64) * logp).sum(axis=-1))


def one_hot(indices, depth, on_value=1.0, off_value=0.0, **kw):
    idx = np.asarray(indices).astype(np.int64)
    out = np.full(idx.shape + (depth,), off_value, np.float64)
    np.put_along_axis(out, idx[..., None], on_value, axis=-1)
    return t(out)


def install():
    tf = _Any("tensorflow")
    tf.float32, tf.int32 = np.float32, np.int32
    tf.variable_scope = lambda *a, **k: contextlib.nullcontext()
    tf.name_scope = lambda *a, **k: contextlib.nullcontext("scope")
    tf.ones = lambda shape, dtype=np.float32: t(np.ones(shape, np.float64))
    tf.ones_like = lambda x, **k: t(np.ones_like(np.asarray(x, np.float64)))
    tf.zeros_like = lambda x, **k: t(np.zeros_like(np.asarray(x, np.float64)))
    tf.expand_dims = lambda x, axis, **k: t(np.expand_dims(np.asarray(x), axis))
    tf.squeeze = lambda x, axis=None, **k: t(np.squeeze(np.asarray(x), axis=axis))
    tf.reshape = lambda x, shape, **k: t(np.reshape(np.asarray(x), shape))
    tf.shape = lambda x, **k: list(np.shape(x))
    tf.reduce_sum = lambda x, axis=None, **k: t(np.sum(np.asarray(x, np.float64), axis=tuple(axis) if isinstance(axis, list) else axis))
    tf.to_float = lambda x, **k: t(np.asarray(x, np.float64))
    tf.divide = lambda a, b, **k: t(np.asarray(a, np.float64) / b)
    tf.where = lambda c, a, b: t(np.where(c, a, b))
    tf.is_nan = lambda x: np.isnan(np.asarray(x, np.float64))
    tf.one_hot = one_hot
    tf.summary = _Any("tensorflow.summary")
    tf.summary.scalar = lambda *a, **k: None
    tf.nn = _Any("tensorflow.nn")
    tf.nn.softmax_cross_entropy_with_logits = softmax_cross_entropy_with_logits
    tf.losses = _Any("tensorflow.losses")
    tf.losses.huber_loss = huber_loss
    tf.losses.Reduction = types.SimpleNamespace(NONE="none", SUM_BY_NONZERO_WEIGHTS="weighted_sum_by_nonzero_weights")
    for name in ("tensorflow.contrib", "tensorflow.contrib.slim", "tensorflow.python", "tensorflow.python.framework",
                 "tensorflow.python.framework.ops", "tensorflow.python.ops", "tensorflow.python.ops.math_ops", "png"):
        sys.modules[name] = _Any(name)
    tf.contrib = sys.modules["tensorflow.contrib"]
    sys.modules["tensorflow"] = tf
    return tf
