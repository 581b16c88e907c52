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
    """tensors are immutable: `a += b` on a TF tensor REBINDS a -- it must not write through to an array another
    name (or a slice of another tensor) still refers to, as numpy's in-place operators would"""

    def get_shape(self):
        return _Shape(self.shape)

    def __iadd__(self, other):
        return np.add(self, other).view(T)

    def __isub__(self, other):
        return np.subtract(self, other).view(T)

    def __imul__(self, other):
        return np.multiply(self, other).view(T)

    def __itruediv__(self, other):
        return np.true_divide(self, other).view(T)


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
    tf.to_float = lambda x, **k: t(np.asarray(x, np.float64))
    tf.cast = lambda x, dtype, **k: t(np.asarray(x, np.float64 if dtype in (np.float32, np.float64) else dtype))
    tf.divide = lambda a, b, **k: t(np.asarray(a, np.float64) / b)
    tf.where = lambda c, a, b: t(np.where(c, a, b))
    tf.is_nan = lambda x: np.isnan(np.asarray(x, np.float64))
    tf.one_hot = one_hot
    # geometry ops (instance_utils.py:567-681,738-788,907-953; transform_utils.py:69-173; calib_utils.py:263-280;
    # monopsr_output_builder.py:407-438,551-571,663-746)
    f64 = lambda x: np.asarray(x, np.float64)
    tf.stack = lambda values, axis=0, **k: t(np.stack([f64(v) for v in values], axis=axis))
    tf.concat = lambda values, axis, **k: t(np.concatenate([f64(v) for v in values], axis=axis))
    tf.sin, tf.cos, tf.tan = (lambda x, **k: t(np.sin(f64(x)))), (lambda x, **k: t(np.cos(f64(x)))), (lambda x, **k: t(np.tan(f64(x))))
    tf.atan = lambda x, **k: t(np.arctan(f64(x)))
    tf.atan2 = lambda y, x, **k: t(np.arctan2(f64(y), f64(x)))
    tf.matmul = lambda a, b, **k: t(np.matmul(f64(a), f64(b)))
    tf.linspace = lambda start, stop, num, **k: t(np.linspace(float(start), float(stop), int(num)))
    tf.transpose = lambda x, perm=None, **k: t(np.transpose(f64(x), perm))
    tf.tile = lambda x, multiples, **k: t(np.tile(f64(x), multiples))
    tf.zeros = lambda shape, dtype=np.float32, **k: t(np.zeros(shape, np.float64))
    tf.less = lambda a, b, **k: f64(a) < b
    tf.round = lambda x, **k: t(np.round(f64(x)))          # both round half to even
    tf.pad = lambda x, paddings, mode="CONSTANT", constant_values=0, **k: t(np.pad(f64(x), paddings, constant_values=constant_values))
    tf.meshgrid = lambda *a, **k: [t(m) for m in np.meshgrid(*[f64(v) for v in a], indexing=k.get("indexing", "xy"))]
    tf.clip_by_value = lambda x, lo, hi, **k: t(np.clip(f64(x), lo, hi))
    tf.abs = lambda x, **k: t(np.abs(f64(x)))
    tf.maximum = lambda a, b, **k: t(np.maximum(f64(a), f64(b)))

    def reduce_sum(x, axis=None, reduction_indices=None, keep_dims=False, keepdims=False, **k):
        ax = axis if axis is not None else reduction_indices
        return t(np.sum(f64(x), axis=tuple(ax) if isinstance(ax, (list, tuple)) else ax, keepdims=keep_dims or keepdims))
    tf.reduce_sum = reduce_sum

    def map_fn(fn, elems, dtype=None, **k):
        if isinstance(elems, (list, tuple)):
            n = len(elems[0])
            results = [fn(tuple(t(e[i]) for e in elems)) for i in range(n)]
        else:
            results = [fn(t(e)) for e in elems]
        if isinstance(results[0], (list, tuple)):
            return type(results[0])(t(np.stack([f64(r[j]) for r in results])) for j in range(len(results[0])))
        return t(np.stack([f64(r) for r in results]))
    tf.map_fn = map_fn
    # ground-truth target synthesis (instance_utils.py:395-481, depth_map_utils.py:161-236)
    tf.to_int32 = lambda x, **k: t(np.asarray(x).astype(np.int32))
    tf.greater_equal = lambda a, b, **k: f64(a) >= b
    tf.reduce_max = lambda x, axis=None, keepdims=False, keep_dims=False, **k: t(np.max(f64(x), axis=axis, keepdims=keepdims or keep_dims))
    tf.stop_gradient = lambda x, **k: x

    def resize_nearest_neighbor(images, size, align_corners=False, **k):
        """TF 1.8 ResizeNearestNeighbor: in = min(round_half_away(out * scale), in_size - 1) with
        scale = (in - 1) / (out - 1) when align_corners (else floor(out * in / out))"""
        x = f64(images)
        n, ih, iw, c = x.shape
        oh, ow = int(size[0]), int(size[1])

        def index(o, i):
            j = np.arange(o, dtype=np.float32)
            if align_corners and o > 1:
                r = j * np.float32((i - 1) / (o - 1))
                r = np.where(r >= 0, np.floor(r + np.float32(0.5)), np.ceil(r - np.float32(0.5)))
            else:
                r = np.floor(j * np.float32(i / o))
            return np.minimum(r.astype(np.int64), i - 1)
        return t(x[:, index(oh, ih)][:, :, index(ow, iw)])
    tf.image = _Any("tensorflow.image")
    tf.image.resize_nearest_neighbor = resize_nearest_neighbor
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
