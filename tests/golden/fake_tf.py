"""TEST INFRASTRUCTURE: a recording stand-in for the small part of the TensorFlow 1.x / TF-slim graph-building API that
the reference's network builders call (object_detection/nets/resnet_v1.py, resnet_utils.py,
monopsr/core/feature_extractors/faster_rcnn_resnet_v1_feature_extractor.py, monopsr/builders/net_builder.py).

No arithmetic happens: "tensors" carry a shape only, and every layer call (conv2d, max_pool2d, pad, crop_and_resize,
resize_images, concat) is appended to RECORD with its full variable-scope name, kernel, stride, atrous rate, padding,
channel counts, activation, normaliser and its parameters -- exactly as the REFERENCE'S OWN builder code passed them
(including everything it sets through slim.arg_scope).  tests/golden/make_arch_golden.py runs the reference's
extract_features through this and stores the record; tests/test_arch_golden.py compares it with the layer tables the
sm_100a engine and the oracle are built from.  slim.arg_scope / add_arg_scope / variable_scope / repeat /
collect_named_outputs follow the documented behaviour of their TF counterparts."""
import contextlib
import functools
import math
import sys
import types

RECORD = []
_SCOPE = [""]                 # current variable-scope name
_ARG_SCOPE = [{}]             # stack of {op key: default kwargs}
_COLLECTIONS = {}


class Shape(list):
    def as_list(self):
        return list(self)

    def with_rank(self, n):
        assert len(self) == n
        return self

    def with_rank_at_least(self, n):
        assert len(self) >= n
        return self

    @property
    def ndims(self):
        return len(self)


class Tensor(object):
    """shape + a provenance tag (the placeholder or the layer scope a value derives from; kept through elementwise
    arithmetic, slicing and reshapes so that the ORDER of the operands of a concat can be recorded)"""

    def __init__(self, shape, name="", tag=None):
        self._shape, self.name, self.tag = Shape(shape), name, tag

    def get_shape(self):
        return Shape(self._shape)

    @property
    def shape(self):
        return Shape(self._shape)

    def _same(self, other):
        return Tensor(self._shape, self.name, self.tag)

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = _same

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        shp = []
        for d, i in zip(self._shape, idx):
            if isinstance(i, slice):
                shp.append(len(range(*i.indices(d))))
        shp += self._shape[len(idx):]
        return Tensor(shp, self.name, self.tag)


class VariableScope(object):
    def __init__(self, name):
        self.name = name
        self.original_name_scope = name + "/" if name else ""


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None):
    if isinstance(name_or_scope, VariableScope):          # re-entering a captured scope: absolute, not nested
        full = name_or_scope.name
    else:
        leaf = name_or_scope if name_or_scope is not None else default_name
        full = (_SCOPE[-1] + "/" + leaf) if _SCOPE[-1] else leaf
    _SCOPE.append(full)
    try:
        yield VariableScope(full)
    finally:
        _SCOPE.pop()


# ---------------------------------------------------------------------------------------------- slim.arg_scope
def _key(op):
    return getattr(op, "_key", None) or (op.__module__ + "." + op.__name__)


@contextlib.contextmanager
def arg_scope(list_ops_or_scope, **kwargs):
    if isinstance(list_ops_or_scope, dict):               # reuse of a captured scope
        assert not kwargs
        new = {k: dict(v) for k, v in list_ops_or_scope.items()}
    else:
        new = {k: dict(v) for k, v in _ARG_SCOPE[-1].items()}
        for op in list_ops_or_scope:
            assert hasattr(op, "_key"), "op was not decorated with add_arg_scope: %r" % (op,)
            new.setdefault(_key(op), {}).update(kwargs)
    _ARG_SCOPE.append(new)
    try:
        yield {k: dict(v) for k, v in new.items()}
    finally:
        _ARG_SCOPE.pop()


def add_arg_scope(func):
    key = func.__module__ + "." + func.__name__

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        merged = dict(_ARG_SCOPE[-1].get(key, {}))
        merged.update(kwargs)
        return func(*args, **merged)
    wrapper._key = key
    return wrapper


# ---------------------------------------------------------------------------------------------- layers
def _pair(v):
    return [v, v] if isinstance(v, int) else list(v)


def _out_len(n, k, stride, rate, padding):
    keff = k + (k - 1) * (rate - 1)
    if padding == "SAME":
        return int(math.ceil(n / float(stride)))
    assert padding == "VALID"
    return (n - keff) // stride + 1


def relu(x, name=None):
    return x


def relu6(x, name=None):
    return x


relu.__name__, relu6.__name__ = "relu", "relu6"


@add_arg_scope
def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, is_training=True, updates_collections=None,
               fused=None, scope=None, **kw):
    return inputs


_BN_DEFAULTS = dict(decay=0.999, center=True, scale=False, epsilon=0.001, is_training=True)


@add_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", rate=1, activation_fn=relu, normalizer_fn=None,
           normalizer_params=None, weights_initializer=None, weights_regularizer=None, biases_initializer="zeros",
           outputs_collections=None, scope=None, **kw):
    kh, kw_ = _pair(kernel_size)
    n, h, w, c = inputs.get_shape().as_list()
    with variable_scope(scope, "Conv", [inputs]) as sc:
        out = Tensor([n, _out_len(h, kh, stride, rate, padding), _out_len(w, kw_, stride, rate, padding), num_outputs])
        bn = None
        if normalizer_fn is not None:
            assert normalizer_fn is batch_norm
            bn = dict(_BN_DEFAULTS)
            bn.update({k: v for k, v in _ARG_SCOPE[-1].get(_key(batch_norm), {}).items() if k in _BN_DEFAULTS})
            bn.update({k: v for k, v in (normalizer_params or {}).items() if k in _BN_DEFAULTS})
        RECORD.append({"op": "conv2d", "scope": sc.name, "kernel": [kh, kw_], "stride": stride, "rate": rate,
                       "padding": padding, "cin": c, "cout": num_outputs,
                       "activation": getattr(activation_fn, "__name__", None) if activation_fn else None,
                       "bias": normalizer_fn is None and biases_initializer is not None, "batch_norm": bn,
                       "in_shape": [n, h, w, c], "out_shape": list(out.get_shape())})
        return collect_named_outputs(outputs_collections, sc.name, out)


@add_arg_scope
def max_pool2d(inputs, kernel_size, stride=2, padding="VALID", outputs_collections=None, scope=None):
    kh, kw_ = _pair(kernel_size)
    n, h, w, c = inputs.get_shape().as_list()
    with variable_scope(scope, "MaxPool2D", [inputs]) as sc:
        out = Tensor([n, _out_len(h, kh, stride, 1, padding), _out_len(w, kw_, stride, 1, padding), c])
        RECORD.append({"op": "max_pool2d", "scope": sc.name, "kernel": [kh, kw_], "stride": stride, "padding": padding,
                       "in_shape": [n, h, w, c], "out_shape": list(out.get_shape())})
        return collect_named_outputs(outputs_collections, sc.name, out)


def repeat(inputs, repetitions, layer, *args, **kwargs):
    scope = kwargs.pop("scope", None)
    with variable_scope(scope, "Repeat", [inputs]):
        scope = scope or getattr(layer, "__name__", "repeat")
        outputs = inputs
        for i in range(repetitions):
            kwargs["scope"] = scope + "_" + str(i + 1)
            outputs = layer(outputs, *args, **kwargs)
        return outputs


def collect_named_outputs(collections, alias, outputs):
    if collections:
        _COLLECTIONS.setdefault(collections, {})[alias] = outputs
    return outputs


def convert_collection_to_dict(collection, clear_collection=False):
    return dict(_COLLECTIONS.get(collection, {}))


def last_dimension(shape, min_rank=1):
    assert len(shape) >= min_rank
    return shape[-1]


def pad(tensor, paddings, **kw):
    shp = [d + p[0] + p[1] for d, p in zip(tensor.get_shape().as_list(), paddings)]
    RECORD.append({"op": "pad", "scope": _SCOPE[-1], "paddings": [list(p) for p in paddings],
                   "in_shape": tensor.get_shape().as_list(), "out_shape": shp})
    return Tensor(shp)


def crop_and_resize(image, boxes, box_ind, crop_size, **kw):
    n = box_ind.get_shape().as_list()[0]
    shp = [n, int(crop_size[0]), int(crop_size[1]), image.get_shape().as_list()[3]]
    RECORD.append({"op": "crop_and_resize", "scope": _SCOPE[-1], "crop_size": [int(crop_size[0]), int(crop_size[1])],
                   "in_shape": image.get_shape().as_list(), "out_shape": shp})
    return Tensor(shp)


def resize_images(images, size, align_corners=False, **kw):
    n, _, _, c = images.get_shape().as_list()
    shp = [n, int(size[0]), int(size[1]), c]
    RECORD.append({"op": "resize_images", "scope": _SCOPE[-1], "size": [int(size[0]), int(size[1])],
                   "align_corners": bool(align_corners), "in_shape": images.get_shape().as_list(), "out_shape": shp})
    return Tensor(shp)


def concat(values, axis, **kw):
    shp = values[0].get_shape().as_list()
    shp[axis] = sum(v.get_shape().as_list()[axis] for v in values)
    RECORD.append({"op": "concat", "scope": _SCOPE[-1], "axis": axis,
                   "in_shapes": [v.get_shape().as_list() for v in values], "in_tags": [v.tag for v in values],
                   "out_shape": shp})
    return Tensor(shp)


def flatten(inputs, outputs_collections=None, scope=None):
    shp = inputs.get_shape().as_list()
    n = 1
    for d in shp[1:]:
        n *= d
    RECORD.append({"op": "flatten", "scope": _SCOPE[-1], "in_shape": shp, "out_shape": [shp[0], n], "in_tag": inputs.tag})
    return Tensor([shp[0], n], tag="flatten(%s)" % inputs.tag)


@add_arg_scope
def fully_connected(inputs, num_outputs, activation_fn=relu, normalizer_fn=None, normalizer_params=None,
                    biases_initializer="zeros", outputs_collections=None, scope=None, **kw):
    n, c = inputs.get_shape().as_list()
    with variable_scope(scope, "fully_connected", [inputs]) as sc:
        RECORD.append({"op": "fully_connected", "scope": sc.name, "cin": c, "cout": num_outputs,
                       "activation": getattr(activation_fn, "__name__", None) if activation_fn else None,
                       "bias": normalizer_fn is None and biases_initializer is not None, "in_tag": inputs.tag})
        return Tensor([n, num_outputs], tag=sc.name)


@add_arg_scope
def dropout(inputs, keep_prob=0.5, is_training=True, scope=None, **kw):
    RECORD.append({"op": "dropout", "scope": (_SCOPE[-1] + "/" + scope) if scope else _SCOPE[-1], "keep_prob": keep_prob,
                   "is_training": is_training})
    return inputs


def expand_dims(t, axis, **kw):
    shp = t.get_shape().as_list()
    shp.insert(axis if axis >= 0 else len(shp) + 1 + axis, 1)
    return Tensor(shp, tag=t.tag)


def squeeze(t, axis=None, **kw):
    return Tensor([d for d in t.get_shape().as_list() if d != 1], tag=t.tag)


def one_hot(indices, depth, **kw):
    return Tensor(indices.get_shape().as_list() + [depth], tag="one_hot(%s)" % indices.tag)


def reshape(t, shape, **kw):
    total = 1
    for d in t.get_shape().as_list():
        total *= d
    known = 1
    for d in shape:
        if d != -1:
            known *= d
    return Tensor([total // known if d == -1 else d for d in shape], tag=t.tag)


def tile(t, multiples, **kw):
    return Tensor([d * m for d, m in zip(t.get_shape().as_list(), multiples)], tag=t.tag)


def install():
    """put the fake modules into sys.modules (tensorflow, tensorflow.contrib, tensorflow.contrib.slim)"""
    tf = types.ModuleType("tensorflow")
    contrib = types.ModuleType("tensorflow.contrib")
    slim = types.ModuleType("tensorflow.contrib.slim")
    slim.arg_scope, slim.add_arg_scope = arg_scope, add_arg_scope
    slim.conv2d, slim.max_pool2d, slim.batch_norm, slim.repeat = conv2d, max_pool2d, batch_norm, repeat
    slim.flatten, slim.fully_connected, slim.dropout = flatten, fully_connected, dropout
    tf.expand_dims, tf.squeeze, tf.one_hot, tf.reshape, tf.tile = expand_dims, squeeze, one_hot, reshape, tile
    slim.utils = types.SimpleNamespace(collect_named_outputs=collect_named_outputs,
                                       convert_collection_to_dict=convert_collection_to_dict,
                                       last_dimension=last_dimension)
    slim.l2_regularizer = lambda scale, scope=None: ("l2", scale)
    slim.variance_scaling_initializer = lambda *a, **k: "variance_scaling"
    tf.contrib, contrib.slim = contrib, slim
    tf.nn = types.SimpleNamespace(relu=relu, relu6=relu6)
    tf.variable_scope, tf.pad, tf.concat = variable_scope, pad, concat
    tf.image = types.SimpleNamespace(crop_and_resize=crop_and_resize, resize_images=resize_images)
    tf.GraphKeys = types.SimpleNamespace(UPDATE_OPS="update_ops")
    tf.AUTO_REUSE, tf.int32, tf.float32 = "auto_reuse", "int32", "float32"
    tf.zeros = lambda shape, dtype=None, name=None: Tensor([shape] if isinstance(shape, int) else list(shape))
    tf.shape = lambda t: t.get_shape().as_list()
    tf.greater_equal = lambda a, b: a >= b
    tf.logical_and = lambda a, b: a and b
    tf.Assert = lambda cond, data: cond
    tf.control_dependencies = lambda deps: contextlib.nullcontext()
    tf.logging = types.SimpleNamespace(set_verbosity=lambda *a: None, ERROR=0)

    # optimizer construction (optimizer_builder.py): every call is recorded with the arguments the reference passes;
    # what it does NOT pass stays at the TensorFlow default, noted here from the TF 1.x signatures
    def optimizer(kind, defaults):
        def make(learning_rate, **kwargs):
            cfg = dict(defaults)
            cfg.update(kwargs)
            RECORD.append({"op": kind, "learning_rate": learning_rate, "passed": sorted(kwargs), "config": cfg})
            return {"optimizer": kind, "learning_rate": learning_rate, "config": cfg}
        return make

    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
        r = {"op": "exponential_decay", "learning_rate": learning_rate, "decay_steps": decay_steps, "decay_rate": decay_rate,
             "staircase": staircase}
        RECORD.append(r)
        return r

    def moving_average_optimizer(opt, average_decay=0.9999, num_updates=None, sequential_update=True):
        RECORD.append({"op": "MovingAverageOptimizer", "average_decay": average_decay, "num_updates": num_updates,
                       "wraps": opt["optimizer"]})
        return {"optimizer": "MovingAverageOptimizer", "inner": opt}
    tf.train = types.SimpleNamespace(
        AdamOptimizer=optimizer("AdamOptimizer", {"beta1": 0.9, "beta2": 0.999, "epsilon": 1e-08}),
        RMSPropOptimizer=optimizer("RMSPropOptimizer", {}), MomentumOptimizer=optimizer("MomentumOptimizer", {}),
        GradientDescentOptimizer=optimizer("GradientDescentOptimizer", {}), exponential_decay=exponential_decay)
    contrib.opt = types.SimpleNamespace(MovingAverageOptimizer=moving_average_optimizer)
    tf.summary = types.SimpleNamespace(scalar=lambda name, tensor, **k: name)
    sys.modules["tensorflow"], sys.modules["tensorflow.contrib"], sys.modules["tensorflow.contrib.slim"] = tf, contrib, slim
    sys.modules.setdefault("png", types.ModuleType("png"))           # pypng: imported by depth_map_utils, never called here
    return tf
