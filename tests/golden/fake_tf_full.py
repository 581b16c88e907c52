"""TEST INFRASTRUCTURE: the numpy-backed TF stand-in (fake_tf_numeric.py) extended with NUMERIC TF-slim layers, so that
the reference's whole graph -- MonoPSRModel.__init__ / build / loss with net_builder, the ResNet builders and the
output builder, all unmodified -- can be executed end to end on arrays and compared with oracle.network.forward / loss.

slim.arg_scope / add_arg_scope / variable_scope / repeat / collect_named_outputs are the implementations of fake_tf.py
(real scoping semantics: layer parameters are looked up by the variable-scope name the reference's code produces).
The TF KERNELS are supplied by the oracle's own primitive functions (convolution, batch norm, pooling, bilinear resize,
crop_and_resize): this run pins the WIRING of the graph -- what feeds what, in which order, with which arguments --
not the arithmetic of those kernels (cross-checked elsewhere against torchvision / torch.nn.functional)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as TF

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fake_tf as R            # noqa: E402  scoping machinery
import fake_tf_numeric as N    # noqa: E402  numeric ops

PARAMS = {}                    # variable name -> numpy array (TF layouts), set by the caller
FEEDS = {}                     # placeholder name -> list of arrays, consumed in creation order
CREATED = []                   # variable names the graph asked for (block4 included)
t = N.t


def _param(name, shape, rng=np.random.RandomState(1234)):
    CREATED.append(name)
    if name not in PARAMS:
        assert "/block4/" in name, "graph asks for a variable the parameter table does not have: " + name
        PARAMS[name] = rng.standard_normal(shape) * 0.01 if name.endswith("weights") else \
            (np.ones(shape) if name.endswith(("gamma", "moving_variance")) else np.zeros(shape))
    p = np.asarray(PARAMS[name], np.float64)
    assert tuple(p.shape) == tuple(shape), (name, p.shape, shape)
    return torch.from_numpy(np.ascontiguousarray(p))


def _nchw(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float64))).permute(0, 3, 1, 2)


def _out(y):
    return t(y.permute(0, 2, 3, 1).contiguous().numpy())


def _same_pad(n, k, s, rate=1):
    keff = k + (k - 1) * (rate - 1)
    total = max((int(np.ceil(n / s)) - 1) * s + keff - n, 0)
    return total // 2, total - total // 2


def relu(x, name=None):
    return t(np.maximum(np.asarray(x, np.float64), 0.0))


@R.add_arg_scope
def batch_norm(inputs, **kw):
    raise AssertionError("batch_norm is only used as a normalizer_fn here")


@R.add_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", rate=1, activation_fn=relu, normalizer_fn=None,
           normalizer_params=None, weights_initializer=None, weights_regularizer=None, biases_initializer="zeros",
           outputs_collections=None, scope=None, **kw):
    kh, kw_ = R._pair(kernel_size)
    x = _nchw(inputs)
    cin = x.shape[1]
    with R.variable_scope(scope, "Conv", [inputs]) as sc:
        w = _param(sc.name + "/weights", (kh, kw_, cin, num_outputs)).permute(3, 2, 0, 1)
        if padding == "SAME":
            pt, pb = _same_pad(x.shape[2], kh, stride, rate)
            pl, pr = _same_pad(x.shape[3], kw_, stride, rate)
            x = TF.pad(x, (pl, pr, pt, pb))
        y = TF.conv2d(x, w, stride=stride, dilation=rate)
        if normalizer_fn is not None:
            bn = dict(decay=0.999, center=True, scale=False, epsilon=0.001, is_training=True)
            bn.update({k: v for k, v in R._ARG_SCOPE[-1].get(R._key(batch_norm), {}).items() if k in bn})
            bn.update({k: v for k, v in (normalizer_params or {}).items() if k in bn})
            b = sc.name + "/BatchNorm/"
            c = num_outputs
            if bn["is_training"]:
                mean, var = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=False)
                _param(b + "moving_mean", (c,)), _param(b + "moving_variance", (c,))
            else:
                mean, var = _param(b + "moving_mean", (c,)), _param(b + "moving_variance", (c,))
            y = (y - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + bn["epsilon"])
            if bn["scale"]:
                y = y * _param(b + "gamma", (c,))[None, :, None, None]
            y = y + _param(b + "beta", (c,))[None, :, None, None]
        elif biases_initializer is not None:
            y = y + _param(sc.name + "/biases", (num_outputs,))[None, :, None, None]
        if activation_fn is not None:
            assert getattr(activation_fn, "__name__", "") == "relu"
            y = torch.relu(y)
        return R.collect_named_outputs(outputs_collections, sc.name, _out(y))


@R.add_arg_scope
def max_pool2d(inputs, kernel_size, stride=2, padding="VALID", outputs_collections=None, scope=None):
    kh, kw_ = R._pair(kernel_size)
    x = _nchw(inputs)
    with R.variable_scope(scope, "MaxPool2D", [inputs]) as sc:
        if padding == "SAME":
            pt, pb = _same_pad(x.shape[2], kh, stride)
            pl, pr = _same_pad(x.shape[3], kw_, stride)
            x = TF.pad(x, (pl, pr, pt, pb), value=float("-inf"))
        return R.collect_named_outputs(outputs_collections, sc.name, _out(TF.max_pool2d(x, (kh, kw_), stride)))


@R.add_arg_scope
def fully_connected(inputs, num_outputs, activation_fn=relu, normalizer_fn=None, biases_initializer="zeros",
                    outputs_collections=None, scope=None, **kw):
    x = torch.from_numpy(np.ascontiguousarray(np.asarray(inputs, np.float64)))
    with R.variable_scope(scope, "fully_connected", [inputs]) as sc:
        y = x @ _param(sc.name + "/weights", (x.shape[1], num_outputs)) + _param(sc.name + "/biases", (num_outputs,))
        if activation_fn is not None:
            assert getattr(activation_fn, "__name__", "") == "relu"
            y = torch.relu(y)
        return t(y.numpy())


def flatten(inputs, outputs_collections=None, scope=None):
    a = np.asarray(inputs, np.float64)
    return t(a.reshape(a.shape[0], -1))


@R.add_arg_scope
def dropout(inputs, keep_prob=0.5, is_training=True, scope=None, **kw):
    assert keep_prob == 1.0
    return inputs


def crop_and_resize(image, boxes, box_ind, crop_size, **kw):
    from oracle import network as onet
    assert not np.asarray(box_ind).any()
    img = torch.from_numpy(np.ascontiguousarray(np.asarray(image, np.float64)))
    return t(onet.crop_and_resize(img, torch.from_numpy(np.asarray(boxes, np.float64)), int(crop_size[0]), int(crop_size[1])).numpy())


def resize_bilinear(images, size, align_corners=False, **kw):
    assert align_corners
    return _out(TF.interpolate(_nchw(images), size=(int(size[0]), int(size[1])), mode="bilinear", align_corners=True))


def resize_images(images, size, align_corners=False, **kw):
    if align_corners:
        return resize_bilinear(images, size, True)
    # TF 1.x legacy mapping (no half-pixel centres): src = dst * in / out
    a = np.asarray(images, np.float64)
    n, h, w, c = a.shape
    oh, ow = int(size[0]), int(size[1])
    sy, sx = np.arange(oh) * (h / oh), np.arange(ow) * (w / ow)
    y0, x0 = np.floor(sy).astype(int), np.floor(sx).astype(int)
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    ly, lx = (sy - y0)[None, :, None, None], (sx - x0)[None, None, :, None]
    top = a[:, y0][:, :, x0] + (a[:, y0][:, :, x1] - a[:, y0][:, :, x0]) * lx
    bot = a[:, y1][:, :, x0] + (a[:, y1][:, :, x1] - a[:, y1][:, :, x0]) * lx
    return t(top + (bot - top) * ly)


def placeholder(dtype, shape=None, name=None):
    return t(FEEDS[name].pop(0))


def one_hot(indices, depth, on_value=1.0, off_value=0.0, **kw):
    idx = np.asarray(indices).astype(np.int64)
    out = np.full(idx.shape + (depth,), off_value, np.float64)
    ok = (idx >= 0) & (idx < depth)                     # out-of-range indices give an all-"off" row, as in TF
    np.put_along_axis(out, np.where(ok, idx, 0)[..., None], np.where(ok, on_value, out[..., 0])[..., None], axis=-1)
    return t(out)


def install():
    tf = N.install()
    slim = sys.modules["tensorflow.contrib.slim"]
    for mod in (slim,):
        mod.arg_scope, mod.add_arg_scope, mod.repeat = R.arg_scope, R.add_arg_scope, R.repeat
        mod.conv2d, mod.max_pool2d, mod.batch_norm, mod.fully_connected = conv2d, max_pool2d, batch_norm, fully_connected
        mod.flatten, mod.dropout = flatten, dropout
        mod.utils = type("U", (), dict(collect_named_outputs=staticmethod(R.collect_named_outputs),
                                       convert_collection_to_dict=staticmethod(R.convert_collection_to_dict),
                                       last_dimension=staticmethod(R.last_dimension)))
        mod.l2_regularizer = lambda *a, **k: None
        mod.variance_scaling_initializer = lambda *a, **k: None
    tf.contrib = sys.modules["tensorflow.contrib"]
    tf.contrib.slim = slim
    tf.variable_scope = R.variable_scope
    tf.nn.relu = relu
    tf.placeholder = placeholder
    tf.one_hot = one_hot
    tf.string = str
    tf.image.crop_and_resize, tf.image.resize_bilinear, tf.image.resize_images = crop_and_resize, resize_bilinear, resize_images
    tf.GraphKeys = type("G", (), dict(UPDATE_OPS="update_ops"))
    tf.shape = lambda x, **k: list(np.shape(x))
    tf.greater_equal = lambda a, b, **k: np.asarray(a) >= b
    tf.logical_and = lambda a, b, **k: np.logical_and(a, b)
    tf.Assert = lambda cond, data, **k: cond
    import contextlib
    tf.control_dependencies = lambda deps: contextlib.nullcontext()
    return tf
