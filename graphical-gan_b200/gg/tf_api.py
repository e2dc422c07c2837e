"""The ~45 `tf.*` symbols the reference's training scripts and tflib use (enumerated by grep over
/root/reference/*_inference_*.py and tflib/), implemented over the gg graph IR.

`import tensorflow as tf` resolves to this module through the shim package graphical-gan_b200/tensorflow/.
Semantics follow TensorFlow 1.x graph mode: everything here *describes* computation; `Session.run` executes it.
"""
import contextlib
import itertools

import numpy as np

from . import graph as G
from . import ops as O
from .graph import Tensor, Operation, float32, int32, int64

float64 = float32   # the hot path is fp32; a float64 request is served in fp32

_opt_ids = itertools.count()
_trainable = []     # param nodes created through tf.Variable(trainable=True)


# ---- graph construction ---------------------------------------------------------------------
def Variable(initial_value, name=None, trainable=True, dtype=None, **_):
    v = O.variable(initial_value, name=name, trainable=trainable)
    if trainable:
        _trainable.append(v)
    return v


def trainable_variables():
    return list(_trainable)


def global_variables_initializer():
    return Operation("noop")


def constant(value, dtype=None, shape=None, name=None):
    return O.constant(value, dtype, shape, name)


def placeholder(dtype, shape=None, name=None):
    return O.placeholder(dtype, shape, name)


@contextlib.contextmanager
def name_scope(name):
    yield name


def shape(x):
    """static shapes only: returns a list of ints (the scripts use it as a shape argument or index it)"""
    return list(x.shape)


def reshape(x, shape, name=None):
    return O.reshape(O.to_tensor(x), list(shape))


def transpose(x, perm=None, name=None):
    if perm is None:
        perm = list(reversed(range(len(x.shape))))
    return O.transpose(x, perm)


def concat(values, axis, name=None):
    if isinstance(values, int):   # tf < 1.0 argument order
        values, axis = axis, values
    return O.concat(list(values), axis)


def stack(values, axis=0, name=None):
    if all(isinstance(v, (int, np.integer)) for v in values):
        return [int(v) for v in values]          # shape vectors stay Python lists
    return O.concat([O.expand_dims(v, axis) for v in values], axis)


pack = stack


def unstack(value, axis=0):
    if isinstance(value, (list, tuple)):
        return list(value)
    return [O.getitem(value, tuple([slice(None)] * axis + [i])) for i in range(value.shape[axis])]


unpack = unstack


def expand_dims(x, axis=None, dim=None):
    return O.expand_dims(x, axis if axis is not None else dim)


def tile(x, multiples):
    return O.tile(x, multiples)


def cast(x, dtype):
    return O.cast(x, dtype)


def identity(x, name=None):
    return x


def stop_gradient(x):
    return O.stop_gradient(x)


def one_hot(indices, depth, **_):
    return O.one_hot(indices, depth)


def argmax(x, axis=None, dimension=None):
    ax = axis if axis is not None else (dimension if dimension is not None else 0)
    return O.argmax(x, ax)


def zeros_like(x):
    return O.zeros_like(x)


def ones_like(x):
    return O.ones_like(x)


# ---- arithmetic -------------------------------------------------------------------------------
def add(a, b):
    return O.add(a, b) if isinstance(a, Tensor) else O.add(O.to_tensor(a), b)


def subtract(a, b):
    return O.sub(a, b)


def multiply(a, b):
    return O.mul(a, b) if isinstance(a, Tensor) or isinstance(b, Tensor) else O.to_tensor(a * b)


def divide(a, b):
    return O.div(a, b)


def maximum(a, b):
    return O.maximum(a, b)


def minimum(a, b):
    return O.binary("min", O.to_tensor(a), O.to_tensor(b))


def square(x):
    return O.unary("square", x)


def sqrt(x):
    return O.unary("sqrt", x)


def exp(x):
    return O.unary("exp", x)


def log(x):
    return O.unary("log", x)


def tanh(x):
    return O.unary("tanh", x)


def sigmoid(x):
    return O.unary("sigmoid", x)


def abs(x):   # noqa: A001  (mirrors tf.abs)
    return O.unary("abs", x)


def negative(x):
    return O.unary("neg", x)


def pow(x, p):   # noqa: A001
    return O.pow_(x, p)


def clip_by_value(x, lo, hi):
    return O.unary("clip", x, float(lo), float(hi))


def scalar_mul(scalar, x):
    return O.mul(x, float(scalar))


def ones(shape, dtype=None, name=None):
    return O.constant(np.ones([int(v) for v in shape], np.float32))


def zeros(shape, dtype=None, name=None):
    return O.constant(np.zeros([int(v) for v in shape], np.float32))


def diag_part(x):
    """diagonal of a square matrix (tflib/objs/mmd.py:28-29): a masked row sum — no gather kernel needed"""
    n = x.shape[0]
    if len(x.shape) != 2 or x.shape[1] != n:
        raise ValueError("diag_part expects a square matrix, got %s" % (tuple(x.shape),))
    return O.reduce("sum", O.mul(x, O.constant(np.eye(n, dtype=np.float32))), [1], False)


def trace(x):
    return O.reduce("sum", diag_part(x), [0], False)


def matmul(a, b, transpose_a=False, transpose_b=False):
    return O.matmul(a, b, transpose_a, transpose_b)


def _axes(axis, reduction_indices):
    return axis if axis is not None else reduction_indices


def reduce_sum(x, axis=None, keep_dims=False, keepdims=None, reduction_indices=None):
    return O.reduce("sum", x, _axes(axis, reduction_indices), keep_dims if keepdims is None else keepdims)


def reduce_mean(x, axis=None, keep_dims=False, keepdims=None, reduction_indices=None):
    return O.reduce("mean", x, _axes(axis, reduction_indices), keep_dims if keepdims is None else keepdims)


def reduce_max(x, axis=None, keep_dims=False, keepdims=None, reduction_indices=None):
    return O.reduce("max", x, _axes(axis, reduction_indices), keep_dims if keepdims is None else keepdims)


def reduce_prod(x, axis=None, **_):
    raise NotImplementedError("tf.reduce_prod is only used on static shapes in the reference; use np.prod")


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    return O.random_normal(list(shape), mean, stddev)


def random_uniform(shape, minval=0, maxval=1, dtype=None, seed=None, name=None):
    return O.random_uniform(list(shape), minval, maxval)


def set_random_seed(seed):
    from .executor import RT
    RT.seed = int(seed)


def gradients(ys, xs, grad_ys=None):
    single = isinstance(xs, Tensor)
    out = O.gradients(ys, [xs] if single else list(xs), grad_ys)
    return out


def group(*ops):
    return Operation("group", attrs={"ops": [o for o in ops if o is not None]})


def assign(var, value):
    return Operation("assign", deps=[O.to_tensor(value)], attrs={"var": var})


@contextlib.contextmanager
def control_dependencies(_ops):
    yield


def cond(pred, true_fn, false_fn):
    raise NotImplementedError("tf.cond: the scripts always pass is_training=None, the branch is never built")


# ---- namespaces ------------------------------------------------------------------------------
class _NN(object):
    @staticmethod
    def relu(x):
        return O.unary("relu", x)

    @staticmethod
    def sigmoid(x):
        return O.unary("sigmoid", x)

    @staticmethod
    def tanh(x):
        return O.unary("tanh", x)

    @staticmethod
    def softsign(x):
        return O.unary("softsign", x)

    @staticmethod
    def softmax(x, dim=-1, axis=None):
        ax = axis if axis is not None else dim
        if ax not in (-1, len(x.shape) - 1):
            raise NotImplementedError("softmax over a non-last axis")
        return O.softmax(x)

    @staticmethod
    def bias_add(x, b, data_format=None):
        if data_format == "NCHW":
            return O.to_nchw(O.bias_add(O.to_nhwc(x), b))
        return O.bias_add(x, b)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(logits=None, labels=None, **_):
        """max(x,0) - x*z + log(1+exp(-|x|)); constant 0/1 labels (every reference call site, gan_inference.py:48-66,
        85-101) take the one-kernel path"""
        if labels.op == "const":
            v = labels.attrs["value"]
            if v.size and np.all(v == v.flat[0]):
                return O.unary("bce", logits, float(v.flat[0]))
        relu_x = O.unary("relu", logits)
        softplus = O.unary("log", O.unary("affine", O.unary("exp", O.unary("neg", O.unary("abs", logits))), 1.0, 1.0))
        return O.add(O.sub(relu_x, O.mul(logits, labels)), softplus)

    @staticmethod
    def softmax_cross_entropy_with_logits(labels=None, logits=None, **_):
        m = O.reduce("max", logits, [-1], keepdims=True)
        z = O.sub(logits, O.stop_gradient(m))
        lse = O.unary("log", O.reduce("sum", O.unary("exp", z), [-1], keepdims=True))
        logp = O.sub(z, lse)
        return O.unary("neg", O.reduce("sum", O.mul(labels, logp), [-1]))

    @staticmethod
    def moments(x, axes, keep_dims=False):
        mean = O.reduce("mean", x, axes, keepdims=True)
        var = O.reduce("mean", O.unary("square", O.sub(x, mean)), axes, keepdims=True)
        if not keep_dims:
            shp = [s for a, s in enumerate(x.shape) if a not in O._norm_axes(axes, len(x.shape))]
            return O.reshape(mean, shp), O.reshape(var, shp)
        return mean, var

    @staticmethod
    def batch_normalization(x, mean, variance, offset, scale, variance_epsilon):
        inv = O.unary("rsqrt", O.add(variance, float(variance_epsilon)))
        if scale is not None:
            inv = O.mul(inv, scale)
        y = O.mul(O.sub(x, mean), inv)
        return O.add(y, offset) if offset is not None else y

    @staticmethod
    def fused_batch_norm(x, scale, offset, epsilon=1e-3, data_format="NHWC", is_training=True, **_):
        if not is_training:
            raise NotImplementedError("inference-mode fused_batch_norm is dead code in the reference")
        xn = O.to_nhwc(x) if data_format == "NCHW" else x
        y = O.batchnorm(xn, scale, offset, epsilon)
        C = xn.shape[-1]
        out = O.to_nchw(y) if data_format == "NCHW" else y
        return out, O.aux(y, 1, (C,)), None

    @staticmethod
    def conv2d(input=None, filter=None, strides=None, padding="SAME", data_format="NHWC", **_):   # noqa: A002
        from .layers import conv2d_nchw, conv2d_nhwc
        if data_format == "NCHW":
            return conv2d_nchw(input, filter, strides[2], padding)
        return conv2d_nhwc(input, filter, strides[1], padding)

    @staticmethod
    def conv3d(input=None, filter=None, strides=None, padding="SAME", data_format="NDHWC", **_):   # noqa: A002
        from .layers import conv3d_ndhwc
        if data_format != "NDHWC" or strides[2] != strides[3] or strides[0] != 1 or strides[4] != 1:
            raise NotImplementedError("tf.nn.conv3d: NDHWC with strides [1, sl, s, s, 1] (tflib/ops/conv3d.py:33-39)")
        return conv3d_ndhwc(input, filter, int(strides[1]), int(strides[2]), padding)

    @staticmethod
    def conv2d_transpose(value=None, filter=None, output_shape=None, strides=None, padding="SAME", **_):   # noqa: A002
        from .layers import conv2d_transpose_nhwc
        return conv2d_transpose_nhwc(value, filter, list(output_shape), strides[1], padding)


nn = _NN()


class _Layers(object):
    @staticmethod
    def dropout(x, rate=0.5, training=False, **_):
        """tf.layers.dropout defaults to training=False -> identity; all 143 reference call sites rely on it."""
        if training is not False:
            raise NotImplementedError("dropout with training=True is not on the reference's path")
        return x


layers = _Layers()


class _Categorical(object):
    def __init__(self, probs=None, logits=None):
        if probs is None:
            raise NotImplementedError("Categorical(logits=...)")
        self.probs = O.to_tensor(probs)

    def sample(self, n):
        return O.categorical_sample(self.probs, n)


class _Distributions(object):
    Categorical = _Categorical


distributions = _Distributions()


class _Optimizer(object):
    kind = None

    def minimize(self, loss, var_list=None):
        if var_list is None:
            var_list = trainable_variables()
        var_list = list(var_list)
        grads = O.gradients(loss, var_list)
        # variables with no path to the loss (e.g. BN moving_* returned by params_with_name) get no update, like TF
        attrs = dict(self.hyper)
        attrs.update(vars=var_list, opt_id=self.opt_id)
        return Operation(self.kind, deps=grads, attrs=attrs)

    def compute_gradients(self, loss, var_list=None):
        var_list = list(var_list if var_list is not None else trainable_variables())
        return list(zip(O.gradients(loss, var_list), var_list))


class AdamOptimizer(_Optimizer):
    kind = "adam"

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **_):
        self.opt_id = next(_opt_ids)
        self.hyper = dict(lr=float(learning_rate), beta1=float(beta1), beta2=float(beta2), eps=float(epsilon))


class RMSPropOptimizer(_Optimizer):
    kind = "rmsprop"

    def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10, **_):
        if momentum != 0.0:
            raise NotImplementedError("RMSProp with momentum")
        self.opt_id = next(_opt_ids)
        self.hyper = dict(lr=float(learning_rate), decay=float(decay), eps=float(epsilon))


class Saver(object):
    """Registry + optimiser-state checkpoint (the reference only ever saves, at the last iteration:
    gmgan_inference_cifar10.py:465,548-549); restore is provided for completeness.

    Parameters are enumerated from the tflib registry and tf.Variable list, NOT from the runtime's lazily filled buffer
    table: build graph -> restore -> train works before any Session.run, and save() includes variables no plan has touched
    yet (e.g. batch-norm moving_mean / moving_variance).  Optimiser slots restored before their plan exists are parked in
    RT.pending_restore and applied when the optimiser step is emitted.  Missing / unexpected names raise."""

    @staticmethod
    def _variables():
        import tflib as lib
        nodes = {}
        for name, node in lib._params.items():
            nodes[name] = node
        for node in _trainable:
            if node.name is not None:
                nodes.setdefault(node.name, node)
        return nodes

    def save(self, sess, path):
        import torch
        from .executor import RT
        nodes = self._variables()
        for i, n in RT.param_nodes.items():
            if n.name is not None:
                nodes.setdefault(n.name, n)
        state = {"params": {name: RT.param_buffer(n).cpu() for name, n in nodes.items()},
                 "slots": {"%d/%s" % (k[0], RT.param_nodes[k[1]].name): (m.cpu(), v.cpu()) for k, (m, v) in RT.slots.items()},
                 "opt_state": {k: v.cpu() for k, v in RT.opt_state.items()}}
        torch.save(state, path)
        return path

    def restore(self, sess, path):
        import torch
        from .executor import RT
        state = torch.load(path)
        nodes = self._variables()
        for i, n in RT.param_nodes.items():
            if n.name is not None:
                nodes.setdefault(n.name, n)
        missing = sorted(set(nodes) - set(state["params"]))
        unexpected = sorted(set(state["params"]) - set(nodes))
        if missing or unexpected:
            raise KeyError("checkpoint %s does not match the graph: missing %s, unexpected %s" % (path, missing[:8], unexpected[:8]))
        for name, t in state["params"].items():
            buf = RT.param_buffer(nodes[name])
            if buf.numel() != t.numel():
                raise ValueError("checkpoint variable %s has %d elements, the graph's has %d" % (name, t.numel(), buf.numel()))
            buf.copy_(t)
        by_name = {n.name: i for i, n in RT.param_nodes.items()}
        pend = RT.pending_restore
        pend.setdefault("slots", {})
        pend.setdefault("opt_state", {})
        for key, (m, v) in state["slots"].items():
            oid, name = key.split("/", 1)
            k = (int(oid), by_name.get(name))
            if k in RT.slots:
                RT.slots[k][0].copy_(m)
                RT.slots[k][1].copy_(v)
            else:
                pend["slots"][key] = (m, v)
        for k, v in state["opt_state"].items():
            if k in RT.opt_state:
                RT.opt_state[k].copy_(v)
            else:
                pend["opt_state"][k] = v


class _Train(object):
    AdamOptimizer = AdamOptimizer
    RMSPropOptimizer = RMSPropOptimizer
    Saver = Saver


train = _Train()


class Session(object):
    """tf.Session over the plan compiler.  `deferred_fetches=True` (an extension; also settable as an attribute) makes
    `run` return executor.Deferred values: the device->host copy of every fetch is still enqueued by the call, but the host
    waits for it only when the value is first used, so the loop can feed step i+1 while step i executes."""

    def __init__(self, *_, **kw):
        self.deferred_fetches = bool(kw.get("deferred_fetches", False))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        from .executor import RT
        return RT.run(fetches, feed_dict, deferred=self.deferred_fetches)

    def close(self):
        pass


def reset_default_graph():
    from .executor import reset_runtime
    global _opt_ids
    del _trainable[:]
    _opt_ids = itertools.count()      # optimiser ids key the checkpointed slots: a rebuilt graph numbers them like a fresh process
    reset_runtime()
