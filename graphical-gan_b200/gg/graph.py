"""Deferred-graph IR of the host side.

The reference builds ONE static TensorFlow graph at import time and then calls
``session.run([cost, train_op], feed_dict)`` twice per iteration (gmgan_inference_cifar10.py:341-410,
:480-494).  This module is the B200-native counterpart of that graph: a tiny static-shape IR whose
nodes map 1:1 onto C-ABI kernel launches (gg/executor.py compiles a fetch set into a launch list and
captures it in a CUDA graph).  Nothing here touches the device.

Layout convention: a node's buffer is always contiguous in its *logical* shape.  Image ops work in
NHWC; the NCHW boundary of tflib is an explicit ``transpose`` node, and the constructors below sink
element-wise ops through those transposes / cancel inverse pairs, so a conv -> BN -> LeakyReLU -> conv
chain never leaves NHWC.
"""
import itertools

import numpy as np

_ids = itertools.count()


class DType(object):
    def __init__(self, name, np_dtype):
        self.name = name
        self.as_numpy_dtype = np_dtype

    def __repr__(self):
        return "tf." + self.name

    def __eq__(self, other):
        if isinstance(other, DType):
            return self.name == other.name
        if isinstance(other, str):
            return self.name == other
        return NotImplemented

    def __hash__(self):
        return hash(self.name)


float32 = DType("float32", np.float32)
int32 = DType("int32", np.int32)
int64 = DType("int64", np.int64)   # stored as int32 on the device


def as_dtype(d):
    if isinstance(d, DType):
        return float32 if d.name.startswith("float") else int32
    if d in ("float32", "float", np.float32, float, "float64", np.float64):
        return float32
    if d in ("int32", "int64", "int", np.int32, np.int64, int):
        return int32
    raise TypeError("unsupported dtype %r" % (d,))


class TensorShape(tuple):
    """tuple of ints with the bits of tf.TensorShape the scripts touch (`.ndims`, `.as_list()`, indexing)."""

    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)


class Tensor(object):
    """A node of the static graph (one output).  `op` names the launcher in gg/executor.py."""

    def __init__(self, op, inputs=(), attrs=None, shape=(), dtype=float32, name=None):
        self.id = next(_ids)
        self.op = op
        self.inputs = tuple(inputs)
        self.attrs = dict(attrs or {})
        self.shape = TensorShape(int(s) for s in shape)
        self.dtype = dtype
        self.name = name or "%s_%d" % (op, self.id)

    # --- tf.Tensor surface -------------------------------------------------------------------
    def get_shape(self):
        return self.shape

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def __repr__(self):
        return "<gg.Tensor %s %s %s>" % (self.name, self.op, tuple(self.shape))

    __hash__ = object.__hash__

    def __add__(self, o):
        from . import ops
        return ops.add(self, o)

    def __radd__(self, o):
        from . import ops
        return ops.add(o, self)

    def __sub__(self, o):
        from . import ops
        return ops.sub(self, o)

    def __rsub__(self, o):
        from . import ops
        return ops.sub(o, self)

    def __mul__(self, o):
        from . import ops
        return ops.mul(self, o)

    def __rmul__(self, o):
        from . import ops
        return ops.mul(o, self)

    def __truediv__(self, o):
        from . import ops
        return ops.div(self, o)

    def __rtruediv__(self, o):
        from . import ops
        return ops.div(o, self)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __neg__(self):
        from . import ops
        return ops.unary("neg", self)

    def __pow__(self, p):
        from . import ops
        return ops.pow_(self, p)

    def __getitem__(self, idx):
        from . import ops
        return ops.getitem(self, idx)

    def __bool__(self):
        raise TypeError("a graph Tensor has no truth value; use `is not None`")

    def __iter__(self):
        raise TypeError("a graph Tensor is not iterable")


class Operation(object):
    """A runnable handle without a value (train ops, tf.group, assign)."""

    def __init__(self, kind, deps=(), attrs=None, name=None):
        self.id = next(_ids)
        self.kind = kind
        self.deps = tuple(deps)       # Tensors that must be computed before it runs
        self.attrs = dict(attrs or {})
        self.name = name or "%s_%d" % (kind, self.id)

    def __repr__(self):
        return "<gg.Operation %s>" % self.name


def prod(xs):
    n = 1
    for x in xs:
        n *= int(x)
    return n
