"""Graph constructors (shape inference + build-time fusion) and reverse-mode gradient rules.

Every constructor returns a gg.graph.Tensor.  Gradients are built symbolically from the same
constructors (like tf.gradients), so a gradient can itself be differentiated — this is what the WGAN-GP
penalty needs (gan_inference_svhn.py:342-357: tf.gradients inside the loss that minimize() differentiates).

Build-time rewrites (all value-preserving):
  * element-wise ops are sunk through `transpose` nodes and inverse transposes cancel, so the NCHW surface
    of tflib.ops.conv2d/deconv2d/batchnorm costs no layout kernels between image layers;
  * `maximum(alpha*x, x)` (the scripts' LeakyReLU, gmgan_inference_cifar10.py:122-123) becomes one op;
  * an activation applied to a conv / matmul / batch-norm node is folded into that kernel's epilogue
    (the un-fused original stays in the graph and is pruned if nothing else reads it).
"""
import os

import numpy as np

from .graph import Tensor, Operation, float32, int32, as_dtype, prod

NHWC2NCHW = (0, 3, 1, 2)
NCHW2NHWC = (0, 2, 3, 1)
FUSABLE = ("conv", "matmul", "bn")
ACT_FNS = ("relu", "leaky", "tanh", "sigmoid")


# ------------------------------------------------------------------------------------------------
# leaves
# ------------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    arr = np.asarray(value)
    if dtype is not None:
        dt = as_dtype(dtype)
    else:
        dt = float32 if arr.dtype.kind == "f" else int32
    arr = arr.astype(dt.as_numpy_dtype)
    if shape is not None:
        arr = np.broadcast_to(arr, tuple(shape)).copy()
    return Tensor("const", (), {"value": np.ascontiguousarray(arr)}, arr.shape, dt, name)


def placeholder(dtype, shape=None, name=None):
    if shape is None or any(s is None for s in shape):
        raise ValueError("placeholders need a fully static shape (the execution plan is static)")
    return Tensor("placeholder", (), {}, tuple(shape), as_dtype(dtype), name)


def variable(value, name=None, trainable=True):
    arr = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
    return Tensor("param", (), {"init": arr, "trainable": trainable}, arr.shape, float32, name)


def to_tensor(x, like=None):
    if isinstance(x, Tensor):
        return x
    return constant(np.asarray(x, dtype=np.float32 if like is None or like.dtype == float32 else np.int32))


def is_scalar_const(x):
    return isinstance(x, (int, float, np.floating, np.integer)) or (isinstance(x, np.ndarray) and x.ndim == 0)


# ------------------------------------------------------------------------------------------------
# layout ops
# ------------------------------------------------------------------------------------------------
def transpose(x, perm):
    perm = tuple(int(p) for p in perm)
    if perm == tuple(range(len(perm))):
        return x
    if x.op == "transpose":
        inner = x.attrs["perm"]
        comp = tuple(inner[p] for p in perm)
        return transpose(x.inputs[0], comp)
    shape = tuple(x.shape[p] for p in perm)
    if len(perm) > 4:
        # the layout kernels take rank <= 4: input axes that stay adjacent and in order travel as ONE axis
        # (tf.transpose(x, [0, 1, 3, 4, 2]) of the SSGAN 3dcnn critic, ssgan_inference_moving_mnist.py:356, is [NL, C, HW] -> [NL, HW, C])
        groups = []
        for q in perm:
            if groups and groups[-1][-1] + 1 == q:
                groups[-1].append(q)
            else:
                groups.append([q])
        if len(groups) > 4:
            raise NotImplementedError("transpose %s of a rank-%d tensor does not reduce to rank <= 4" % (perm, len(perm)))
        in_order = sorted(range(len(groups)), key=lambda g: groups[g][0])
        merged = reshape(x, [prod(x.shape[a] for a in groups[g]) for g in in_order])
        return reshape(transpose(merged, [in_order.index(g) for g in range(len(groups))]), shape)
    return Tensor("transpose", (x,), {"perm": perm}, shape, x.dtype)


def to_nhwc(x):
    return transpose(x, NCHW2NHWC)


def to_nchw(x):
    return transpose(x, NHWC2NCHW)


def reshape(x, shape):
    shape = [int(s) for s in shape]
    if -1 in shape:
        known = prod(s for s in shape if s != -1)
        shape[shape.index(-1)] = x.size // known
    shape = tuple(shape)
    if prod(shape) != x.size:
        raise ValueError("cannot reshape %s to %s" % (tuple(x.shape), shape))
    if shape == tuple(x.shape):
        return x
    if x.op == "reshape":
        return reshape(x.inputs[0], shape)
    return Tensor("reshape", (x,), {}, shape, x.dtype)


def expand_dims(x, axis):
    shape = list(x.shape)
    if axis < 0:
        axis += len(shape) + 1
    shape.insert(axis, 1)
    return reshape(x, shape)


def concat(values, axis):
    values = [to_tensor(v) for v in values]
    if len(values) == 1:
        return values[0]
    nd = len(values[0].shape)
    if axis < 0:
        axis += nd
    shape = list(values[0].shape)
    for v in values[1:]:
        if len(v.shape) != nd or any(a != b for i, (a, b) in enumerate(zip(v.shape, shape)) if i != axis) or v.dtype != values[0].dtype:
            raise ValueError("concat along axis %d: shapes %s do not agree off the axis" % (axis, [tuple(t.shape) for t in values]))
    shape[axis] = sum(v.shape[axis] for v in values)
    return Tensor("concat", values, {"axis": axis}, shape, values[0].dtype)


def slice_axis(x, axis, start, size):
    if axis < 0:
        axis += len(x.shape)
    if start == 0 and size == x.shape[axis]:
        return x
    if x.op == "transpose":
        # crop in the producer's own layout (gmgan_inference_mnist.py:179 crops the NCHW view `output[:, :, :7, :7]` between
        # two deconvolutions: the data stays NHWC and no layout kernel is materialised)
        perm = x.attrs["perm"]
        return transpose(slice_axis(x.inputs[0], perm[axis], start, size), perm)
    shape = list(x.shape)
    shape[axis] = size
    return Tensor("slice", (x,), {"axis": axis, "start": int(start), "size": int(size)}, shape, x.dtype)


def pad_axis(x, axis, before, total):
    """inverse of slice_axis: zeros of extent `total` along axis with x placed at `before`."""
    shape = list(x.shape)
    shape[axis] = total
    return Tensor("pad", (x,), {"axis": axis, "start": int(before), "total": int(total)}, shape, x.dtype)


def getitem(x, idx):
    if not isinstance(idx, tuple):
        idx = (idx,)
    out = x
    axis = 0
    squeeze = []
    for it in idx:
        if it is None:
            raise NotImplementedError("None indexing: use tf.expand_dims")
        if isinstance(it, slice):
            n = out.shape[axis]
            start, stop, step = it.indices(n)
            if step != 1:
                raise NotImplementedError("strided slices")
            out = slice_axis(out, axis, start, stop - start)
        else:
            i = int(it)
            if i < 0:
                i += out.shape[axis]
            out = slice_axis(out, axis, i, 1)
            squeeze.append(axis)
        axis += 1
    if squeeze:
        out = reshape(out, [s for a, s in enumerate(out.shape) if a not in squeeze])
    return out


def tile(x, multiples):
    multiples = [int(m) for m in multiples]
    if all(m == 1 for m in multiples):
        return x
    shape = [s * m for s, m in zip(x.shape, multiples)]
    return Tensor("tile", (x,), {"multiples": tuple(multiples)}, shape, x.dtype)


def stop_gradient(x):
    return Tensor("stop_gradient", (x,), {}, x.shape, x.dtype)


def identity(x):
    return x


def aux(parent, k, shape):
    return Tensor("aux", (parent,), {"k": k}, shape, float32)


# ------------------------------------------------------------------------------------------------
# element-wise
# ------------------------------------------------------------------------------------------------
def _fuse_act(x, fn, alpha):
    """fold an activation into the producing conv / matmul / bn kernel when it has none yet"""
    if x.op in FUSABLE and x.attrs.get("act") is None and x.attrs.get("mode") != "wgrad":
        attrs = dict(x.attrs)
        attrs["act"] = fn
        attrs["alpha"] = float(alpha)
        return Tensor(x.op, x.inputs, attrs, x.shape, x.dtype)
    return None


def unary(fn, x, a=0.0, b=0.0):
    x = to_tensor(x)
    if x.op == "transpose":
        return transpose(unary(fn, x.inputs[0], a, b), x.attrs["perm"])
    if x.op == "reshape" and fn in ACT_FNS and x.inputs[0].op in FUSABLE:
        f = _fuse_act(x.inputs[0], fn, a)
        if f is not None:
            return reshape(f, x.shape)
    if fn in ACT_FNS:
        f = _fuse_act(x, fn, a)
        if f is not None:
            return f
    if fn == "affine":
        if a == 1.0 and b == 0.0:
            return x
        if x.op == "unary" and x.attrs["fn"] == "affine" and b == 0.0 and x.attrs["b"] == 0.0:
            # only exact merges: powers of two commute with rounding
            if float(a) in (2.0, 0.5, -1.0, 4.0, 0.25):
                return unary("affine", x.inputs[0], x.attrs["a"] * a, 0.0)
    return Tensor("unary", (x,), {"fn": fn, "a": float(a), "b": float(b)}, x.shape, float32)


def _bshape(sa, sb):
    nd = max(len(sa), len(sb))
    sa = (1,) * (nd - len(sa)) + tuple(sa)
    sb = (1,) * (nd - len(sb)) + tuple(sb)
    out = []
    for x, y in zip(sa, sb):
        if x != y and x != 1 and y != 1:
            raise ValueError("shapes %s and %s do not broadcast" % (sa, sb))
        out.append(max(x, y))
    return tuple(out)


def binary(fn, a, b, alpha=0.0):
    a, b = to_tensor(a), to_tensor(b)
    if a.op == "transpose" and b.op == "transpose" and a.attrs["perm"] == b.attrs["perm"] and \
            a.inputs[0].shape == b.inputs[0].shape:
        return transpose(binary(fn, a.inputs[0], b.inputs[0], alpha), a.attrs["perm"])
    if fn in ("max",):
        # LeakyReLU written as tf.maximum(alpha*x, x)
        for p, q in ((a, b), (b, a)):
            if p.op == "unary" and p.attrs["fn"] == "affine" and p.attrs["b"] == 0.0 and p.inputs[0] is q \
                    and 0.0 < p.attrs["a"] < 1.0:
                return unary("leaky", q, p.attrs["a"])
    shape = _bshape(a.shape, b.shape)
    return Tensor("binary", (a, b), {"fn": fn, "alpha": float(alpha)}, shape, float32)


def add(a, b):
    if is_scalar_const(b):
        return a if float(b) == 0.0 else unary("affine", a, 1.0, float(b))
    if is_scalar_const(a):
        return b if float(a) == 0.0 else unary("affine", b, 1.0, float(a))
    return binary("add", a, b)


def sub(a, b):
    if is_scalar_const(b):
        return unary("affine", a, 1.0, -float(b))
    if is_scalar_const(a):
        return unary("affine", b, -1.0, float(a))
    return binary("sub", a, b)


def mul(a, b):
    if is_scalar_const(b):
        return unary("affine", a, float(b), 0.0)
    if is_scalar_const(a):
        return unary("affine", b, float(a), 0.0)
    return binary("mul", a, b)


def div(a, b):
    if is_scalar_const(b):
        return unary("divc", a, float(b))
    if is_scalar_const(a):
        return unary("rdivc", b, float(a))
    return binary("div", a, b)


def pow_(x, p):
    if is_scalar_const(p):
        if float(p) == 2.0:
            return unary("square", x)
        if float(p) == 1.0:
            return x
        return unary("pow", x, float(p))
    return binary("pow", x, p)


def maximum(a, b):
    if is_scalar_const(b):
        b = constant(np.float32(b))
    if is_scalar_const(a):
        a = constant(np.float32(a))
    return binary("max", a, b)


def add_n(ts):
    ts = [t for t in ts if t is not None]
    if not ts:
        return None
    if len(ts) == 1:
        return ts[0]
    if all(t.op == "transpose" and t.attrs["perm"] == ts[0].attrs["perm"] for t in ts):
        return transpose(add_n([t.inputs[0] for t in ts]), ts[0].attrs["perm"])
    if all(t.op == "pad" and t.attrs["axis"] == ts[0].attrs["axis"] and t.attrs["total"] == ts[0].attrs["total"] for t in ts):
        # gradients of the slices that tile a tensor (the two halves of a sibling-batched tower, gg/rewrite.py): the sum of
        # zero-padded pieces with disjoint, covering extents IS their concatenation — no zero fill, no add
        axis, total = ts[0].attrs["axis"], ts[0].attrs["total"]
        parts = sorted(ts, key=lambda t: t.attrs["start"])
        pos = 0
        for t in parts:
            if t.attrs["start"] != pos:
                break
            pos += t.inputs[0].shape[axis]
        else:
            if pos == total:
                return concat([t.inputs[0] for t in parts], axis)
    shape = ts[0].shape
    for t in ts:
        if tuple(t.shape) != tuple(shape):
            raise ValueError("add_n shape mismatch %s vs %s" % (t.shape, shape))
    out = None
    for i in range(0, len(ts), 16):
        chunk = ts[i:i + 16] + ([out] if out is not None else [])
        if len(chunk) > 16:
            out = Tensor("add_n", chunk[:16], {}, shape, float32)
            out = Tensor("add_n", [out] + chunk[16:], {}, shape, float32)
        else:
            out = Tensor("add_n", chunk, {}, shape, float32)
    return out


def cast(x, dtype):
    dt = as_dtype(dtype)
    x = to_tensor(x)
    if x.dtype == dt:
        return x
    return Tensor("cast", (x,), {}, x.shape, dt)


def zeros_like(x):
    return constant(np.zeros(x.shape, np.float32))


def ones_like(x):
    return constant(np.ones(x.shape, np.float32))


def broadcast_to(x, shape):
    shape = tuple(shape)
    if tuple(x.shape) == shape:
        return x
    return Tensor("broadcast", (x,), {}, shape, float32)


# ------------------------------------------------------------------------------------------------
# reductions
# ------------------------------------------------------------------------------------------------
def _norm_axes(axes, nd):
    if axes is None:
        return tuple(range(nd))
    if isinstance(axes, (int, np.integer)):
        axes = [axes]
    return tuple(sorted(set(a + nd if a < 0 else a for a in axes)))


def reduce(fn, x, axes=None, keepdims=False):
    x = to_tensor(x)
    nd = len(x.shape)
    axes = _norm_axes(axes, nd)
    if not axes:
        return x
    if x.op == "transpose" and len(axes) == nd:
        x = x.inputs[0]   # a full reduction does not care about the layout
    # contiguous groups, reduced from the last group to the first so earlier axis numbers stay valid
    groups = []
    for a in axes:
        if groups and groups[-1][-1] == a - 1:
            groups[-1].append(a)
        else:
            groups.append([a])
    out = x
    total = prod(x.shape[a] for a in axes)
    fused_mean = fn == "mean" and len(groups) == 1      # one kernel: sum and divide by the row count
    for g in reversed(groups):
        shape = list(out.shape)
        kept = shape[:g[0]] + [1] * len(g) + shape[g[-1] + 1:]
        sub_fn = "mean" if fused_mean else ("sum" if fn == "mean" else fn)
        out = Tensor("reduce", (out,), {"fn": sub_fn, "axes": tuple(g)}, kept, float32)
    if fn == "mean" and not fused_mean:
        out = unary("divc", out, float(total))
    if not keepdims:
        out = reshape(out, [s for a, s in enumerate(x.shape) if a not in axes])
    return out


def softmax(x):
    return Tensor("softmax", (x,), {}, x.shape, float32)


def argmax(x, axis=-1):
    nd = len(x.shape)
    if axis < 0:
        axis += nd
    if axis != nd - 1:
        raise NotImplementedError("argmax over a non-last axis")
    return Tensor("argmax", (x,), {}, x.shape[:-1], int32)


def one_hot(indices, depth):
    indices = to_tensor(indices)
    if indices.dtype != int32:
        indices = cast(indices, int32)
    return Tensor("one_hot", (indices,), {"depth": int(depth)}, tuple(indices.shape) + (int(depth),), float32)


# ------------------------------------------------------------------------------------------------
# random
# ------------------------------------------------------------------------------------------------
def random_normal(shape, mean=0.0, stddev=1.0):
    return Tensor("random", (), {"kind": "normal", "a": float(mean), "b": float(stddev)}, tuple(int(s) for s in shape), float32)


def random_uniform(shape, minval=0.0, maxval=1.0):
    return Tensor("random", (), {"kind": "uniform", "a": float(minval), "b": float(maxval)}, tuple(int(s) for s in shape), float32)


def categorical_sample(probs, n):
    probs = to_tensor(probs)
    return Tensor("random", (probs,), {"kind": "categorical"}, (int(n),), int32)


# ------------------------------------------------------------------------------------------------
# dense / conv / batch-norm
# ------------------------------------------------------------------------------------------------
def matmul(a, b, transpose_a=False, transpose_b=False, bias=None):
    a, b = to_tensor(a), to_tensor(b)
    M = a.shape[1] if transpose_a else a.shape[0]
    Ka = a.shape[0] if transpose_a else a.shape[1]
    Kb = b.shape[1] if transpose_b else b.shape[0]
    N = b.shape[0] if transpose_b else b.shape[1]
    if Ka != Kb:
        raise ValueError("matmul inner dimensions differ: %s x %s" % (tuple(a.shape), tuple(b.shape)))
    if Ka == 1 and bias is None and os.environ.get("GG_RANK1_MUL", "1") != "0":
        # a rank-1 product is a broadcast multiply: out[m, n] = a[m] * b[n] — the input gradient of a 1-output dense layer
        # (`Discriminator.Output`, gmgan_inference_cifar10.py:300, backward: dy [B,1] x W^T [1,512]).  fmaf(a, b, 0) of the GEMM
        # kernels and a * b round identically; as an element-wise node it fuses with the activation gradient that follows it
        return binary("mul", reshape(a, (M, 1)), reshape(b, (1, N)))
    inputs = (a, b) if bias is None else (a, b, bias)
    return Tensor("matmul", inputs, {"ta": bool(transpose_a), "tb": bool(transpose_b), "act": None, "alpha": 0.0}, (M, N), float32)


def bias_add(x, b):
    """x [..., C] + b [C]; folded into the producing conv / matmul when that node has neither bias nor activation"""
    if x.op in ("conv", "matmul") and len(x.inputs) == 2 and x.attrs.get("act") is None and x.attrs.get("mode") != "wgrad":
        return Tensor(x.op, x.inputs + (b,), x.attrs, x.shape, x.dtype)
    return binary("add", x, reshape(b, (1,) * (len(x.shape) - 1) + (b.size,)))


def conv(mode, a, b, geom, bias=None):
    """mode fwd:   a = x [B,H,W,Ci],    b = w [k,k,Ci,Co]  -> y  [B,Ho,Wo,Co]
       mode dgrad: a = dy [B,Ho,Wo,Co], b = w [k,k,Ci,Co]  -> dx [B,H,W,Ci]
       mode wgrad: a = x [B,H,W,Ci],    b = dy [B,Ho,Wo,Co] -> dw [k,k,Ci,Co]"""
    g = dict(geom)
    if mode == "fwd":
        shape = (g["B"], g["Ho"], g["Wo"], g["Co"])
    elif mode == "dgrad":
        shape = (g["B"], g["H"], g["W"], g["Ci"])
    else:
        shape = (g["k"], g["k"], g["Ci"], g["Co"])
    g.update(mode=mode, act=None, alpha=0.0)
    inputs = (a, b) if bias is None else (a, b, bias)
    return Tensor("conv", inputs, g, shape, float32)


def batchnorm(x, gamma, beta, eps=1e-5):
    """x [..., C] normalised per channel over all leading dims with batch statistics"""
    return Tensor("bn", (x, gamma, beta), {"eps": float(eps), "act": None, "alpha": 0.0}, x.shape, float32)


# ------------------------------------------------------------------------------------------------
# gradients
# ------------------------------------------------------------------------------------------------
def _unbroadcast(g, shape):
    """sum g over the axes that were broadcast to reach g.shape from `shape`"""
    shape = tuple(shape)
    if tuple(g.shape) == shape:
        return g
    nd = len(g.shape)
    padded = (1,) * (nd - len(shape)) + shape
    axes = [i for i in range(nd) if padded[i] == 1 and g.shape[i] != 1]
    out = reduce("sum", g, axes, keepdims=True) if axes else g
    return reshape(out, shape)


def _act_grad(y, g, act, alpha):
    if act is None:
        return g
    return binary(act + "_grad", y, g, alpha)


def _grad_unary(n, g, need):
    x = n.inputs[0]
    fn, a = n.attrs["fn"], n.attrs["a"]
    if fn in ("relu", "leaky", "tanh", "sigmoid"):
        return [binary(fn + "_grad", n, g, a)]
    if fn == "copy":
        return [g]
    if fn == "exp":
        return [mul(g, n)]
    if fn == "log":
        return [div(g, x)]
    if fn == "sqrt":
        return [div(mul(g, 0.5), n)]
    if fn == "square":
        return [mul(mul(g, x), 2.0)]
    if fn == "neg":
        return [unary("neg", g)]
    if fn == "abs":
        return [binary("abs_grad", x, g)]
    if fn == "affine":
        return [mul(g, a)]
    if fn == "divc":
        return [div(g, a)]
    if fn == "rdivc":   # a/x -> -a/x^2
        return [mul(g, mul(div(n, x), -1.0))]
    if fn == "pow":
        return [mul(g, mul(unary("pow", x, a - 1.0), a))]
    if fn == "recip":
        return [unary("neg", mul(g, unary("square", n)))]
    if fn == "rsqrt":
        return [mul(g, mul(mul(n, unary("square", n)), -0.5))]
    if fn == "bce":
        return [binary("bce_grad", x, g, a)]
    if fn == "clip":
        lo, hi = a, n.attrs["b"]
        m = mul(binary("ge_mask", x, constant(np.float32(lo))), binary("ge_mask", constant(np.float32(hi)), x))
        return [mul(g, m)]
    if fn == "softsign":
        return [div(g, unary("square", unary("affine", unary("abs", x), 1.0, 1.0)))]
    if fn == "sign":
        return [None]
    raise NotImplementedError("gradient of unary %s" % fn)


def _grad_binary(n, g, need):
    a, b = n.inputs
    fn, alpha = n.attrs["fn"], n.attrs["alpha"]
    ga = gb = None
    if fn == "add":
        ga, gb = g, g
    elif fn == "sub":
        ga, gb = g, (unary("neg", g) if need[1] else None)
    elif fn == "mul":
        ga = mul(g, b) if need[0] else None
        gb = mul(g, a) if need[1] else None
    elif fn == "div":
        ga = div(g, b) if need[0] else None
        gb = unary("neg", mul(g, div(n, b))) if need[1] else None
    elif fn in ("max", "min"):
        m = binary("ge_mask", a, b) if fn == "max" else binary("ge_mask", b, a)   # tf: x >= y picks x
        ga = mul(g, m) if need[0] else None
        gb = mul(g, unary("affine", m, -1.0, 1.0)) if need[1] else None
    elif fn in ("relu_grad", "leaky_grad"):
        # f(y, g) = mask(y) * g : piecewise linear in g, zero derivative w.r.t. y almost everywhere
        ga, gb = None, (binary(fn, a, g, alpha) if need[1] else None)
    elif fn == "tanh_grad":      # (1-y^2) g
        ga = mul(mul(g, mul(a, b)), -2.0) if need[0] else None
        gb = binary("tanh_grad", a, g) if need[1] else None
    elif fn == "sigmoid_grad":   # y(1-y) g
        ga = mul(mul(g, b), unary("affine", a, -2.0, 1.0)) if need[0] else None
        gb = binary("sigmoid_grad", a, g) if need[1] else None
    elif fn == "bce_grad":       # (sigmoid(x)-z) g
        s = unary("sigmoid", a)
        ga = mul(binary("sigmoid_grad", s, b), g) if need[0] else None
        gb = binary("bce_grad", a, g, alpha) if need[1] else None
    elif fn == "abs_grad":
        ga, gb = None, (binary("abs_grad", a, g) if need[1] else None)
    elif fn in ("ge_mask", "gt_mask"):
        return [None, None]
    else:
        raise NotImplementedError("gradient of binary %s" % fn)
    return [_unbroadcast(ga, a.shape) if ga is not None and need[0] else None,
            _unbroadcast(gb, b.shape) if gb is not None and need[1] else None]


def _grad_reduce(n, g, need):
    x = n.inputs[0]
    fn = n.attrs["fn"]
    if fn == "sum":
        return [broadcast_to(g, x.shape)]
    if fn == "mean":
        count = prod(x.shape[a] for a in n.attrs["axes"])
        return [broadcast_to(div(g, float(count)), x.shape)]
    if fn == "max":
        m = binary("ge_mask", x, broadcast_to(n, x.shape))
        return [mul(broadcast_to(g, x.shape), m)]
    raise NotImplementedError(fn)


def _grad_matmul(n, g, need):
    a, b = n.inputs[0], n.inputs[1]
    ta, tb = n.attrs["ta"], n.attrs["tb"]
    g = _act_grad(n, g, n.attrs["act"], n.attrs["alpha"])
    out = [None] * len(n.inputs)
    if need[0]:
        if not ta:
            out[0] = matmul(g, b, False, not tb)        # dA = g op(B)^T
        else:
            out[0] = matmul(b, g, tb, True)             # dA^T: A stored [K,M] -> op(B) g^T
    if need[1]:
        if not tb:
            out[1] = matmul(a, g, not ta, False)        # dB = op(A)^T g
        else:
            out[1] = matmul(g, a, True, ta)             # B stored [N,K] -> g^T op(A)
    if len(n.inputs) == 3 and need[2]:
        out[2] = reshape(reduce("sum", g, [0]), n.inputs[2].shape)
    return out


def _geom(n):
    return {k: n.attrs[k] for k in ("B", "H", "W", "Ci", "Co", "k", "stride", "pad_t", "pad_l", "Ho", "Wo")}


def _grad_conv(n, g, need):
    mode = n.attrs["mode"]
    geom = _geom(n)
    out = [None] * len(n.inputs)
    if mode == "fwd":
        x, w = n.inputs[0], n.inputs[1]
        g = _act_grad(n, g, n.attrs["act"], n.attrs["alpha"])
        if need[0]:
            out[0] = conv("dgrad", g, w, geom)
        if need[1]:
            out[1] = conv("wgrad", x, g, geom)
        if len(n.inputs) == 3 and need[2]:
            out[2] = reshape(reduce("sum", reshape(g, (-1, geom["Co"])), [0]), n.inputs[2].shape)
    elif mode == "dgrad":
        dy, w = n.inputs[0], n.inputs[1]
        g = _act_grad(n, g, n.attrs["act"], n.attrs["alpha"])
        if need[0]:
            out[0] = conv("fwd", g, w, geom)
        if need[1]:
            out[1] = conv("wgrad", g, dy, geom)
        if len(n.inputs) == 3 and need[2]:
            out[2] = reshape(reduce("sum", reshape(g, (-1, geom["Ci"])), [0]), n.inputs[2].shape)
    else:  # wgrad(x, dy) -> dw ; g has the filter's shape
        x, dy = n.inputs
        if need[0]:
            out[0] = conv("dgrad", dy, g, geom)
        if need[1]:
            out[1] = conv("fwd", x, g, geom)
    return out


def _grad_bn(n, g, need):
    x, gamma, beta = n.inputs
    C = x.shape[-1]
    mean, rstd = aux(n, 1, (C,)), aux(n, 2, (C,))
    bg = Tensor("bn_grad", (g, x, n, mean, rstd, gamma), {"act": n.attrs["act"], "alpha": n.attrs["alpha"]}, x.shape, float32)
    return [bg, reshape(aux(bg, 1, (C,)), gamma.shape), reshape(aux(bg, 2, (C,)), beta.shape)]


def substitute(roots, mapping):
    """clone the expression graph above the keys of `mapping` ({node: replacement}); untouched sub-graphs are shared"""
    memo = {k.id: v for k, v in mapping.items()}

    def rec(t):
        if t.id in memo:
            return memo[t.id]
        new_in = [rec(i) for i in t.inputs]
        r = t if all(a is b for a, b in zip(new_in, t.inputs)) else Tensor(t.op, new_in, t.attrs, t.shape, t.dtype)
        memo[t.id] = r
        return r
    return [rec(t) if t is not None else None for t in roots]


def _grad_bn_grad(n, g, need):
    """Second-order gradient through batch norm: gan_inference_mnist.py MODE='wali-gp' puts BN inside the critic (:225,230)
    and differentiates tf.gradients(D(x_hat), x_hat) again (:346-357).  The first-order node dx = bn_grad(gy, x, ...) is
    re-expressed over proxy leaves with primitive ops whose gradient rules exist — statistics recomputed from x, so the
    dependence of mean / rstd on x is differentiated too:
        xh = (x - mean(x)) rsqrt(var(x) + eps);  ga = act'(y) gy;  dx = gamma rstd (ga - mean(ga) - xh mean(ga xh))
    — that expression is differentiated symbolically w.r.t. (gy, x, gamma) and the proxies are replaced by the real nodes.
    The fused activation's mask act'(y) is piecewise constant in x (zero derivative almost everywhere, like TF's ReluGrad).
    Gradients flowing into the dgamma / dbeta outputs of a bn_grad node are not propagated (no reference graph does that)."""
    gy, x, y, mean, rstd, gamma = n.inputs
    bn_node = y if y.op == "bn" else None
    eps = bn_node.attrs["eps"] if bn_node is not None else 1e-5
    act, alpha = n.attrs["act"], n.attrs["alpha"]
    C = x.shape[-1]
    axes = list(range(len(x.shape) - 1))
    px = placeholder(float32, x.shape, name="bn2_x")
    pgy = placeholder(float32, gy.shape, name="bn2_gy")
    pgamma = placeholder(float32, (C,), name="bn2_gamma")
    py = placeholder(float32, y.shape, name="bn2_y")
    mu = reduce("mean", px, axes, keepdims=True)
    xc = sub(px, mu)
    var = reduce("mean", unary("square", xc), axes, keepdims=True)
    rs = unary("rsqrt", unary("affine", var, 1.0, float(eps)))
    xh = mul(xc, rs)
    ga = _act_grad(py, pgy, act, alpha)
    m1 = reduce("mean", ga, axes, keepdims=True)
    m2 = reduce("mean", mul(ga, xh), axes, keepdims=True)
    gam = reshape(pgamma, (1,) * len(axes) + (C,))
    dx = mul(mul(gam, rs), sub(sub(ga, m1), mul(xh, m2)))
    d_gy, d_x, d_gamma = gradients(dx, [pgy, px, pgamma], grad_ys=[g])
    d_gy, d_x, d_gamma = substitute([d_gy, d_x, d_gamma], {px: x, pgy: gy, pgamma: reshape(gamma, (C,)), py: y})
    out = [None] * 6
    if need[0]:
        out[0] = d_gy
    if need[1]:
        out[1] = d_x
    if need[5] and d_gamma is not None:
        out[5] = reshape(d_gamma, gamma.shape)
    return out


def _grad_concat(n, g, need):
    axis = n.attrs["axis"]
    out, start = [], 0
    # Row-concatenated inputs of a sibling-batched tower where only SOME pieces need a gradient (D(fake_x) vs D(real_x):
    # real_x is data): do not compute the producing conv-dgrad / dense-dgrad for rows nobody reads — run it on the row
    # slice of ITS input instead (an axis-0 slice is a zero-copy view).
    partial_rows = axis == 0 and not all(need) and g.op in ("conv", "matmul") and \
        (g.attrs.get("mode", "fwd") in ("fwd", "dgrad")) and not g.attrs.get("ta", False)
    for inp, nd in zip(n.inputs, need):
        size = inp.shape[axis]
        if not nd:
            out.append(None)
        elif partial_rows:
            out.append(_rows_of(g, start, size))
        else:
            out.append(slice_axis(g, axis, start, size))
        start += size
    return out


def _rows_of(g, start, size):
    """rows [start, start+size) of a conv / dense node, computed from the same rows of its activation input"""
    x = slice_axis(g.inputs[0], 0, start, size)
    if g.op == "conv":
        geom = _geom(g)
        geom["B"] = size
        t = conv(g.attrs["mode"], x, g.inputs[1], geom, g.inputs[2] if len(g.inputs) == 3 else None)
    else:
        t = matmul(x, g.inputs[1], False, g.attrs["tb"], g.inputs[2] if len(g.inputs) == 3 else None)
    t.attrs["act"], t.attrs["alpha"] = g.attrs["act"], g.attrs["alpha"]
    return t


def _grad_tile(n, g, need):
    x = n.inputs[0]
    mult = n.attrs["multiples"]
    shape, axes = [], []
    for i, (s, m) in enumerate(zip(x.shape, mult)):
        if m > 1:
            axes.append(len(shape))
            shape.append(m)
        shape.append(s)
    return [reshape(reduce("sum", reshape(g, shape), axes), x.shape)]


def _grad_softmax(n, g, need):
    return [Tensor("softmax_grad", (n, g), {}, n.shape, float32)]


def _inv_perm(perm):
    inv = [0] * len(perm)
    for i, p in enumerate(perm):
        inv[p] = i
    return tuple(inv)


GRADS = {
    "unary": _grad_unary,
    "binary": _grad_binary,
    "reduce": _grad_reduce,
    "matmul": _grad_matmul,
    "conv": _grad_conv,
    "bn": _grad_bn,
    "bn_grad": _grad_bn_grad,
    "concat": _grad_concat,
    "tile": _grad_tile,
    "softmax": _grad_softmax,
    "reshape": lambda n, g, need: [reshape(g, n.inputs[0].shape)],
    "transpose": lambda n, g, need: [transpose(g, _inv_perm(n.attrs["perm"]))],
    "slice": lambda n, g, need: [pad_axis(g, n.attrs["axis"], n.attrs["start"], n.inputs[0].shape[n.attrs["axis"]])],
    "pad": lambda n, g, need: [slice_axis(g, n.attrs["axis"], n.attrs["start"], n.inputs[0].shape[n.attrs["axis"]])],
    "broadcast": lambda n, g, need: [_unbroadcast(g, n.inputs[0].shape)],
    "add_n": lambda n, g, need: [g] * len(n.inputs),
    "cast": lambda n, g, need: [g if n.inputs[0].dtype == float32 else None],
    "stop_gradient": lambda n, g, need: [None],
    "one_hot": lambda n, g, need: [None],
    "argmax": lambda n, g, need: [None],
    "random": lambda n, g, need: [None] * len(n.inputs),
}


def toposort(roots):
    order, seen = [], set()
    stack = [(r, False) for r in roots]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if node.id in seen:
            continue
        seen.add(node.id)
        stack.append((node, True))
        for inp in node.inputs:
            if inp.id not in seen:
                stack.append((inp, False))
    return order


def gradients(ys, xs, grad_ys=None):
    """Symbolic reverse-mode differentiation, tf.gradients semantics: d(sum ys)/d(x) for each x (None if unconnected)."""
    single = isinstance(xs, Tensor)
    ys = [ys] if isinstance(ys, Tensor) else list(ys)
    xs = [xs] if single else list(xs)
    order = toposort(ys)
    x_ids = set(x.id for x in xs)
    # forward reachability from xs
    reach = set()
    for node in order:
        if node.id in x_ids or any(i.id in reach for i in node.inputs):
            reach.add(node.id)
    pending = {}
    for i, y in enumerate(ys):
        gy = grad_ys[i] if grad_ys is not None and grad_ys[i] is not None else constant(np.ones(y.shape, np.float32))
        pending.setdefault(y.id, []).append(gy)
    result = {}
    for node in reversed(order):
        if node.id not in pending or node.id not in reach:
            continue
        g = add_n(pending.pop(node.id))
        if node.id in x_ids:
            result[node.id] = g
        if not node.inputs or node.op in ("param", "const", "placeholder"):
            continue
        need = [inp.id in reach for inp in node.inputs]
        if not any(need):
            continue
        if node.op == "aux":
            continue
        rule = GRADS.get(node.op)
        if rule is None:
            raise NotImplementedError("no gradient rule for op %r" % node.op)
        in_grads = rule(node, g, need)
        for inp, ig, nd in zip(node.inputs, in_grads, need):
            if ig is not None and nd:
                if tuple(ig.shape) != tuple(inp.shape):
                    raise AssertionError("gradient shape %s != input shape %s for %s -> %s" %
                                         (tuple(ig.shape), tuple(inp.shape), node, inp))
                pending.setdefault(inp.id, []).append(ig)
    out = [result.get(x.id) for x in xs]
    return out
