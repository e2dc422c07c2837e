"""Sibling batching: run the two applications of a weight-sharing network as ONE batched application.

Every Graphical-GAN objective evaluates its discriminators twice per step — D(fake_x, p_z) and D(real_x, q_z) share all
parameters (gmgan_inference_cifar10.py:367-370; lib.param returns the same variable on the second call,
tflib/__init__.py:22-27).  TensorFlow runs the two towers as separate op sequences.  On a B200 the towers' GEMMs are far
too small to fill the machine (bs=64: 16-32 output tiles for 148 SMs) and every launch pays ~10 us of fixed pipeline /
split-K cost, so the plan is throughput-bound by the NUMBER of tensor-core launches.  This pass matches the two towers
structurally from the logits downward and rebuilds the matched region once on the row-concatenated inputs: twice the
rows per launch, half the launches, and the weight gradients of the two towers come out of one wgrad launch already
summed (tf.gradients' add_n over the two uses of each variable disappears).  Values are unchanged — every op in the
matched region is row-wise independent; ops that couple rows (batch norm, reductions over axis 0, random draws) end the
region, and the un-matched inputs are concatenated along axis 0 there (zero-copy in the plan: gg/executor.py places the
producers' buffers inside the concat buffer).  The results are axis-0 slices (views) of the batched output.
"""
import os

from . import ops as O
from .graph import Tensor, float32

HEAVY = ("conv", "matmul")
_BOUNDARY = {}     # ids of the joined tensors -> concat node, shared by all towers of a graph: p_z ++ q_z feeds both HyperD and D


def enabled():
    return os.environ.get("GG_BATCH_SIBLINGS", "1") != "0"


class _Matcher(object):
    """structural match of N >= 2 towers (tuples `ts` of tensors, one per tower) that apply the same weights"""

    def __init__(self):
        self.memo = {}
        self.heavy = 0

    def boundary(self, ts):
        key = tuple(t.id for t in ts)
        if key not in _BOUNDARY:
            _BOUNDARY[key] = O.concat(list(ts), 0)
        return _BOUNDARY[key]

    @staticmethod
    def same_attrs(ts, skip=()):
        first = {k: v for k, v in ts[0].attrs.items() if k not in skip}
        return all({k: v for k, v in t.attrs.items() if k not in skip} == first for t in ts[1:])

    def match(self, ts):
        """tensor T whose row blocks (axis 0) are ts[0], ts[1], ... in this order"""
        key = tuple(t.id for t in ts)
        if key not in self.memo:
            t = self._match(ts)
            if t is None:
                t = self.boundary(ts)
            self.memo[key] = t
        return self.memo[key]

    def _inputs(self, ts, i):
        return tuple(t.inputs[i] for t in ts)

    def _match(self, ts):
        a = ts[0]
        if any(t is a for t in ts[1:]) or len(set(t.id for t in ts)) != len(ts):
            return None
        if any(t.op != a.op or len(t.inputs) != len(a.inputs) or t.dtype != float32 for t in ts) or a.dtype != float32:
            return None
        if len(a.shape) == 0 or any(len(t.shape) != len(a.shape) or tuple(t.shape[1:]) != tuple(a.shape[1:]) for t in ts):
            return None
        rows = [t.shape[0] for t in ts]
        total = sum(rows)
        op = a.op
        shared = lambda i: all(t.inputs[i] is a.inputs[i] for t in ts)       # the SAME variable in every tower
        if op == "conv":
            if a.attrs["mode"] not in ("fwd", "dgrad") or not self.same_attrs(ts, skip=("B",)):
                return None
            if not all(shared(i) for i in range(1, len(a.inputs))):
                return None                                        # filters / bias must be the SAME variable
            x = self.match(self._inputs(ts, 0))
            geom = {k: a.attrs[k] for k in ("H", "W", "Ci", "Co", "k", "stride", "pad_t", "pad_l", "Ho", "Wo")}
            geom["B"] = total
            t = O.conv(a.attrs["mode"], x, a.inputs[1], geom, a.inputs[2] if len(a.inputs) == 3 else None)
            t.attrs["act"], t.attrs["alpha"] = a.attrs["act"], a.attrs["alpha"]
            self.heavy += 1
            return t
        if op == "matmul":
            if a.attrs["ta"] or not self.same_attrs(ts):
                return None
            if not all(shared(i) for i in range(1, len(a.inputs))):
                return None
            x = self.match(self._inputs(ts, 0))
            t = O.matmul(x, a.inputs[1], False, a.attrs["tb"], a.inputs[2] if len(a.inputs) == 3 else None)
            t.attrs["act"], t.attrs["alpha"] = a.attrs["act"], a.attrs["alpha"]
            self.heavy += 1
            return t
        if op == "unary":
            if not self.same_attrs(ts) or any(t.inputs[0].shape[0] != r for t, r in zip(ts, rows)):
                return None
            x = self.match(self._inputs(ts, 0))
            return Tensor("unary", (x,), a.attrs, (total,) + tuple(a.shape[1:]), float32)
        if op == "binary":
            if not self.same_attrs(ts):
                return None
            ins = []
            for i in range(len(a.inputs)):
                ps = self._inputs(ts, i)
                full = all(len(p.shape) == len(t.shape) and p.shape[0] == r for p, t, r in zip(ps, ts, rows))
                same = all(p is ps[0] for p in ps)
                if full and (max(rows) > 1 or not same):
                    ins.append(self.match(ps))
                elif same and (len(ps[0].shape) < len(a.shape) or ps[0].shape[0] == 1):
                    ins.append(ps[0])                              # shared operand broadcast along the rows
                else:
                    return None
            return Tensor("binary", ins, a.attrs, (total,) + tuple(a.shape[1:]), float32)
        if op == "reshape":
            ps = self._inputs(ts, 0)
            if len(ps[0].shape) == 0 or any(p.shape[0] != r or tuple(p.shape[1:]) != tuple(ps[0].shape[1:]) for p, r in zip(ps, rows)):
                return None
            return O.reshape(self.match(ps), (total,) + tuple(a.shape[1:]))
        if op == "transpose":
            if a.attrs["perm"][0] != 0 or not self.same_attrs(ts):
                return None
            return O.transpose(self.match(self._inputs(ts, 0)), a.attrs["perm"])
        if op == "concat":
            if a.attrs["axis"] == 0 or not self.same_attrs(ts):
                return None
            if any(p.shape[0] != r for t, r in zip(ts, rows) for p in t.inputs):
                return None
            return O.concat([self.match(self._inputs(ts, i)) for i in range(len(a.inputs))], a.attrs["axis"])
        if op == "slice":
            if a.attrs["axis"] == 0 or not self.same_attrs(ts) or \
                    any(tuple(t.inputs[0].shape[1:]) != tuple(a.inputs[0].shape[1:]) for t in ts):
                return None
            return O.slice_axis(self.match(self._inputs(ts, 0)), a.attrs["axis"], a.attrs["start"], a.attrs["size"])
        if op == "softmax":
            if len(a.shape) < 2:
                return None
            return O.softmax(self.match(self._inputs(ts, 0)))
        if op == "reduce":
            if 0 in a.attrs["axes"] or not self.same_attrs(ts) or \
                    any(tuple(t.inputs[0].shape[1:]) != tuple(a.inputs[0].shape[1:]) for t in ts):
                return None
            x = self.match(self._inputs(ts, 0))
            return Tensor("reduce", (x,), a.attrs, (total,) + tuple(a.shape[1:]), float32)
        return None                                                # bn, random, one_hot, ... : rows are coupled or not worth it


def batch_group(ts):
    """tensors computing the same function with the same weights on different inputs -> the same values as row slices of ONE
    batched application, or `ts` unchanged when they share no conv / dense layer with common weights"""
    ts = list(ts)
    if not enabled() or len(ts) < 2 or any(not isinstance(t, Tensor) for t in ts):
        return ts
    if len(ts[0].shape) == 0 or any(len(t.shape) != len(ts[0].shape) for t in ts):
        return ts
    m = _Matcher()
    t = m._match(tuple(ts))
    if t is None or m.heavy == 0:
        return ts
    out, off = [], 0
    for x in ts:
        out.append(O.slice_axis(t, 0, off, x.shape[0]))
        off += x.shape[0]
    return out


def batch_pair(a, b):
    """(a, b) -> (a', b') computing the same values from ONE batched application of the shared network, or (a, b) unchanged
    when the two graphs share no conv / dense layer with common weights."""
    if not isinstance(a, Tensor) or not isinstance(b, Tensor):
        return a, b
    out = batch_group([a, b])
    return out[0], out[1]


def _signature(t, depth=0):
    """structural key of a tower (ops, attributes and the identity of its weights): towers with equal keys can be batched"""
    if t.op in ("param", "const", "placeholder", "random") or depth > 64:
        return (t.op, t.id) if t.op == "param" else (t.op, tuple(t.shape[1:]))
    if t.op in HEAVY:
        return (t.op, tuple(sorted((k, str(v)) for k, v in t.attrs.items() if k != "B")), tuple(i.id for i in t.inputs[1:]),
                _signature(t.inputs[0], depth + 1))
    if t.op in ("unary", "reshape", "transpose", "softmax") and t.inputs:
        return (t.op, tuple(t.shape[1:]), _signature(t.inputs[0], depth + 1))
    return (t.op, tuple(t.shape[1:]))


def batch_pairs(fakes, reals):
    """list form (local_ep / weighted_local_epce take lists of logit tensors).  Towers are grouped over the WHOLE list: the
    LEN-1 pairwise latent discriminators of the SSGAN scripts (ssgan_inference_moving_mnist.py:531-534) share one set of
    weights, so their 2(LEN-1) applications (fake and real) run as ONE batched tower instead of LEN-1 pairs."""
    if not isinstance(fakes, (list, tuple)):
        return batch_pair(fakes, reals)
    fakes, reals = list(fakes), list(reals)
    if os.environ.get("GG_BATCH_GROUPS", "1") == "0" or not all(isinstance(t, Tensor) for t in fakes + reals):
        out = [batch_pair(f, r) for f, r in zip(fakes, reals)]
        return [o[0] for o in out], [o[1] for o in out]
    groups = {}
    for i, (f, r) in enumerate(zip(fakes, reals)):
        groups.setdefault((_signature(f), _signature(r)), []).append(i)
    new_f, new_r = list(fakes), list(reals)
    for (sf, sr), idx in groups.items():
        members = [fakes[i] for i in idx] + [reals[i] for i in idx] if sf == sr else None
        if members is not None and len(members) > 2:
            out = batch_group(members)
            if out is not members and len(out) == len(members) and any(o is not m for o, m in zip(out, members)):
                for k, i in enumerate(idx):
                    new_f[i], new_r[i] = out[k], out[len(idx) + k]
                continue
        for i in idx:
            new_f[i], new_r[i] = batch_pair(fakes[i], reals[i])
    return new_f, new_r
