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
_BOUNDARY = {}     # (a.id, b.id) -> concat node, shared by all towers of a graph: p_z ++ q_z feeds both HyperD and D


def enabled():
    return os.environ.get("GG_BATCH_SIBLINGS", "1") != "0"


class _Matcher(object):
    def __init__(self):
        self.memo = {}
        self.heavy = 0

    def boundary(self, a, b):
        key = (a.id, b.id)
        if key not in _BOUNDARY:
            _BOUNDARY[key] = O.concat([a, b], 0)
        return _BOUNDARY[key]

    def same_attrs(self, a, b, skip=()):
        ka = {k: v for k, v in a.attrs.items() if k not in skip}
        kb = {k: v for k, v in b.attrs.items() if k not in skip}
        return ka == kb

    def match(self, a, b):
        """tensor T with T[:Ba] == a and T[Ba:] == b (rows = axis 0)"""
        key = (a.id, b.id)
        if key not in self.memo:
            t = self._match(a, b)
            if t is None:
                t = self.boundary(a, b)
            self.memo[key] = t
        return self.memo[key]

    def _batched_inputs(self, a, b, idx):
        return [self.match(a.inputs[i], b.inputs[i]) for i in idx]

    def _match(self, a, b):
        if a is b or a.op != b.op or len(a.inputs) != len(b.inputs) or a.dtype != float32 or b.dtype != float32:
            return None
        if len(a.shape) == 0 or len(a.shape) != len(b.shape) or tuple(a.shape[1:]) != tuple(b.shape[1:]):
            return None
        Ba, Bb = a.shape[0], b.shape[0]
        op = a.op
        if op == "conv":
            if a.attrs["mode"] not in ("fwd", "dgrad") or not self.same_attrs(a, b, skip=("B",)):
                return None
            if any(a.inputs[i] is not b.inputs[i] for i in range(1, len(a.inputs))):
                return None                                        # filters / bias must be the SAME variable
            x = self.match(a.inputs[0], b.inputs[0])
            geom = {k: a.attrs[k] for k in ("H", "W", "Ci", "Co", "k", "stride", "pad_t", "pad_l", "Ho", "Wo")}
            geom["B"] = Ba + Bb
            t = O.conv(a.attrs["mode"], x, a.inputs[1], geom, a.inputs[2] if len(a.inputs) == 3 else None)
            t.attrs["act"], t.attrs["alpha"] = a.attrs["act"], a.attrs["alpha"]
            self.heavy += 1
            return t
        if op == "matmul":
            if a.attrs["ta"] or not self.same_attrs(a, b):
                return None
            if any(a.inputs[i] is not b.inputs[i] for i in range(1, len(a.inputs))):
                return None
            x = self.match(a.inputs[0], b.inputs[0])
            t = O.matmul(x, a.inputs[1], False, a.attrs["tb"], a.inputs[2] if len(a.inputs) == 3 else None)
            t.attrs["act"], t.attrs["alpha"] = a.attrs["act"], a.attrs["alpha"]
            self.heavy += 1
            return t
        if op == "unary":
            if not self.same_attrs(a, b) or a.inputs[0].shape[0] != Ba or b.inputs[0].shape[0] != Bb:
                return None
            x = self.match(a.inputs[0], b.inputs[0])
            return Tensor("unary", (x,), a.attrs, (Ba + Bb,) + tuple(a.shape[1:]), float32)
        if op == "binary":
            if not self.same_attrs(a, b):
                return None
            ins = []
            for p, q in zip(a.inputs, b.inputs):
                full_p = len(p.shape) == len(a.shape) and p.shape[0] == Ba
                full_q = len(q.shape) == len(b.shape) and q.shape[0] == Bb
                if full_p and full_q and (Ba > 1 or Bb > 1 or p is not q):
                    ins.append(self.match(p, q))
                elif p is q and (len(p.shape) < len(a.shape) or p.shape[0] == 1):
                    ins.append(p)                                  # shared operand broadcast along the rows
                else:
                    return None
            return Tensor("binary", ins, a.attrs, (Ba + Bb,) + tuple(a.shape[1:]), float32)
        if op == "reshape":
            pa, pb = a.inputs[0], b.inputs[0]
            if len(pa.shape) == 0 or pa.shape[0] != Ba or pb.shape[0] != Bb or tuple(pa.shape[1:]) != tuple(pb.shape[1:]):
                return None
            return O.reshape(self.match(pa, pb), (Ba + Bb,) + tuple(a.shape[1:]))
        if op == "transpose":
            if a.attrs["perm"][0] != 0 or not self.same_attrs(a, b):
                return None
            return O.transpose(self.match(a.inputs[0], b.inputs[0]), a.attrs["perm"])
        if op == "concat":
            if a.attrs["axis"] == 0 or not self.same_attrs(a, b):
                return None
            if any(p.shape[0] != Ba for p in a.inputs) or any(q.shape[0] != Bb for q in b.inputs):
                return None
            return O.concat(self._batched_inputs(a, b, range(len(a.inputs))), a.attrs["axis"])
        if op == "slice":
            if a.attrs["axis"] == 0 or not self.same_attrs(a, b) or tuple(a.inputs[0].shape[1:]) != tuple(b.inputs[0].shape[1:]):
                return None
            return O.slice_axis(self.match(a.inputs[0], b.inputs[0]), a.attrs["axis"], a.attrs["start"], a.attrs["size"])
        if op == "softmax":
            if len(a.shape) < 2:
                return None
            return O.softmax(self.match(a.inputs[0], b.inputs[0]))
        if op == "reduce":
            if 0 in a.attrs["axes"] or not self.same_attrs(a, b) or tuple(a.inputs[0].shape[1:]) != tuple(b.inputs[0].shape[1:]):
                return None
            x = self.match(a.inputs[0], b.inputs[0])
            return Tensor("reduce", (x,), a.attrs, (Ba + Bb,) + tuple(a.shape[1:]), float32)
        return None                                                # bn, random, one_hot, ... : rows are coupled or not worth it


def batch_pair(a, b):
    """(a, b) -> (a', b') computing the same values from ONE batched application of the shared network, or (a, b) unchanged
    when the two graphs share no conv / dense layer with common weights."""
    if not enabled() or not isinstance(a, Tensor) or not isinstance(b, Tensor):
        return a, b
    if len(a.shape) == 0 or len(a.shape) != len(b.shape):
        return a, b
    m = _Matcher()
    t = m._match(a, b)
    if t is None or m.heavy == 0:
        return a, b
    Ba, Bb = a.shape[0], b.shape[0]
    return O.slice_axis(t, 0, 0, Ba), O.slice_axis(t, 0, Ba, Bb)


def batch_pairs(fakes, reals):
    """list form (local_ep / weighted_local_epce take lists of logit tensors)"""
    if isinstance(fakes, (list, tuple)):
        out = [batch_pair(f, r) for f, r in zip(fakes, reals)]
        return [o[0] for o in out], [o[1] for o in out]
    return batch_pair(fakes, reals)
