"""TEST INFRASTRUCTURE — a CPU interpreter of the host-side graph IR (gg/graph.py nodes), float64, PyTorch-CPU.

It evaluates a fetch set node by node with the same per-op semantics the C-ABI kernels implement (oracle/tf_ops.py for
conv / batch norm), so graph-level logic — shape inference, build-time fusion, the symbolic gradient rules of gg/ops.py,
the sibling-batching rewrite — can be checked without a GPU: against torch autograd, and rewritten-vs-original graph.
Nothing under graphical-gan_b200/ imports this file; the product path has no CPU execution.
"""
import numpy as np
import torch

from gg.graph import Tensor, Operation, float32, prod
from oracle import tf_ops as O

DT = torch.float64


def _act(v, act, alpha):
    if act in (None, "none"):
        return v
    if act == "relu":
        return torch.clamp(v, min=0)
    if act == "leaky":
        return torch.maximum(alpha * v, v)
    if act == "tanh":
        return torch.tanh(v)
    if act == "sigmoid":
        return torch.sigmoid(v)
    raise NotImplementedError(act)


def _act_grad_from_out(y, g, act, alpha):
    if act in (None, "none"):
        return g
    if act == "relu":
        return torch.where(y > 0, g, torch.zeros_like(g))
    if act == "leaky":
        return torch.where(y > 0, g, alpha * g)
    if act == "tanh":
        return (1 - y * y) * g
    if act == "sigmoid":
        return y * (1 - y) * g
    raise NotImplementedError(act)


UNARY = {
    "copy": lambda x, a, b: x, "relu": lambda x, a, b: torch.clamp(x, min=0), "leaky": lambda x, a, b: torch.maximum(a * x, x),
    "tanh": lambda x, a, b: torch.tanh(x), "sigmoid": lambda x, a, b: torch.sigmoid(x), "exp": lambda x, a, b: torch.exp(x),
    "log": lambda x, a, b: torch.log(x), "sqrt": lambda x, a, b: torch.sqrt(x), "square": lambda x, a, b: x * x,
    "neg": lambda x, a, b: -x, "abs": lambda x, a, b: x.abs(), "affine": lambda x, a, b: a * x + b,
    "pow": lambda x, a, b: torch.pow(x, a), "rsqrt": lambda x, a, b: torch.rsqrt(x), "recip": lambda x, a, b: 1.0 / x,
    "bce": lambda x, a, b: torch.clamp(x, min=0) - x * a + torch.log1p(torch.exp(-x.abs())),
    "clip": lambda x, a, b: torch.clamp(x, a, b), "sign": lambda x, a, b: torch.sign(x),
    "softsign": lambda x, a, b: x / (1 + x.abs()), "divc": lambda x, a, b: x / a, "rdivc": lambda x, a, b: a / x,
}
BINARY = {
    "add": lambda a, b, al: a + b, "sub": lambda a, b, al: a - b, "mul": lambda a, b, al: a * b, "div": lambda a, b, al: a / b,
    "max": lambda a, b, al: torch.maximum(a, b), "min": lambda a, b, al: torch.minimum(a, b),
    "relu_grad": lambda a, b, al: torch.where(a > 0, b, torch.zeros_like(b) + 0 * a),
    "leaky_grad": lambda a, b, al: torch.where(a > 0, b + 0 * a, al * b + 0 * a),
    "tanh_grad": lambda a, b, al: (1 - a * a) * b, "sigmoid_grad": lambda a, b, al: a * (1 - a) * b,
    "bce_grad": lambda a, b, al: (torch.sigmoid(a) - al) * b, "ge_mask": lambda a, b, al: (a >= b).to(DT),
    "gt_mask": lambda a, b, al: (a > b).to(DT), "abs_grad": lambda a, b, al: torch.sign(a) * b,
    "pow": lambda a, b, al: torch.pow(a, b),
}


class Interp(object):
    def __init__(self, feeds=None, params=None, seed=0):
        """feeds: {Tensor: array} (any node may be fed, like TF); params: {param node id or name: array} overrides"""
        self.val = {}
        self.aux = {}
        self.params = params or {}
        self.gen = torch.Generator().manual_seed(seed)
        for t, v in (feeds or {}).items():
            self.val[t.id] = torch.as_tensor(np.asarray(v)).to(DT).reshape(tuple(t.shape))

    def run(self, fetches):
        single = isinstance(fetches, Tensor)
        fl = [fetches] if single else list(fetches)
        out = [self.eval(f).numpy() for f in fl]
        return out[0] if single else out

    def eval(self, root):
        stack = [(root, False)]
        while stack:
            node, done = stack.pop()
            if node.id in self.val:
                continue
            if done:
                self.val[node.id] = self._compute(node)
                assert tuple(self.val[node.id].shape) == tuple(node.shape), (node, tuple(self.val[node.id].shape))
                continue
            stack.append((node, True))
            for inp in node.inputs:
                if inp.id not in self.val:
                    stack.append((inp, False))
        return self.val[root.id]

    # ------------------------------------------------------------------------------------------
    def _compute(self, n):
        a = n.attrs
        I = [self.val[i.id] for i in n.inputs]
        op = n.op
        if op == "const":
            return torch.as_tensor(np.asarray(a["value"])).to(DT).reshape(tuple(n.shape))
        if op == "param":
            v = self.params.get(n.id, self.params.get(n.name, a["init"]))
            return torch.as_tensor(np.asarray(v)).to(DT).reshape(tuple(n.shape))
        if op == "placeholder":
            raise KeyError("placeholder %s must be fed" % n.name)
        if op == "random":
            if a["kind"] == "normal":
                return torch.randn(tuple(n.shape), generator=self.gen, dtype=DT) * a["b"] + a["a"]
            if a["kind"] == "uniform":
                return torch.rand(tuple(n.shape), generator=self.gen, dtype=DT) * (a["b"] - a["a"]) + a["a"]
            p = I[0] / I[0].sum()
            return torch.multinomial(p, n.shape[0], replacement=True, generator=self.gen).to(DT)
        if op in ("reshape", "stop_gradient"):
            return I[0].reshape(tuple(n.shape))
        if op == "aux":
            return self.aux[(n.inputs[0].id, a["k"])].reshape(tuple(n.shape))
        if op == "unary":
            return UNARY[a["fn"]](I[0], a["a"], a["b"])
        if op == "binary":
            return BINARY[a["fn"]](I[0], I[1], a["alpha"]).expand(tuple(n.shape)).clone()
        if op == "broadcast":
            return I[0].expand(tuple(n.shape)).clone()
        if op == "add_n":
            out = I[0].clone()
            for t in I[1:]:
                out = out + t
            return out
        if op == "cast":
            return torch.trunc(I[0]) if n.dtype != float32 else I[0]
        if op == "reduce":
            axes = list(a["axes"])
            if a["fn"] == "sum":
                return I[0].sum(dim=axes, keepdim=True)
            if a["fn"] == "mean":
                return I[0].mean(dim=axes, keepdim=True)
            return I[0].amax(dim=axes, keepdim=True)
        if op == "softmax":
            return torch.softmax(I[0], dim=-1)
        if op == "softmax_grad":
            y, g = I
            return y * (g - (g * y).sum(dim=-1, keepdim=True))
        if op == "argmax":
            return I[0].argmax(dim=-1).to(DT)
        if op == "one_hot":
            return torch.nn.functional.one_hot(I[0].long(), a["depth"]).to(DT)
        if op == "transpose":
            return I[0].permute(*a["perm"]).contiguous()
        if op == "concat":
            return torch.cat(I, dim=a["axis"])
        if op == "slice":
            return I[0].narrow(a["axis"], a["start"], a["size"]).contiguous()
        if op == "pad":
            out = torch.zeros(tuple(n.shape), dtype=DT)
            out.narrow(a["axis"], a["start"], n.inputs[0].shape[a["axis"]]).copy_(I[0])
            return out
        if op == "tile":
            return I[0].repeat(*a["multiples"])
        if op == "matmul":
            A = I[0].t() if a["ta"] else I[0]
            Bm = I[1].t() if a["tb"] else I[1]
            out = A @ Bm
            if len(I) - (1 if a.get("mask_act") else 0) == 3:
                out = out + I[2].reshape(1, -1)
            out = _act(out, a["act"], a["alpha"])
            if a.get("mask_act"):
                out = _act_grad_from_out(I[-1], out, a["mask_act"], a["mask_alpha"])
            return out
        if op == "conv":
            return self._conv(n, I)
        if op == "bn":
            x, gamma, beta = I
            C = x.shape[-1]
            x2 = x.reshape(-1, C)
            mean = x2.mean(0)
            var = x2.var(0, unbiased=False)
            rstd = torch.rsqrt(var + a["eps"])
            self.aux[(n.id, 1)], self.aux[(n.id, 2)] = mean, rstd
            y = (x2 - mean) * rstd * gamma.reshape(-1) + beta.reshape(-1)
            return _act(y, a["act"], a["alpha"]).reshape(tuple(n.shape))
        if op == "bn_grad":
            gy, x, y, mean, rstd, gamma = I
            C = x.shape[-1]
            g2 = _act_grad_from_out(y.reshape(-1, C), gy.reshape(-1, C), a["act"], a["alpha"])
            xh = (x.reshape(-1, C) - mean.reshape(-1)) * rstd.reshape(-1)
            dbeta, dgamma = g2.sum(0), (g2 * xh).sum(0)
            R = g2.shape[0]
            dx = gamma.reshape(-1) * rstd.reshape(-1) * (g2 - dbeta / R - xh * dgamma / R)
            self.aux[(n.id, 1)], self.aux[(n.id, 2)] = dgamma, dbeta
            return dx.reshape(tuple(n.shape))
        raise NotImplementedError("graph_interp: op %r" % op)

    def _conv(self, n, I):
        a = n.attrs
        k, s = a["k"], a["stride"]
        pads = self._pads(a)
        mode = a["mode"]
        nchw = lambda t: t.permute(0, 3, 1, 2)
        nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
        F = torch.nn.functional
        if mode == "fwd":
            x, w = I[0], I[1]
            y = F.conv2d(F.pad(nchw(x), pads), w.permute(3, 2, 0, 1), stride=s)
            if len(I) - (1 if a.get("mask_act") else 0) == 3:
                y = y + I[2].reshape(1, -1, 1, 1)
            out = _act(nhwc(y), a["act"], a["alpha"])
        elif mode == "dgrad":
            dy, w = I[0], I[1]
            xz = torch.zeros(a["B"], a["Ci"], a["H"], a["W"], dtype=DT, requires_grad=True)
            yy = F.conv2d(F.pad(xz, pads), w.permute(3, 2, 0, 1), stride=s)
            dx, = torch.autograd.grad(yy, xz, nchw(dy))
            if len(I) - (1 if a.get("mask_act") else 0) == 3:
                dx = dx + I[2].reshape(1, -1, 1, 1)
            out = _act(nhwc(dx.detach()), a["act"], a["alpha"])
        else:
            x, dy = I[0], I[1]
            wz = torch.zeros(k, k, a["Ci"], a["Co"], dtype=DT, requires_grad=True)
            yy = F.conv2d(F.pad(nchw(x), pads), wz.permute(3, 2, 0, 1), stride=s)
            dw, = torch.autograd.grad(yy, wz, nchw(dy))
            return dw.detach()
        if a.get("mask_act"):
            out = _act_grad_from_out(I[-1], out, a["mask_act"], a["mask_alpha"])
        return out

    @staticmethod
    def _pads(a):
        k, s = a["k"], a["stride"]
        ph = max((a["Ho"] - 1) * s + k - a["H"], 0)
        pw = max((a["Wo"] - 1) * s + k - a["W"], 0)
        return (a["pad_l"], pw - a["pad_l"], a["pad_t"], ph - a["pad_t"])
