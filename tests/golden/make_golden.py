"""Writes the golden vectors under tests/golden/ from the CPU oracle (fp64 compute, stored fp32).

The reference ships no fixtures and cannot run here (no TensorFlow, Python 2 sources), so these vectors pin THIS
repository's oracle: CPU tests check the oracle still reproduces them, GPU tests check the CUDA path against them.
    python -m tests.golden.make_golden          # regenerates every .npz (deterministic seeds)
"""
import os

import numpy as np
import torch

from oracle import tf_ops as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


def compute(kind, d):
    """recompute the outputs of a golden case from its stored inputs"""
    if kind == "conv":
        x, w, b = _t(d["x"]).requires_grad_(True), _t(d["w"]).requires_grad_(True), _t(d["b"])
        stride, padding = int(d["stride"]), str(d["padding"])
        y = O.conv2d(x, w, stride, padding, b)
        dx, dw = torch.autograd.grad(y, (x, w), _t(d["gy"]))
        return {"y": y.detach().numpy(), "dx": dx.numpy(), "dw": dw.numpy()}
    if kind == "deconv":
        x, w, b = _t(d["x"]), _t(d["w"]), _t(d["b"])
        return {"y": O.conv2d_transpose(x, w, 2, 'SAME', b).numpy()}
    if kind == "bn":
        x, sc, of = _t(d["x"]).requires_grad_(True), _t(d["scale"]).requires_grad_(True), _t(d["offset"]).requires_grad_(True)
        y = O.batchnorm(x, sc, of, [int(a) for a in d["axes"]])
        dx, ds, do = torch.autograd.grad(y, (x, sc, of), _t(d["gy"]))
        return {"y": y.detach().numpy(), "dx": dx.numpy(), "dscale": ds.numpy(), "doffset": do.numpy()}
    if kind == "adam":
        p = _t(d["p"]).clone()
        opt = O.TFAdam([p], lr=float(d["lr"]), beta1=float(d["beta1"]), beta2=float(d["beta2"]))
        for g in d["grads"]:
            opt.step([_t(g)])
        return {"p": p.numpy()}
    if kind == "losses":
        df, dr = _t(d["disc_fake"]), _t(d["disc_real"])
        gen, disc = O.local_ep_costs([df[0], df[1]], [dr[0], dr[1]])
        return {"gen": np.asarray(float(gen)), "disc": np.asarray(float(disc)),
                "l2": np.asarray(float(O.distance(_t(d["a"]), _t(d["b2"]), 'l2'))),
                "gp": np.asarray(float(O.gradient_penalty(_t(d["a"]) * 0.05, 10.0)))}
    raise ValueError(kind)


def _save(name, kind, inputs):
    out = compute(kind, inputs)
    payload = {k: (np.asarray(v, dtype=np.float32) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in inputs.items()}
    payload.update({"out_" + k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
    payload["kind"] = kind
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)


def main():
    rs = np.random.RandomState(2026)
    f32 = lambda *s: rs.randn(*s).astype(np.float32)
    # conv layers of the hot path at batch 2 (Extractor/Discriminator .1/.2/.3, MNIST 7->4 with (2,2) padding)
    # (channel counts reduced to keep the fixtures small; still multiples of 32 so the tcgen05 path is the one exercised)
    for name, (B, H, Ci, Co) in {"conv_e1": (2, 32, 3, 32), "conv_e2": (2, 16, 32, 64), "conv_e3": (2, 8, 64, 64),
                                 "conv_mnist3": (2, 7, 32, 32)}.items():
        Ho = -(-H // 2)
        _save(name, "conv", dict(x=f32(B, Ci, H, H), w=f32(5, 5, Ci, Co) * 0.05, b=f32(Co) * 0.1, gy=f32(B, Co, Ho, Ho),
                                 stride=2, padding="SAME"))
    for name, (B, H, Cin, Cout) in {"deconv_g2": (8, 4, 64, 32), "deconv_g3": (2, 8, 64, 32), "deconv_g5": (2, 16, 32, 3)}.items():
        _save(name, "deconv", dict(x=f32(B, Cin, H, H), w=f32(5, 5, Cout, Cin) * 0.03, b=f32(Cout) * 0.1))
    _save("bn_spatial", "bn", dict(x=f32(8, 16, 4, 4) * 1.5 + 0.2, scale=(rs.rand(16) + .5).astype(np.float32), offset=f32(16),
                                   gy=f32(8, 16, 4, 4), axes=np.array([0, 2, 3])))
    _save("bn_dense", "bn", dict(x=f32(16, 64), scale=(rs.rand(1, 64) + .5).astype(np.float32), offset=f32(1, 64), gy=f32(16, 64),
                                 axes=np.array([0])))
    _save("adam_5steps", "adam", dict(p=f32(300), grads=np.stack([f32(300) * 10.0 ** (i - 2) for i in range(5)]), lr=2e-4, beta1=.5,
                                      beta2=.999))
    _save("losses", "losses", dict(disc_fake=f32(2, 64) * 2, disc_real=f32(2, 64) * 2, a=f32(64, 128), b2=f32(64, 128)))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
