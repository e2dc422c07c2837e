"""Do two independent tensor-core conv launches on two streams of one CUDA graph overlap?

Captures (a) one chain of N launches, (b) two independent chains of N launches on forked streams, (c) a chain of N
launches next to a chain of N small element-wise kernels, and reports time per graph replay.  Perfect overlap: (b) == (a);
no overlap: (b) == 2 x (a).  Run under the split-K variants (default cooperative rendezvous, GG_TC_CLUSTER=1,
GG_TC_SPLITS=1) to see which launch kinds the hardware co-schedules.

    python tools/exp_concurrency.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from gg import cabi

N = 16


def spin_up(seconds=0.5):
    import time
    a = torch.randn(4096, 4096, device="cuda")
    t0 = time.time()
    while time.time() - t0 < seconds:
        (a @ a).sum().item()


def zws(n):
    return torch.zeros(max(int(n), 256), dtype=torch.uint8, device="cuda")


def conv_launcher(B, H, W, Ci, Co, k=5, s=2, mode="fwd"):
    Ho, Wo = H // s, W // s
    pt = max((Ho - 1) * s + k - H, 0) // 2
    x = torch.randn(B, H, W, Ci, device="cuda"); w = torch.randn(k, k, Ci, Co, device="cuda") * .05
    b = torch.zeros(Co, device="cuda"); dy = torch.randn(B, Ho, Wo, Co, device="cuda")
    y = torch.empty(B, Ho, Wo, Co, device="cuda"); dx = torch.empty_like(x); dw = torch.empty_like(w)
    geo = (B, H, W, Ci, Co, k, s, pt, pt, Ho, Wo)
    m = {"fwd": 0, "dgrad": 1, "wgrad": 2}[mode]
    ws = zws(cabi.lib.gg_conv2d_workspace(m, B, H, W, Ci, Co, k, s, Ho, Wo))
    keep = (x, w, b, dy, y, dx, dw, ws)
    if mode == "fwd":
        return lambda st: cabi.call("gg_conv2d_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *geo, 2, 0.2, ws.data_ptr(), ws.numel(), st), keep
    if mode == "dgrad":
        return lambda st: cabi.call("gg_conv2d_dgrad", dy.data_ptr(), w.data_ptr(), None, dx.data_ptr(), *geo, 0, 0.0, ws.data_ptr(), ws.numel(), st), keep
    return lambda st: cabi.call("gg_conv2d_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), *geo, ws.data_ptr(), ws.numel(), st), keep


def small_launcher(n=64 * 512):
    a = torch.randn(n, device="cuda"); o = torch.empty_like(a)
    return lambda st: cabi.call("gg_unary", cabi.UNARY["leaky"], a.data_ptr(), o.data_ptr(), n, 0.2, 0.0, st), (a, o)


def time_graph(chains):
    """chains: list of lists of launch closures; chain 0 runs on the capture stream, the others on forked streams"""
    main = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in chains[1:]]
    with torch.cuda.stream(main):
        for ch in chains:
            for f in ch[:1]:
                f(main.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            cs = torch.cuda.current_stream()
            for sd in side:
                sd.wait_stream(cs)
            for st, ch in zip([cs] + side, chains):
                for f in ch:
                    f(st.cuda_stream)
            for sd in side:
                cs.wait_stream(sd)
        ts = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main); g.replay(); e1.record(main); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.min(ts))


def main():
    cabi.call("gg_set_tc_max_ctas", int(os.environ.get("GG_TC_MAX_CTAS", "74")))
    spin_up()
    tag = "coop=%s cluster=%s splits=%s max_ctas=%s" % (os.environ.get("GG_TC_COOP", "-"), os.environ.get("GG_TC_CLUSTER", "0"),
                                                       os.environ.get("GG_TC_SPLITS", "auto"), os.environ.get("GG_TC_MAX_CTAS", "74"))
    for label, geo in (("E.2 fwd", (64, 16, 16, 64, 128)), ("E.3 fwd", (64, 8, 8, 128, 256))):
        fa, ka = conv_launcher(*geo)
        fb, kb = conv_launcher(*geo)
        fs, ks = small_launcher()
        one = time_graph([[fa] * N])
        two = time_graph([[fa] * N, [fb] * N])
        mix = time_graph([[fa] * N, [fs] * N])
        small = time_graph([[fs] * N])
        print("[%s] %s: 1 chain %.1f us/launch | 2 chains %.1f us/pair (overlap %.0f%%) | conv+small chain %.1f us/pair (small alone %.1f)" %
              (tag, label, one / N, two / N, 100.0 * (2 * one - two) / one, mix / N, small / N), flush=True)


if __name__ == "__main__":
    main()
