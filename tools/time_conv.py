"""Device-side timing of the conv / dense kernels at the gmgan-CIFAR shapes.

Each measurement replays a CUDA graph of N back-to-back launches (buffers pre-allocated, tensor maps encoded at capture:
no host overhead inside the timed region).  `hot` = operands stay in L2; `cold` = a 256 MiB memset between launches
(its own time, measured the same way, is subtracted)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from gg import cabi

N = 20
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def graph_time(fn, flush):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                if flush:
                    flush_buf.zero_()
                fn()
        ts = []
        spin_up(0.05)
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); e1.record(s); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / N)
    return float(np.min(ts))


def spin_up(seconds=0.6):
    """ramp the SM clocks before timing: short bursts otherwise run at idle clocks"""
    import time
    a = torch.randn(4096, 4096, device="cuda")
    t0 = time.time()
    while time.time() - t0 < seconds:
        (a @ a).sum().item()


spin_up()
cabi.call("gg_set_pdl", int(os.environ.get("GG_PDL", "0")))
cabi.call("gg_set_tc_stages", int(os.environ.get("GG_TC_STAGES", "3")))
flush_only = graph_time(lambda: None, True)


def zws(n):
    return torch.zeros(max(int(n), 256), dtype=torch.uint8, device="cuda")


def conv_case(B, H, W, Ci, Co, k=5, s=2):
    Ho, Wo = H // s, W // s
    pt = max((Ho - 1) * s + k - H, 0) // 2
    x = torch.randn(B, H, W, Ci, device="cuda"); w = torch.randn(k, k, Ci, Co, device="cuda") * .05
    b = torch.zeros(Co, device="cuda"); dy = torch.randn(B, Ho, Wo, Co, device="cuda")
    y = torch.empty(B, Ho, Wo, Co, device="cuda"); dx = torch.empty_like(x); dw = torch.empty_like(w)
    geo = (B, H, W, Ci, Co, k, s, pt, pt, Ho, Wo)
    w0, w1, w2 = (zws(cabi.lib.gg_conv2d_workspace(m, B, H, W, Ci, Co, k, s, Ho, Wo)) for m in (0, 1, 2))
    st = lambda: cabi.stream_ptr()
    fns = {
        "fwd": lambda: cabi.call("gg_conv2d_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *geo, 2, 0.2, w0.data_ptr(), w0.numel(), st()),
        "dgrad": lambda: cabi.call("gg_conv2d_dgrad", dy.data_ptr(), w.data_ptr(), None, dx.data_ptr(), *geo, 0, 0.0, w1.data_ptr(), w1.numel(), st()),
        "wgrad": lambda: cabi.call("gg_conv2d_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), *geo, w2.data_ptr(), w2.numel(), st()),
    }
    gf = 2.0 * B * Ho * Wo * Co * Ci * k * k / 1e9
    for name, fn in fns.items():
        hot = graph_time(fn, False)
        cold = graph_time(fn, True) - flush_only
        info = cabi.last_tc_info()
        print("conv %-5s B%d %dx%d %d->%d k%d s%d  hot %6.1f us (%6.1f TF/s)  cold %6.1f us  backend %d tiles %d splits %d n_tile %d stages %d" %
              (name, B, H, W, Ci, Co, k, s, hot, gf / hot * 1e3, cold, cabi.lib.gg_last_backend(), info["tiles"], info["splits"],
               info["n_tile"], info["stages"]), flush=True)


def gemm_case(M, N, K, ta=0, tb=0):
    A = torch.randn((K, M) if ta else (M, K), device="cuda"); Bm = torch.randn((N, K) if tb else (K, N), device="cuda")
    bias = torch.zeros(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    ws = zws(cabi.lib.gg_gemm_workspace(M, N, K))
    fn = lambda: cabi.call("gg_gemm", A.data_ptr(), Bm.data_ptr(), None if ta else bias.data_ptr(), C.data_ptr(), M, N, K, ta, tb, 0, 0.0,
                           ws.data_ptr(), ws.numel(), cabi.stream_ptr())
    hot = graph_time(fn, False)
    cold = graph_time(fn, True) - flush_only
    print("gemm M%d N%d K%d ta%d tb%d  hot %6.1f us  cold %6.1f us  backend %d" % (M, N, K, ta, tb, hot, cold, cabi.lib.gg_last_backend()), flush=True)


if __name__ == "__main__":
    print("flush-only %.1f us" % flush_only)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "dom":
        conv_case(128, 16, 16, 64, 128)
        sys.exit(0)
    conv_case(64, 16, 16, 64, 128)
    conv_case(64, 8, 8, 128, 256)
    if which in ("all", "batched"):
        # the shapes the step actually launches after sibling batching (fake + real towers: 128 rows)
        conv_case(128, 32, 32, 3, 64)
        conv_case(64, 32, 32, 3, 64)
        conv_case(128, 16, 16, 64, 128)
        conv_case(128, 8, 8, 128, 256)
    if which in ("all", "3x3"):
        # 3x3 stride 1 / 2 at the same channel counts: a SYNTHETIC extra for BASELINE.json's "conv3x3" wording — the
        # reference has no 3x3 convolution anywhere; filter_size is a free argument of the op
        conv_case(64, 16, 16, 64, 128, k=3, s=1)
        conv_case(64, 16, 16, 64, 128, k=3, s=2)
        conv_case(64, 8, 8, 128, 256, k=3, s=1)
    if which == "all":
        conv_case(128, 32, 32, 32, 64)
        gemm_case(64, 512, 4608); gemm_case(64, 4608, 512, 0, 1); gemm_case(4608, 512, 64, 1, 0)
        gemm_case(64, 4096, 128); gemm_case(64, 512, 512); gemm_case(64, 1, 512); gemm_case(64, 512, 158)
