"""Device-side timing (CUDA graph of N back-to-back launches) of the non-GEMM kernels on the step's critical path, and the
graph's kernel-node dispatch rate with 1..8 parallel branches.   python tools/time_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from gg import cabi

N = 20


def spin_up(seconds=0.5):
    import time
    a = torch.randn(4096, 4096, device="cuda")
    t0 = time.time()
    while time.time() - t0 < seconds:
        (a @ a).sum().item()


def graph_time(chains, reps=10):
    main = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in chains[1:]]
    with torch.cuda.stream(main):
        for ch in chains:
            ch[0](main.cuda_stream)
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
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main); g.replay(); e1.record(main); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.min(ts))


def bn_case(R, C, act="relu"):
    x = torch.randn(R, C, device="cuda"); y = torch.empty_like(x); gy = torch.randn(R, C, device="cuda"); dx = torch.empty_like(x)
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    mean = torch.empty(C, device="cuda"); rstd = torch.empty(C, device="cuda"); dg = torch.empty(C, device="cuda"); db = torch.empty(C, device="cuda")
    S = cabi.lib.gg_bn_slices(R, C)
    part = torch.empty(S, 2, C, device="cuda"); dgb = torch.empty(2, C, device="cuda")
    a = cabi.ACT[act]
    p = lambda t: t.data_ptr()

    def old_fwd(st):
        cabi.call("gg_bn_stats", p(x), p(part), R, C, st)
        cabi.call("gg_bn_apply", p(x), p(part), S, float(R), p(gamma), p(beta), 1e-5, p(y), p(mean), p(rstd), R, C, a, 0.2, st)

    def new_fwd(st):
        cabi.call("gg_bn_fwd_fused", p(x), p(gamma), p(beta), 1e-5, p(y), p(mean), p(rstd), R, C, a, 0.2, st)

    def old_bwd(st):
        cabi.call("gg_bn_bwd_reduce", p(gy), p(x), p(y), p(mean), p(rstd), p(gamma), None, p(part), R, C, a, 0.2, st)
        cabi.call("gg_bn_fold_partials", p(part), S, p(dgb), C, st)
        cabi.call("gg_bn_bwd_apply", p(gy), p(x), p(y), p(mean), p(rstd), p(gamma), None, p(dgb), 1, float(R), p(dx), None, None, R, C, a, 0.2, st)

    def new_bwd(st):
        cabi.call("gg_bn_bwd_fused", p(gy), p(x), p(y), p(mean), p(rstd), p(gamma), p(dx), p(dg), p(db), R, C, a, 0.2, st)
    r = [graph_time([[f] * N]) / N for f in (old_fwd, new_fwd, old_bwd, new_bwd)]
    print("bn R=%-6d C=%-5d  fwd: 2-kernel %5.1f us, fused %5.1f us | bwd: 3-kernel %5.1f us, fused %5.1f us" % (R, C, *r), flush=True)


def elementwise_cases():
    for n in (64 * 512, 64 * 4096, 64 * 16 * 16 * 64):
        a = torch.randn(n, device="cuda"); b = torch.randn(n, device="cuda"); o = torch.empty_like(a)
        dims, sa, sb = cabi.int4([1, 1, 1, n]), cabi.int4([0, 0, 0, 1]), cabi.int4([0, 0, 0, 1])
        un = lambda st: cabi.call("gg_unary", cabi.UNARY["leaky"], a.data_ptr(), o.data_ptr(), n, 0.2, 0.0, st)
        bi = lambda st: cabi.call("gg_binary", cabi.BINARY["leaky_grad"], a.data_ptr(), b.data_ptr(), o.data_ptr(), dims, sa, sb, 0.2, st)
        print("elementwise n=%-8d unary %5.1f us  binary %5.1f us" % (n, graph_time([[un] * N]) / N, graph_time([[bi] * N]) / N), flush=True)
    x = torch.randn(64, 4, 4, 256, device="cuda"); y = torch.empty_like(x)
    tr = lambda st: cabi.call("gg_transpose_b2d", x.data_ptr(), y.data_ptr(), 64, 16, 256, st)
    print("transpose_b2d 64x16x256  %5.1f us" % (graph_time([[tr] * N]) / N), flush=True)


def dispatch_rate():
    n = 64 * 128
    bufs = [(torch.randn(n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(12)]
    mk = lambda a, o: (lambda st: cabi.call("gg_unary", cabi.UNARY["leaky"], a.data_ptr(), o.data_ptr(), n, 0.2, 0.0, st))
    M = 40
    for S in (1, 2, 4, 6, 8, 12):
        chains = [[mk(*bufs[s])] * M for s in range(S)]
        t = graph_time(chains)
        print("dispatch: %2d parallel chains x %d tiny kernels: %7.1f us total = %5.2f us per chain step, %5.2f us per kernel node" %
              (S, M, t, t / M, t / (M * S)), flush=True)


if __name__ == "__main__":
    spin_up()
    dispatch_rate()
    for R, C in ((64, 4096), (4096, 128), (16384, 64), (1024, 256)):
        bn_case(R, C)
    elementwise_cases()
