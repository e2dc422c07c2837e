"""Per-launch device time of the one-launch SyncBN kernels (gg_bn_fwd_fused_dp / gg_bn_bwd_fused_dp) against the
single-GPU kernels on the same shapes: what one in-kernel cross-GPU statistic exchange costs, without the rank skew of a
whole training step.  Run under torchrun (N ranks); rank 0 prints.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 tools/time_bn_dp.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from gg import cabi, dist as ggdist

rank, world = ggdist.init_from_env()
arena = ggdist.peer_arena()
N = 20


def graph_us(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(s.cuda_stream)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                fn(torch.cuda.current_stream().cuda_stream)
        ts = []
        for _ in range(8):
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); e1.record(s); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / N)
    return float(np.median(ts[2:]))


for R, C in ((64, 4096), (64 * 8 * 8, 128), (64 * 16 * 16, 64), (64 * 4 * 4, 256)):
    x = torch.randn(R, C, device="cuda"); y = torch.empty_like(x); dy = torch.randn(R, C, device="cuda"); dx = torch.empty_like(x)
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    mean = torch.zeros(C, device="cuda"); rstd = torch.ones(C, device="cuda"); dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    p = [t.data_ptr() for t in (x, gamma, beta, y, mean, rstd, dy, dx, dg, db)]
    single_f = graph_us(lambda st: cabi.call("gg_bn_fwd_fused", p[0], p[1], p[2], 1e-5, p[3], p[4], p[5], R, C, 1, 0.0, st))
    single_b = graph_us(lambda st: cabi.call("gg_bn_bwd_fused", p[6], p[0], p[3], p[4], p[5], p[1], p[7], p[8], p[9], R, C, 1, 0.0, st))
    line = "bn [%5d x %4d] grid %3d  single fwd %5.1f us bwd %5.1f us" % (R, C, cabi.lib.gg_bn_fused_grid(R, C), single_f, single_b)
    if world > 1:
        sf = arena.alloc(cabi.lib.gg_bn_dp_site_bytes(C, world)); sb = arena.alloc(cabi.lib.gg_bn_dp_site_bytes(C, world))
        dp_f = graph_us(lambda st: cabi.call("gg_bn_fwd_fused_dp", p[0], p[1], p[2], 1e-5, p[3], p[4], p[5], R, C, 1, 0.0,
                                             arena.peers, rank, world, sf, st))
        dp_b = graph_us(lambda st: cabi.call("gg_bn_bwd_fused_dp", p[6], p[0], p[3], p[4], p[5], p[1], p[7], p[8], p[9], R, C, 1, 0.0,
                                             arena.peers, rank, world, sb, st))
        line += "   dp%d fwd %5.1f us bwd %5.1f us" % (world, dp_f, dp_b)
    if rank == 0:
        print(line, flush=True)
if world > 1:
    torch.distributed.barrier()
    ggdist.shutdown()
