"""Where a launch of the one-launch small-channel filter gradient (gg_conv_small.cu) spends its time: per-CTA %globaltimer stamps
(gg_debug_set_small_buffer) of a few launches at the bench geometry (B=128 32x32 3->64, the batched Discriminator.1), and the
launch plan with the number of 8-CTA clusters the device keeps resident at once.   python tools/timeline_small_wgrad.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from gg import cabi

NAMES = ["entry", "tiles staged", "FMA done", "cluster rendezvous", "partial in L2", "ticket/flag", "dw written (last cluster)"]


def run(B, H, W, Ci, Co, k=5, s=2):
    Ho, Wo = H // s, W // s
    out8 = (C.c_int * 8)()
    cabi.call("gg_debug_small_wgrad_info", B, H, W, Ci, Co, k, s, Ho, Wo, out8)
    print("geometry B%d %dx%d %d->%d k%d s%d: served %d, rows/unit %d, units %d, threads %d, pixel groups %d, clusters %d, smem %d B, "
          "resident clusters (occupancy query) %d" % ((B, H, W, Ci, Co, k, s) + tuple(out8)))
    x = torch.randn(B, H, W, Ci, device="cuda")
    dy = torch.randn(B, Ho, Wo, Co, device="cuda")
    dw = torch.empty(k, k, Ci, Co, device="cuda")
    need = cabi.lib.gg_conv2d_wgrad_workspace(B, H, W, Ci, Co, k, s, Ho, Wo)
    ws = torch.zeros(max(need, 256), dtype=torch.uint8, device="cuda")
    dbg = torch.zeros(8 * 160, dtype=torch.int64, device="cuda")
    geo = (B, H, W, Ci, Co, k, s, 1, 1, Ho, Wo)
    for it in range(4):
        dbg.zero_()
        cabi.call("gg_debug_set_small_buffer", dbg.data_ptr())
        cabi.call("gg_conv2d_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), *geo, ws.data_ptr(), ws.numel(), cabi.stream_ptr())
        torch.cuda.synchronize()
        cabi.call("gg_debug_set_small_buffer", 0)
    t = dbg.cpu().numpy().reshape(160, 8)
    live = t[:, 0] > 0
    t = t[live]
    t0 = t[:, 0].min()
    print("  %d CTAs stamped; ns relative to the first CTA's entry: min / median / max over CTAs" % len(t))
    for i, n in enumerate(NAMES):
        col = t[:, i]
        col = col[col > 0] - t0
        if len(col):
            print("    %-28s %7d %7d %7d   (n=%d)" % (n, col.min(), np.median(col), col.max(), len(col)))
    ent = np.sort(t[:, 0] - t0)
    print("  CTA entry times (sorted, every 8th):", ent[::8].tolist())


if __name__ == "__main__":
    run(128, 32, 32, 3, 64)
    run(64, 32, 32, 3, 64)
