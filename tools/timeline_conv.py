import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import gpu_util as U
from gg import cabi
dbg = torch.zeros(256, dtype=torch.int64, device="cuda")
cabi.call("gg_debug_set_buffer", dbg.data_ptr())
B, H, W, Ci, Co, k, s = 64, 16, 16, 64, 128, 5, 2
x = torch.randn(B, H, W, Ci, device="cuda"); w = torch.randn(k, k, Ci, Co, device="cuda") * .05
b = torch.zeros(Co, device="cuda"); dy = torch.randn(B, H // 2, W // 2, Co, device="cuda")
import time
a_ = torch.randn(4096, 4096, device='cuda')
t_ = time.time()
while time.time() - t_ < 0.6: (a_ @ a_).sum().item()
A_ = torch.randn(64, 512, device="cuda"); B_ = torch.randn(512, 512, device="cuda"); bias_ = torch.zeros(512, device="cuda")
for name, fn in (("gemm64x512x512", lambda: U.gemm(A_, B_, bias_, 64, 512, 512, 0, 0, act="leaky")), ("fwd", lambda: U.conv_fwd(x, w, b, s, 'SAME', act="leaky")), ("dgrad", lambda: U.conv_dgrad(dy, w, None, H, W, s, 'SAME')),
                 ("wgrad", lambda: U.conv_wgrad(x, dy, k, s, 'SAME'))):
    for rep in range(50):
        dbg.zero_(); dbg[201] = 2**62; fn(); torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    t0 = t[0]
    issue = [int(v - t0) for v in t[1:61] if v]
    land = [int(v - t0) for v in t[64:124] if v]
    print(name, "issue(ns):", issue[:16])
    print(name, "landed(ns):", land[:16])
    print(name, "ALL CTAs: first start %d, last end %d ; some last-CTA (tile %d) reduce start %d end %d" % (t[201] - t0, t[200] - t0, t[204], t[202] - t0, t[203] - t0))
    print(name, "start->kernel timeline (ns): first TMA issue %d, accum_ready %d, partial_written %d, staged(after rendezvous+reduce) %d, stored %d, last CTA end %d" % (issue[0] if issue else -1, t[128]-t0, (t[129]-t0) if t[129] else -1, t[140]-t0, t[203]-t0, t[200]-t0))
    print(name, "this/last CTA: staged %d barrier %d stored %d" % (t[140] - t0, t[141] - t0, t[203] - t0))
    print(name, "cycles: idx->loop %d, tmem_ld %d, bias/act/STS %d, whole staging loop %d" % (t[151] - t[150], t[152] - t[151], t[153] - t[152], t[154] - t[150]))
    print(name, "accum_ready %d partial_written %d epilogue_done %d" % (t[128] - t0, (t[129] - t0) if t[129] else -1, t[131] - t0))
