"""In-kernel timeline of the tensor-core conv / dense kernel (globaltimer stamps written by CTA (0,0) and min/max over CTAs).

Needs the instrumented build:  GG_BUILD_TIMELINE=1 bash graphical-gan_b200/build.sh
    GG_LIB=graphical-gan_b200/lib/libgg_b200_tl.so python tools/timeline_conv.py

All launches are enqueued back to back on one stream (stamp buffer reset -> launch -> snapshot, no host sync in between);
reported numbers are medians over the repetitions, in ns relative to the earliest CTA entry of the launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gpu_util as U
from gg import cabi

REPS = 30
dbg = torch.zeros(256, dtype=torch.int64, device="cuda")
template = torch.zeros(256, dtype=torch.int64, device="cuda")
template[201] = 2 ** 62
template[210] = 2 ** 62
snaps = torch.zeros(REPS, 256, dtype=torch.int64, device="cuda")
cabi.call("gg_debug_set_buffer", dbg.data_ptr())


def spin_up(seconds=0.5):
    import time
    a = torch.randn(4096, 4096, device="cuda")
    t0 = time.time()
    while time.time() - t0 < seconds:
        (a @ a).sum().item()


def run(name, fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    for r in range(REPS):
        dbg.copy_(template)
        fn()
        snaps[r].copy_(dbg)
    torch.cuda.synchronize()
    t = snaps.cpu().numpy().astype(np.int64)

    def med(slot, base=210):
        v = [int(row[slot] - row[base]) for row in t if row[slot] != 0 and row[slot] != 2 ** 62]
        return int(np.median(v)) if v else -1
    issue = [med(1 + i) for i in range(12)]
    land = [med(64 + i) for i in range(12)]
    print("%-28s first-CTA entry 0 | CTA00 entry %d | setup done(min over CTAs) %d | CTA00 setup done %d" %
          (name, med(211), med(201), med(0)))
    print("%-28s   TMA issue  %s" % ("", [v for v in issue if v >= 0]))
    print("%-28s   stage full %s" % ("", [v for v in land if v >= 0]))
    print("%-28s   accum ready %d | partial written %d | rendezvous passed %d | reduced+staged %d | stored %d | CTA00 epilogue end %d | last CTA end %d" %
          ("", med(128), med(129), med(212), med(140), med(203), med(131), med(200)))
    sys.stdout.flush()


spin_up()
cabi.call("gg_set_tc_stages", int(os.environ.get("GG_TC_STAGES", "3")))
B, H, W, Ci, Co, k, s = 64, 16, 16, 64, 128, 5, 2
x = torch.randn(B, H, W, Ci, device="cuda"); w = torch.randn(k, k, Ci, Co, device="cuda") * .05
b = torch.zeros(Co, device="cuda"); dy = torch.randn(B, H // 2, W // 2, Co, device="cuda")
A_ = torch.randn(64, 512, device="cuda"); B_ = torch.randn(512, 512, device="cuda"); bias_ = torch.zeros(512, device="cuda")
A2 = torch.randn(64, 4608, device="cuda"); B2 = torch.randn(4608, 512, device="cuda")
x1 = torch.randn(64, 32, 32, 3, device="cuda"); w1 = torch.randn(5, 5, 3, 64, device="cuda") * .05; b1 = torch.zeros(64, device="cuda")
x3 = torch.randn(B, 8, 8, 128, device="cuda"); w3 = torch.randn(k, k, 128, 256, device="cuda") * .05; b3 = torch.zeros(256, device="cuda")
xb = torch.randn(2 * B, H, W, Ci, device="cuda")
# the face / SSGAN discriminators (DIM 32, 64x64 inputs): D.2 on the batched towers, 256 x 32x32x32 -> 64
xf = torch.randn(256, 32, 32, 32, device="cuda"); wf = torch.randn(k, k, 32, 64, device="cuda") * .05; bf = torch.zeros(64, device="cuda")
dyf = torch.randn(256, 16, 16, 64, device="cuda")
cases = (
    ("face D.2 fwd B=256 32->64", lambda: U.conv_fwd(xf, wf, bf, s, 'SAME', act="leaky")),
    ("face D.2 dgrad B=256", lambda: U.conv_dgrad(dyf, wf, None, 32, 32, s, 'SAME')),
    ("face D.2 wgrad B=256", lambda: U.conv_wgrad(xf, dyf, k, s, 'SAME')),
    ("D.2 fwd batched B=128", lambda: U.conv_fwd(xb, w, b, s, 'SAME', act="leaky")),
    ("gemm 64x512x512", lambda: U.gemm(A_, B_, bias_, 64, 512, 512, 0, 0, act="leaky")),
    ("gemm 64x512x4608", lambda: U.gemm(A2, B2, bias_, 64, 512, 4608, 0, 0, act="leaky")),
    ("conv1 fwd 3->64 (im2col+tc)", lambda: U.conv_fwd(x1, w1, b1, s, 'SAME', act="leaky")),
    ("E.2 fwd", lambda: U.conv_fwd(x, w, b, s, 'SAME', act="leaky")),
    ("E.3 fwd", lambda: U.conv_fwd(x3, w3, b3, s, 'SAME', act="leaky")),
    ("E.2 dgrad", lambda: U.conv_dgrad(dy, w, None, H, W, s, 'SAME')),
    ("E.2 wgrad", lambda: U.conv_wgrad(x, dy, k, s, 'SAME')),
)
only = sys.argv[1:] 
for name, fn in cases:
    if only and not any(o in name for o in only):
        continue
    run(name, fn)
    print("%-28s   tc info %s" % ("", cabi.last_tc_info()))
