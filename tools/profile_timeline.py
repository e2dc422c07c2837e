"""True GPU timeline of one captured training step via CUPTI (torch.profiler, kernel activity records): per kernel
name / stream / start / duration inside the multi-stream CUDA-graph replay.  Summarises concurrency and idle time.

    python tools/profile_timeline.py [gen|disc] > profiles/timeline_<tag>.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import tensorflow as tf
import gmgan_inference_cifar10 as S
from gg.executor import RT
from gg import dist as ggdist, cabi

rank, world = ggdist.init_from_env()     # under torchrun: the data-parallel step of THIS rank (rank 0 prints)

which = sys.argv[1] if len(sys.argv) > 1 else "gen"
np.random.seed(1234)
g = S.build_graph(BATCH_SIZE=64)
sess = tf.Session()
rs = np.random.RandomState(0)
batches = [torch.from_numpy(rs.randint(0, 256, size=(64, 3072)).astype(np.int32)).cuda() for _ in range(4)]
fet = {"gen": [g.gen_cost, g.gen_train_op], "disc": [g.disc_cost, g.disc_train_op]}
for i in range(6):
    for k in ("gen", "disc"):
        RT.run(fet[k], {g.real_x_int: batches[i % 4]}, to_host=False)
torch.cuda.synchronize()
plan = RT.plans[[q for q in RT.plans if q[0][0] == fet[which][0].id][0]]
def replay():
    for seg in plan.graph:
        if callable(seg):
            seg(cabi.stream_ptr())
        else:
            seg.replay()


with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for r in range(3):
        replay()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
if rank != 0:
    sys.exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
evs = evs[2 * n:]                     # the last replay
t0 = evs[0].time_range.start
rows = [(e.time_range.start - t0, e.time_range.end - t0, e.name) for e in evs]
span = max(r[1] for r in rows)
busy = sum(r[1] - r[0] for r in rows)
# time with k kernels in flight
pts = sorted([(r[0], 1) for r in rows] + [(r[1], -1) for r in rows])
hist, cur, last = {}, 0, 0.0
for t, d in pts:
    hist[cur] = hist.get(cur, 0.0) + (t - last)
    cur += d
    last = t
print("%s step: %d kernels, span %.1f us, sum of kernel durations %.1f us (avg concurrency %.2f)" % (which, len(rows), span, busy, busy / span))
print("time with k kernels in flight: " + ", ".join("%d: %.0f us" % (k, v) for k, v in sorted(hist.items())))
agg = {}
for s, e, name in rows:
    key = name.split("(")[0].replace("void ", "").replace("gg::", "").replace("<unnamed>::", "")[:40]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += e - s
for key, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("   %-42s n=%3d total %7.1f us avg %5.1f" % (key, c, us, us / c))
print("timeline (start, dur, name):")
for s, e, name in rows:
    print("  %7.1f %6.1f  %s" % (s, e - s, name.replace("void ", "").replace("gg::", "").replace("<unnamed>::", "")[:70]))
