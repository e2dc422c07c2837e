"""Timeline of one captured training step: timed CUDA event nodes around every kernel group of the CUDA graph
(GG_TRACE=1 in gg/executor.py), read back after a replay.  Prints per group: stream, start, duration, what it is —
the measured counterpart of tools/critical_path.py's model (which kernels the step actually waits for).

    GG_TRACE=1 python tools/trace_step.py [gen|disc]
"""
import os
import sys

os.environ["GG_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)
import numpy as np
import torch
import tensorflow as tf
import gmgan_inference_cifar10 as S
from gg.executor import RT

which = sys.argv[1] if len(sys.argv) > 1 else "gen"
np.random.seed(1234)
g = S.build_graph(BATCH_SIZE=64)
sess = tf.Session()
rs = np.random.RandomState(0)
batches = [torch.from_numpy(rs.randint(0, 256, size=(64, 3072)).astype(np.int32)).cuda() for _ in range(4)]
fet = {"gen": [g.gen_cost, g.gen_train_op], "disc": [g.disc_cost, g.disc_train_op]}


def describe(plan, gi):
    grp = plan.groups[gi]
    w = grp["writes"]
    node = next((n for n in plan.order if n.id == w), None)
    if node is None:
        return str(w)
    a = node.attrs
    if node.op == "conv":
        return "conv %s B%d %dx%d %d->%d k%d" % (a["mode"], a["B"], a["H"], a["W"], a["Ci"], a["Co"], a["k"])
    if node.op in ("unary", "binary", "reduce"):
        return "%s %s %s" % (node.op, a.get("fn"), tuple(node.shape))
    return "%s %s" % (node.op, tuple(node.shape))


for i in range(6):
    for k in ("gen", "disc"):
        RT.run(fet[k], {g.real_x_int: batches[i % 4]}, to_host=False)
torch.cuda.synchronize()
plan = RT.plans[[k for k in RT.plans if k[0][0] == fet[which][0].id][0]]
acc = {}
REPS = 10
for rep in range(REPS):
    RT.run(fet[which], {g.real_x_int: batches[rep % 4]}, to_host=False)
    torch.cuda.synchronize()
    for gi, sidx, ta, tb, waits in plan.trace:
        s, e = plan.trace_elapsed_us(plan.trace_t0, ta), plan.trace_elapsed_us(plan.trace_t0, tb)
        acc.setdefault(gi, []).append((s, e))
rows = []
for gi, sidx, ta, tb, waits in plan.trace:
    s = float(np.median([v[0] for v in acc[gi]])); e = float(np.median([v[1] for v in acc[gi]]))
    rows.append((s, e, sidx, gi, waits))
end = max(r[1] for r in rows)
print("%s step: %d groups, graph span %.1f us (event nodes add overhead; relative picture)" % (which, len(rows), end))
# critical chain: walk back from the last-finishing group through the dependency (or stream predecessor) that finished last
by_gi = {r[3]: r for r in rows}
prev_on_stream = {}
last = {}
for r in sorted(rows, key=lambda r: r[0]):
    prev_on_stream[r[3]] = last.get(r[2])
    last[r[2]] = r[3]
cur = max(rows, key=lambda r: r[1])[3]
chain = []
while cur is not None:
    chain.append(cur)
    r = by_gi[cur]
    cands = [d for d in r[4] if d in by_gi]
    if prev_on_stream.get(cur) is not None:
        cands.append(prev_on_stream[cur])
    cur = max(cands, key=lambda d: by_gi[d][1]) if cands else None
chain.reverse()
print("critical chain (%d groups):" % len(chain))
tprev = 0.0
for gi in chain:
    s, e, sidx, _, _ = by_gi[gi]
    print("  st%-2d start %7.1f  dur %6.1f  wait-before %5.1f  %s" % (sidx, s, e - s, s - tprev, describe(plan, gi)))
    tprev = e
print("all groups by start time:")
for s, e, sidx, gi, waits in sorted(rows):
    print("  st%-2d %7.1f -> %7.1f (%5.1f)  %s" % (sidx, s, e, e - s, describe(plan, gi)))
