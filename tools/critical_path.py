"""Offline critical-path analysis of the compiled training step (runs on the CPU box, no GPU needed).

The plan compiler is run against a CPU "device" (buffers are ordinary host tensors, nothing is launched): that gives
the exact launch list, the kernel groups and their data dependencies.  Each C-ABI call of the list is then matched
with the kernels of an ncu launch list of the same iteration (profiles/launches_*.csv, eager order = plan order), so
every group gets a measured duration.  Output: serial sum, critical path (longest dependency chain, + a fixed
per-kernel dependency latency), and the chain itself — i.e. what bounds the step when the SMs are not the limit.

    python tools/critical_path.py [profiles/launches_r1_iteration.csv] [--gap-us 2.0] [--show 80]
"""
import argparse
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)

import numpy as np
import torch

from gg import cabi, executor
from gg.executor import RT, Plan

# ---- CPU "device": buffers only, no launches -----------------------------------------------------------------
executor.Runtime.dev = lambda self: torch.device("cpu")
_orig_pin = torch.Tensor.pin_memory
torch.Tensor.pin_memory = lambda self, *a, **k: self
_real_call = cabi.call
HOST_ONLY = ("gg_set_tc_max_ctas", "gg_set_tc_stages", "gg_set_pdl")
_record = []


def _fake_call(name, *args):
    if name in HOST_ONLY:
        return _real_call(name, *args)
    _record.append((name, args))


cabi.call = _fake_call

# kernels each C-ABI entry point may launch (prefix match on the ncu kernel name)
KERNELS = {
    "gg_gemm": ("gemm_kernel", "gemm_splitk_finish", "conv_tc_kernel", "im2col", "col2im"),
    "gg_conv2d_fwd": ("conv_tc_kernel<0>", "im2col_kernel", "conv_fwd_kernel", "conv_small_fwd_kernel"),
    "gg_conv2d_dgrad": ("conv_tc_kernel<1>", "col2im_kernel", "conv_dgrad_kernel", "conv_small_dgrad_kernel"),
    "gg_conv2d_wgrad": ("conv_tc_kernel<2>", "im2col_kernel", "conv_wgrad_kernel", "wgrad_finish", "conv_small_wgrad_kernel"),
    "gg_reduce_ws": ("reduce_cols_sliced_kernel", "reduce_cols_kernel", "reduce_rows_kernel"),
    "gg_reduce": ("reduce_cols_kernel", "reduce_rows_kernel", "reduce_all"),
    "gg_adam_multi": ("adam_tick_kernel", "adam_multi_kernel"),
}
MAX_K = {"gg_gemm": 2, "gg_conv2d_fwd": 2, "gg_conv2d_dgrad": 2, "gg_conv2d_wgrad": 2, "gg_reduce_ws": 2, "gg_reduce": 1,
         "gg_adam_multi": 2}


def load_launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) > vi and r[0].isdigit():
            name = r[ki].replace("void ", "").replace("gg::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
            out.append((name, float(r[vi]) / 1e3))
    return out


def match(calls, launches, pos):
    """greedy alignment of the C-ABI call list with the kernel list; returns per-call [(kernel, us)], new position"""
    per_call = []
    for name, _ in calls:
        allowed = KERNELS.get(name)
        got = []
        if allowed is None:
            got.append(launches[pos]); pos += 1
        else:
            while pos < len(launches) and len(got) < MAX_K[name] and any(launches[pos][0].startswith(a) for a in allowed):
                # a following call of the same entry point must keep at least one kernel: stop at a repeated "first" kernel
                if got and name in ("gg_gemm",) and launches[pos][0].startswith(("gemm_kernel", "conv_tc_kernel", "im2col")) \
                        and not got[-1][0].startswith("im2col"):
                    break
                if got and name.startswith("gg_conv2d") and launches[pos][0].startswith("im2col") :
                    break
                if got and got[-1][0].startswith("conv_small"):      # the one-launch small-channel kernels stand alone
                    break
                if got and name.startswith("gg_conv2d") and launches[pos][0].startswith("conv_tc") and got[-1][0].startswith("conv_tc"):
                    break
                if got and name in ("gg_reduce_ws",) and launches[pos][0].startswith("reduce_cols_sliced"):
                    break
                got.append(launches[pos]); pos += 1
            if not got:
                raise RuntimeError("cannot match %s at launch %d (%s)" % (name, pos, launches[pos][0]))
        per_call.append(got)
    return per_call, pos


def build_plans(batch=64):
    import tensorflow as tf
    import gmgan_inference_cifar10 as S
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=batch)
    feeds = [g.real_x_int]
    gp = Plan(RT, [g.gen_cost, g.gen_train_op], feeds)
    dp = Plan(RT, [g.disc_cost, g.disc_train_op], feeds)
    return g, gp, dp


def record_steps(plan):
    calls = []
    for f in plan.steps:
        del _record[:]
        f(0)
        assert len(_record) == 1, _record
        calls.append(_record[0])
    return calls


def describe(plan, gi):
    grp = plan.groups[gi]
    w = grp["writes"]
    node = next((n for n in plan.order if n.id == w), None)
    if node is None:
        return str(w)
    extra = ""
    if node.op == "conv":
        a = node.attrs
        extra = " %s B%d %dx%d %d->%d k%d" % (a["mode"], a["B"], a["H"], a["W"], a["Ci"], a["Co"], a["k"])
    elif node.op == "matmul":
        extra = " %s" % (tuple(node.shape),)
    elif node.op in ("unary", "binary", "reduce"):
        extra = " %s %s" % (node.attrs.get("fn"), tuple(node.shape))
    else:
        extra = " %s" % (tuple(node.shape),)
    return node.op + extra


def analyse(plan, calls, per_call, gap_us, show, label):
    n = len(plan.groups)
    dur = []
    for grp in plan.groups:
        ks = [k for c in per_call[grp["start"]:grp["end"]] for k in c]
        dur.append((sum(u for _, u in ks), len(ks), ks))
    producer, finish, pred = {}, [0.0] * n, [None] * n
    last_barrier = None
    all_done = []
    for gi, grp in enumerate(plan.groups):
        deps = [producer[o] for o in grp["reads"] if o in producer]
        if last_barrier is not None:
            deps.append(last_barrier)
        if grp["barrier"]:
            deps += all_done
        start, who = 0.0, None
        for d in deps:
            if finish[d] > start:
                start, who = finish[d], d
        finish[gi] = start + dur[gi][0] + gap_us * dur[gi][1]
        pred[gi] = who
        producer[grp["writes"]] = gi
        all_done.append(gi)
        if grp["barrier"]:
            last_barrier = gi
    end = max(range(n), key=lambda i: finish[i])
    chain = []
    i = end
    while i is not None:
        chain.append(i)
        i = pred[i]
    chain.reverse()
    serial = sum(d[0] for d in dur)
    nk = sum(d[1] for d in dur)
    print("== %s: %d kernels, serial sum %.0f us, critical path %.0f us (%d kernels on it, gap %.1f us/kernel)" %
          (label, nk, serial, finish[end], sum(dur[i][1] for i in chain), gap_us))
    agg = {}
    for i in chain:
        for kname, us in dur[i][2]:
            key = kname.split("(")[0]
            a = agg.setdefault(key, [0, 0.0])
            a[0] += 1; a[1] += us
    for key, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("   on path: %-40s n=%3d  %7.1f us" % (key[:40], cnt, us))
    if show:
        for i in chain[:show]:
            print("   %6.1f us  %-60s %s" % (dur[i][0], describe(plan, i)[:60], ",".join(k.split("(")[0][:24] for k, _ in dur[i][2])))
    return finish[end], serial


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="?", default=os.path.join(ROOT, "profiles", "launches_r1_iteration.csv"))
    ap.add_argument("--gap-us", type=float, default=2.0)
    ap.add_argument("--show", type=int, default=0)
    args = ap.parse_args()
    launches = load_launches(args.csv)
    g, gp, dp = build_plans()
    pos = 0
    tot_cp = tot_serial = 0.0
    for label, plan in (("G step", gp), ("D step", dp)):
        calls = record_steps(plan)
        per_call, pos = match(calls, launches, pos)
        cp, serial = analyse(plan, calls, per_call, args.gap_us, args.show, label)
        tot_cp += cp; tot_serial += serial
    print("matched %d of %d launches; iteration: serial %.0f us, critical path %.0f us" % (pos, len(launches), tot_serial, tot_cp))


if __name__ == "__main__":
    main()
