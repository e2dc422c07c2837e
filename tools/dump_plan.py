"""OFFLINE: print the launch list of the compiled gmgan-CIFAR training step (CPU box, nothing is launched).

One line per kernel group: the C-ABI calls it makes, the graph node (op, fn, shape), the nodes it reads and how many
consumers its output has — the worksheet for launch-count work (which glue launches can fold into a neighbour).

    python tools/dump_plan.py [gen|disc|both] [--config cifar|face|ssgan] [--hist]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)

import numpy as np
import torch

from gg import cabi, executor
from gg.executor import RT, Plan

executor.Runtime.dev = lambda self: torch.device("cpu")
torch.Tensor.pin_memory = lambda self, *a, **k: self
_real_call = cabi.call
HOST_ONLY = ("gg_set_tc_max_ctas", "gg_set_tc_stages", "gg_set_pdl")
_record = []


def _fake_call(name, *args):
    if name in HOST_ONLY:
        return _real_call(name, *args)
    _record.append((name, args))


cabi.call = _fake_call


def build(config, batch):
    import tensorflow as tf  # noqa: F401  (the shim)
    np.random.seed(1234)
    if config == "cifar":
        import gmgan_inference_cifar10 as S
        g = S.build_graph(BATCH_SIZE=batch or 64)
        feeds = [g.real_x_int]
    elif config == "face":
        import gan_inference_face as S
        g = S.build_graph(BATCH_SIZE=batch or 128)
        feeds = [g.real_x_int]
    else:
        import ssgan_inference_moving_mnist as S
        g = S.build_graph(BATCH_SIZE=batch or 32)
        feeds = [getattr(g, n) for n in ("real_x_int", "real_x", "real_x_unit") if hasattr(g, n)][:1]
        from gg.ops import toposort
        roots = [g.gen_cost, g.disc_cost] + [d for op in (g.gen_train_op, g.disc_train_op) for d in op.deps if d is not None]
        feeds = [n for n in toposort(roots) if n.op == "placeholder"]
    return g, feeds


def short(node):
    if node is None:
        return "-"
    a = node.attrs
    if node.op == "conv":
        s = "conv.%s B%d %dx%d %d->%d k%d act=%s%s" % (a["mode"], a["B"], a["H"], a["W"], a["Ci"], a["Co"], a["k"], a.get("act"),
                                                      " +bias" if len(node.inputs) == 3 else "")
    elif node.op == "matmul":
        s = "matmul%s ta%d tb%d act=%s%s" % (tuple(node.shape), a["ta"], a["tb"], a.get("act"), " +bias" if len(node.inputs) == 3 else "")
    elif node.op in ("unary", "binary", "reduce"):
        s = "%s.%s%s" % (node.op, a.get("fn"), tuple(node.shape))
        if node.op == "reduce":
            s += " axes=%s of %s" % (a.get("axes"), tuple(node.inputs[0].shape))
    elif node.op == "transpose":
        s = "transpose%s perm=%s" % (tuple(node.shape), a.get("perm"))
    elif node.op in ("slice", "concat", "pad"):
        s = "%s%s axis=%s" % (node.op, tuple(node.shape), a.get("axis"))
    else:
        s = "%s%s" % (node.op, tuple(node.shape))
    return "#%d %s" % (node.id, s)


def dump(plan, label, hist):
    uses = collections.Counter()
    for n in plan.order:
        if n.id in plan.fed:
            continue
        for i in n.inputs:
            uses[i.id] += 1
    calls = collections.Counter()
    print("== %s: %d kernel steps, %d groups" % (label, len(plan.steps), len(plan.groups)))
    for gi, grp in enumerate(plan.groups):
        names = []
        for f in plan.steps[grp["start"]:grp["end"]]:
            del _record[:]
            try:
                f(0)
            except Exception as e:      # collectives etc.
                _record.append(("<%s>" % type(e).__name__, ()))
            names += [r[0] for r in _record]
        for nm in names:
            calls[nm] += 1
        node = grp.get("node")
        if not hist:
            ins = " <- " + ", ".join(short(i) for i in node.inputs) if node is not None else ""
            print("%4d %-34s %s  uses=%d%s" % (gi, "+".join(n.replace("gg_", "") for n in names), short(node) if node is not None else grp["writes"],
                                              uses.get(node.id, 0) if node is not None else 0, ins[:230]))
    print("-- calls:", ", ".join("%s x%d" % kv for kv in calls.most_common()))
    return calls


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="both")
    ap.add_argument("--config", default="cifar")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--hist", action="store_true")
    args = ap.parse_args()
    g, feeds = build(args.config, args.batch)
    total = 0
    if args.which in ("gen", "both"):
        gp = Plan(RT, [g.gen_cost, g.gen_train_op], feeds)
        dump(gp, "G step", args.hist)
        total += len(gp.steps)
    if args.which in ("disc", "both"):
        dp = Plan(RT, [g.disc_cost, g.disc_train_op], feeds)
        dump(dp, "D step", args.hist)
        total += len(dp.steps)
    print("steps per iteration: %d" % total)


if __name__ == "__main__":
    main()
