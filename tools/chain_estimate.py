"""OFFLINE: longest dependency chain of the compiled step under a per-call duration table (averages of the committed ncu launch
list, profiles/launches_r2_iteration.txt) — which glue launches sit ON the chain, i.e. what a fusion would actually shorten.

    python tools/chain_estimate.py [gen|disc] [--config cifar] [--gap-us 1.5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dump_plan as DP
from dump_plan import Plan, RT, _record

US = {"gg_unary": 2.9, "gg_binary": 3.2, "gg_ew_run": 3.5, "gg_reduce": 4.0, "gg_reduce_ws": 9.0, "gg_transpose_b2d": 5.0,
      "gg_transpose_b2d_ex": 5.0, "gg_copy2d": 3.8, "gg_add_n": 3.7, "gg_softmax_fwd": 3.0, "gg_softmax_bwd": 3.0, "gg_one_hot": 2.5,
      "gg_rng_tick": 2.0, "gg_rng_normal": 3.0, "gg_rng_uniform": 3.0, "gg_rng_categorical": 3.0, "gg_cast_i32_f32": 3.0,
      "gg_bn_fwd_fused": 7.6, "gg_bn_bwd_fused": 9.6, "gg_adam_multi": 16.0, "gg_conv2d_dgrad_actgrad": 14.6, "gg_fill": 2.5}


def cost(name, node):
    if name in ("gg_conv2d_fwd", "gg_conv2d_dgrad", "gg_conv2d_wgrad"):
        a = node.attrs
        if min(a["Ci"], 10 ** 9) <= 4:
            return {"gg_conv2d_fwd": 14.5, "gg_conv2d_dgrad": 18.2, "gg_conv2d_wgrad": 22.1}[name]
        return {"gg_conv2d_fwd": 13.0, "gg_conv2d_dgrad": 14.6, "gg_conv2d_wgrad": 11.2}[name]
    if name == "gg_gemm":
        M, N = node.shape
        K = node.inputs[0].shape[0] if node.attrs["ta"] else node.inputs[0].shape[1]
        if N % 32 == 0 and K % 32 == 0 and M * N > 8192:
            return 12.0
        return 7.0
    return US.get(name, 3.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="gen")
    ap.add_argument("--config", default="cifar")
    ap.add_argument("--gap-us", type=float, default=1.5)
    args = ap.parse_args()
    g, feeds = DP.build(args.config, 0)
    fetch = [g.gen_cost, g.gen_train_op] if args.which == "gen" else [g.disc_cost, g.disc_train_op]
    plan = Plan(RT, fetch, feeds)
    n = len(plan.groups)
    dur, names = [], []
    for grp in plan.groups:
        t, nm = 0.0, []
        for f in plan.steps[grp["start"]:grp["end"]]:
            del _record[:]
            f(0)
            for r in _record:
                t += cost(r[0], grp.get("node")) + args.gap_us
                nm.append(r[0].replace("gg_", ""))
        dur.append(t)
        names.append("+".join(nm))
    producer, finish, pred, last_barrier, done = {}, [0.0] * n, [None] * n, None, []
    for gi, grp in enumerate(plan.groups):
        deps = [producer[o] for o in grp["reads"] if o in producer]
        if last_barrier is not None:
            deps.append(last_barrier)
        if grp["barrier"]:
            deps += done
        start, who = 0.0, None
        for d in deps:
            if finish[d] > start:
                start, who = finish[d], d
        finish[gi] = start + dur[gi]
        pred[gi] = who
        producer[grp["writes"]] = gi
        done.append(gi)
        if grp["barrier"]:
            last_barrier = gi
    end = max(range(n), key=lambda i: finish[i])
    chain, i = [], end
    while i is not None:
        chain.append(i)
        i = pred[i]
    chain.reverse()
    print("%s step: %d groups, serial %.0f us, chain %.0f us over %d groups" % (args.which, n, sum(dur), finish[end], len(chain)))
    for i in chain:
        print("  %7.1f  %5.1f  %-28s %s" % (finish[i], dur[i], names[i][:28], DP.short(plan.groups[i].get("node"))[:110]))


if __name__ == "__main__":
    main()
