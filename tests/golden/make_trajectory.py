"""Generates tests/golden/trajectory100.npz: the first 100 training iterations of gmgan_inference_cifar10.py (MODE local_ep,
bs 64; loop :480-494 — iteration 0 runs the D step only) evaluated by the float64 CPU oracle (oracle/gmgan_cifar10.py) with
tflib-initialised weights (np.random.seed(1234), graph-construction order) and injected per-step noise
(oracle.gmgan_cifar10.synthetic_inputs).  Stored: both cost curves and the generator's fixed-noise samples
(gmgan_inference_cifar10.py:413-419, first N_KEEP of the 300) at a few checkpoints.  SURVEY.md §8(c) "golden vectors to
create"; parity unpinned (the oracle is a restatement — TensorFlow cannot run here).

    python tests/golden/make_trajectory.py        # ~5-10 min on 8 cores
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)

ITERS = 100
CHECKPOINTS = (1, 2, 5, 10, 20, 50, 100)
N_KEEP = 12
BATCH = 64


def main():
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    from oracle import gmgan_cifar10 as OM
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=BATCH)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    k1h, noise = g.np_fixed_k.astype(np.float32), g.np_fixed_noise      # all N_VIS = 300: the generator's batch norm uses THEIR batch statistics
    runs = {}
    # float64 = the arbiter.  float32 = the SAME oracle in the reference's own precision: how far an fp32 implementation of
    # the identical algorithm drifts from the arbiter is the noise floor any other implementation is measured against.
    for tag, dt in (("", torch.float64), ("32", torch.float32)):
        oracle = OM.GMGANCifar10(params, dtype=dt)
        gen_costs, disc_costs, samples, stats = np.full(ITERS, np.nan), np.zeros(ITERS), {}, {}
        step, t0 = 0, time.time()
        for it in range(ITERS):
            if it > 0:
                gen_costs[it], _ = oracle.gen_step(**OM.synthetic_inputs(BATCH, step)); step += 1
            disc_costs[it], _ = oracle.disc_step(**OM.synthetic_inputs(BATCH, step)); step += 1
            if it + 1 in CHECKPOINTS:
                full = oracle.sample(k1h, noise).numpy().astype(np.float32)
                samples[it + 1] = full[:N_KEEP]
                stats[it + 1] = (float(full.mean()), float(full.std()))
                print("%s iteration %d  gen %.6f disc %.6f  (%.0f s)" % (dt, it + 1, gen_costs[it], disc_costs[it], time.time() - t0), flush=True)
        runs["gen_costs" + tag], runs["disc_costs" + tag] = gen_costs, disc_costs
        runs["samples" + tag] = np.stack([samples[c] for c in CHECKPOINTS])
        runs["stats" + tag] = np.asarray([stats[c] for c in CHECKPOINTS])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trajectory100.npz")
    np.savez_compressed(out, checkpoints=np.asarray(CHECKPOINTS), n_keep=N_KEEP, batch=BATCH, **runs)
    d = runs["samples32"] - runs["samples"]
    print("fp32 oracle vs fp64 oracle, sample rel-L2 per checkpoint:",
          [float(np.linalg.norm(d[i]) / np.linalg.norm(runs["samples"][i])) for i in range(len(CHECKPOINTS))])
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
