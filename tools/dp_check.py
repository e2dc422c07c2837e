"""Data-parallel consistency check: one D step + one G step of GMGAN-CIFAR10 on a GLOBAL batch of 64, run either on one
GPU (batch 64) or sharded over N ranks (batch 64/N each, gradient all-reduce + SyncBN).  Writes costs and a few updated
parameters to --out so the caller can compare the two runs.

  python tools/dp_check.py --out /tmp/single.npz
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py --out /tmp/dp2.npz
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--global-batch", type=int, default=64)
args = ap.parse_args()

from gg import dist as ggdist
rank, world = ggdist.init_from_env()
import tensorflow as tf
import tflib as lib
import gmgan_inference_cifar10 as S
from oracle import gmgan_cifar10 as OM          # only for the shared synthetic input generator

np.random.seed(1234)
local = args.global_batch // world
g = S.build_graph(BATCH_SIZE=local)
sess = tf.Session()
costs = []
snap = {}
for step, (cost, op) in enumerate(((g.disc_cost, g.disc_train_op), (g.gen_cost, g.gen_train_op), (g.disc_cost, g.disc_train_op))):
    inp = OM.synthetic_inputs(args.global_batch, step)
    lo, hi = ggdist.shard_bounds(args.global_batch, rank, world)
    feeds = {g.real_x_int: inp["real_x_int"][lo:hi], g.hyper_p_z: inp["hyper_p_z"][lo:hi], g.hyper_p_k_idx: inp["k_idx"][lo:hi],
             g.gumbel_uniforms[0]: inp["U"][lo:hi]}
    c, _ = sess.run([cost, op], feed_dict=feeds)
    t = torch.tensor([float(c)], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t)
    costs.append(float(t) / world)               # mean of equal-sized shard means == global mean
    from gg.executor import RT as _RT
    for n_ in ('Discriminator.zx1.W', 'Discriminator.2.Filters', 'Generator.3.Filters', 'Generator.BN2.scale'):
        snap['s%d_%s' % (step, n_.replace('.', '_'))] = _RT.get_param(lib._params[n_]).reshape(-1)[:4096].copy()
from gg.executor import RT
names = ['Discriminator.2.Filters', 'Discriminator.zx1.W', 'Generator.3.Filters', 'Generator.BN2.scale', 'Extractor.BN3.offset',
         'Generator.Hyper.Mu', 'Extractor.Output.W']
params = {n: RT.get_param(lib._params[n]).reshape(-1)[:65536] for n in names}
if world > 1:
    # every rank must hold identical parameters after the all-reduced update
    for n in names:
        t = torch.from_numpy(params[n]).cuda()
        lo_, hi_ = t.clone(), t.clone()
        torch.distributed.all_reduce(lo_, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi_, op=torch.distributed.ReduceOp.MAX)
        assert torch.equal(lo_, hi_), "parameter %s diverged across ranks" % n
    # in-graph random draws must differ per rank (the rank is mixed into the Philox key): each rank owns ITS slice of the
    # global noise batch, not a copy of rank 0's
    z = torch.from_numpy(np.ascontiguousarray(sess.run(g.hyper_p_z))).cuda()
    zs = [torch.empty_like(z) for _ in range(world)]
    torch.distributed.all_gather(zs, z)
    for r in range(1, world):
        assert not torch.equal(zs[0], zs[r]), "ranks 0 and %d drew identical in-graph noise" % r
if rank == 0:
    np.savez(args.out, costs=np.array(costs), **snap, **{n.replace('.', '_'): v for n, v in params.items()})
    print("dp_check world=%d costs=%s" % (world, costs))
if world > 1:
    torch.distributed.barrier()
    ggdist.shutdown()      # drops the captured graphs (NCCL nodes inside) BEFORE the process group goes away
