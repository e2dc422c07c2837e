"""How much of the training step is the CUDA graph's own launch + dependency cost?

Captures the gmgan-CIFAR G and D steps twice: as they are, and with every kernel launch of libgg_b200 replaced by an empty one-warp
kernel (gg_set_null_launch): same nodes, same edges, same streams / priorities, no work.  The second number is the floor the
dependency structure alone imposes on the step (node dispatch + edge resolution along the longest chain); the difference is what
kernel durations add on top.   python tools/exp_null_step.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)
import numpy as np
import torch
from gg import cabi
from gg.executor import RT
import tensorflow as tf
import gmgan_inference_cifar10 as S

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
np.random.seed(1234)
g = S.build_graph(BATCH_SIZE=64)
rs = np.random.RandomState(0)
batch = torch.from_numpy(rs.randint(0, 256, size=(64, 3072)).astype(np.int32)).cuda()


def iteration():
    RT.run([g.gen_cost, g.gen_train_op], {g.real_x_int: batch}, to_host=False)
    RT.run([g.disc_cost, g.disc_train_op], {g.real_x_int: batch}, to_host=False)


def timed(label):
    for _ in range(10):
        iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    plans = list(RT.plans.values())
    print("%-34s %.4f ms per iteration  (%d kernel nodes, %d groups)" %
          (label, ms, sum(p.kernel_launches for p in plans), sum(len(p.groups) for p in plans)))
    return ms


real = timed("real kernels")
cabi.call("gg_set_null_launch", 1)
for p in RT.plans.values():
    p.graph = None                       # re-capture: same launch list, empty kernels
null = timed("empty kernels, same graph")
cabi.call("gg_set_null_launch", 0)
print("graph structure alone: %.0f us of %.0f us (%.0f %%)" % (null * 1e3, real * 1e3, 100 * null / real))
