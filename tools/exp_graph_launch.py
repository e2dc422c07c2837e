"""Is the captured step bound by the CPU cost of cudaGraphLaunch rather than by the GPU?
For each of the two step graphs: CPU time of replay() with an idle GPU, GPU span of one replay, and back-to-back rate."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)
import numpy as np
import torch
import tensorflow as tf
import gmgan_inference_cifar10 as S
from gg.executor import RT

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
for k in ("gen", "disc"):
    plan = RT.plans[[q for q in RT.plans if q[0][0] == fet[k][0].id][0]]
    gr = plan.graph[0]
    cpu, tot, gpu = [], [], []
    for r in range(30):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        gr.replay()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        cpu.append((t1 - t0) * 1e6); tot.append((t2 - t0) * 1e6); gpu.append(e0.elapsed_time(e1) * 1e3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(200):
        gr.replay()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%s graph: %d kernels | idle-GPU replay(): CPU %.0f us, GPU span %.0f us, launch->done %.0f us | 200 back-to-back: CPU enqueue %.0f us each, "
          "wall %.0f us each" % (k, plan.kernel_launches, np.median(cpu), np.median(gpu), np.median(tot), (t1 - t0) * 1e6 / 200, (t2 - t0) * 1e6 / 200), flush=True)
# whole iteration through RT.run (python + feed copy + 2 graph launches), CPU enqueue cost only
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(100):
    RT.run(fet["gen"], {g.real_x_int: batches[i % 4]}, to_host=False)
    RT.run(fet["disc"], {g.real_x_int: batches[(i + 1) % 4]}, to_host=False)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("RT.run iteration: CPU enqueue %.0f us, wall %.0f us per iteration" % ((t1 - t0) * 1e4, (t2 - t0) * 1e4))
