"""Run one steady-state GMGAN-CIFAR10 iteration (G step + D step) inside a cudaProfilerStart/Stop range.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -c 6 \
      -o gpurun_out/prof_conv python tools/profile_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)
os.environ.setdefault("GG_CUDA_GRAPH", "0")     # eager by default; GG_CUDA_GRAPH=1 profiles the kernel nodes of the captured graphs

import numpy as np
import torch
import tensorflow as tf
import gmgan_inference_cifar10 as S

np.random.seed(1234)
g = S.build_graph(BATCH_SIZE=64)
sess = tf.Session()
rs = np.random.RandomState(0)
batches = [rs.randint(0, 256, size=(64, 3072)).astype(np.int32) for _ in range(4)]


def iteration(i):
    sess.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: batches[(2 * i) % 4]})
    sess.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: batches[(2 * i + 1) % 4]})


for i in range(3):
    iteration(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
iteration(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one iteration")
