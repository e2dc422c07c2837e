"""Calibration target for the ncu tensor-pipe metrics (`sm__pipe_tensor_cycles_active_realtime` read 1.6 % on
a launch whose FLOP-derived tensor utilisation is ~19 %).  Runs two library GEMMs whose tensor utilisation is known from their
own timing — cuBLAS fp32-with-tf32 and bf16, 8192^3 — plus one Discriminator.2 forward of this library, so that one
`ncu --metrics sm__pipe_tensor_cycles_active_realtime...,sm__inst_executed_pipe_tensor_subpipe_hmma...` pass shows what the
counters report for UTCHMMA-class work at a known rate.

    ncu --metrics <list> --clock-control none --csv --log-file gpurun_out/calib.csv python tools/ncu_calib_tensor.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(2):
    c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
print("cublas tf32 %d^3: %.1f us, %.1f TFLOP/s" % (n, e0.elapsed_time(e1) * 1e3, 2 * n ** 3 / e0.elapsed_time(e1) / 1e9))
ah, bh = a.bfloat16(), b.bfloat16()
for _ in range(2):
    ch = ah @ bh
torch.cuda.synchronize()
e0.record(); ch = ah @ bh; e1.record(); torch.cuda.synchronize()
print("cublas bf16 %d^3: %.1f us, %.1f TFLOP/s" % (n, e0.elapsed_time(e1) * 1e3, 2 * n ** 3 / e0.elapsed_time(e1) / 1e9))

import gpu_util as U
from gg import cabi
x = torch.randn(128, 16, 16, 64, device="cuda")
w = torch.randn(5, 5, 64, 128, device="cuda") * 0.05
bias = torch.zeros(128, device="cuda")
for _ in range(3):
    y = U.conv_fwd(x, w, bias, 2, 'SAME', act="leaky")
torch.cuda.synchronize()
print("conv_tc D.2 batched forward launched (3.355 GFLOP per launch)", cabi.last_tc_info())
