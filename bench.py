#!/usr/bin/env python
"""bench.py — images/sec of the GMGAN CIFAR-10 LOCAL_EP training hot path (BASELINE.json metric, configs[1]).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling
    python bench.py --impl reference ...                          # the reference's algorithm on the host cores (oracle port)

A "step" is one training iteration of gmgan_inference_cifar10.py:480-494: one generator/extractor step and one
discriminator step, each on its own synthetic 64x3072 int32 batch; images/sec = (per-GPU batch x N) x steps / time.
`value` is device-timed with inputs resident in HBM; `e2e` goes through Session.run with host numpy batches (pinned
H2D copy in, cost scalar D2H out, every step).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "graphical-gan_b200")
for p in (ROOT, PKG, os.path.join(PKG, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)

BATCH = 64                     # per GPU (gmgan_inference_cifar10.py:61)
GF_PER_ITER = 63.6             # algorithmic GFLOP per iteration at B=64
METRIC = "images/sec GMGAN CIFAR-10 LOCAL_EP bs=64/GPU (1 iteration = G step + D step)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_ips(iters, warmup, threads=None):
    """the reference's algorithm (oracle port, PyTorch-CPU fp32, all host cores) on the same workload"""
    import torch
    from oracle import gmgan_cifar10 as OM
    import tflib as lib
    import tensorflow as tf
    import gmgan_inference_cifar10 as S
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    if not lib._params:
        np.random.seed(1234)
        S.build_graph(BATCH_SIZE=BATCH)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    model = OM.GMGANCifar10(params, dtype=torch.float32)
    step = 0
    t0 = None
    for it in range(warmup + iters):
        if it == warmup:
            t0 = time.perf_counter()
        model.gen_step(**OM.synthetic_inputs(BATCH, step)); step += 1
        model.disc_step(**OM.synthetic_inputs(BATCH, step)); step += 1
    dt = time.perf_counter() - t0
    return BATCH * iters / dt, dt / iters, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    iters = max(1, min(args.steps, 20))
    ips, sec_per, threads = cpu_oracle_ips(iters, min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": args.gpus, "steps": iters,
        "warmup": min(args.warmup, 3), "ms_per_step": sec_per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "gmgan_inference_cifar10.py MODE=local_ep bs=64 32x32x3 (configs[1]), CPU oracle port of the "
                               "reference's tflib/TensorFlow path (TensorFlow itself cannot run here)"},
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": threads, "kind": "port",
                         "sample": "%d iterations (G step + D step, bs=64) of the same workload" % iters},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_dominant_kernel(torch, cabi, flush_buf):
    """live CUDA-event timing of the dominant tensor kernel: the tcgen05 conv2d forward at the Discriminator.2 shape of the
    step (fake and real towers batched: 128x16x16x64 -> 128x8x8x128, 5x5 stride 2; 3.355 GFLOP per launch, 2x SURVEY.md
    §8(d)'s 1.678 GF per 64 images).  N launches are replayed from a CUDA graph between two events, so the number is the
    kernel's device time (launch gap included), not the host's launch overhead; `cold` puts a 256 MiB memset between
    launches (its own time, measured the same way, subtracted)."""
    import numpy as np
    B, H, W, Ci, Co, k = 2 * BATCH, 16, 16, 64, 128, 5
    x = torch.randn(B, H, W, Ci, device="cuda")
    w = torch.randn(k, k, Ci, Co, device="cuda") * 0.05
    b = torch.zeros(Co, device="cuda")
    y = torch.empty(B, 8, 8, Co, device="cuda")
    ws = torch.zeros(max(int(cabi.lib.gg_conv2d_workspace(0, B, H, W, Ci, Co, k, 2, 8, 8)), 256), dtype=torch.uint8, device="cuda")
    N = 20

    def launch(st):
        cabi.call("gg_conv2d_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, H, W, Ci, Co, k, 2, 1, 1, 8, 8,
                  2, 0.2, ws.data_ptr(), ws.numel(), st)

    def graph_us(with_kernel, with_flush):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            launch(s.cuda_stream)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(N):
                    if with_flush:
                        flush_buf.zero_()
                    if with_kernel:
                        launch(torch.cuda.current_stream().cuda_stream)
            ts = []
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); g.replay(); e1.record(s); e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / N)
        return float(np.mean(ts[2:]))
    hot = graph_us(True, False)
    cold = graph_us(True, True) - graph_us(False, True)
    flops = 2.0 * B * 8 * 8 * Co * Ci * k * k
    return hot * 1e-3, cold * 1e-3, flops, cabi.lib.gg_last_backend()


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/ncu_dominant_kernel.json, written by tools/summarize_ncu.py); None when no capture is committed"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")) as f:
            return json.load(f)
    except Exception:
        return None


def run_ours(args):
    import torch
    from gg import cabi, dist as ggdist
    from gg.executor import RT
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S

    rank, world = ggdist.init_from_env()
    if world == 1 and torch.cuda.is_available():
        torch.cuda.set_device(0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    np.random.seed(1234)                      # identical initial weights on every rank
    g = S.build_graph(BATCH_SIZE=BATCH)
    sess = tf.Session()
    rs = np.random.RandomState(100 + rank)
    ring_h = [rs.randint(0, 256, size=(BATCH, 3072)).astype(np.int32) for _ in range(8)]
    ring_d = [torch.from_numpy(a).cuda() for a in ring_h]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def iteration_async(i):
        RT.run([g.gen_cost, g.gen_train_op], {g.real_x_int: ring_d[(2 * i) % 8]}, to_host=False)
        return RT.run([g.disc_cost, g.disc_train_op], {g.real_x_int: ring_d[(2 * i + 1) % 8]}, to_host=False)

    def iteration_e2e(i):
        gc, _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: ring_h[(2 * i) % 8]})
        dc, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: ring_h[(2 * i + 1) % 8]})
        return gc, dc

    # ---- warm-up (also captures the two CUDA graphs) ----
    for i in range(max(args.warmup, 3)):
        iteration_async(i)
    barrier()

    # ---- kernel-only: per-step CUDA events, L2 flushed between steps outside the event brackets ----
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    barrier()
    evs = []
    for i in range(args.steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iteration_async(i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    per_step = [e0.elapsed_time(e1) for e0, e1 in evs]
    dev_ms = sum(per_step)
    pct = [float(np.percentile(per_step, q)) for q in (10, 50, 90)]
    # hot-L2 variant (no flush), back-to-back, one event pair around all K steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        iteration_async(i)
    e1.record()
    barrier()
    hot_ms = e0.elapsed_time(e1)

    if args.quick:
        # experiment mode (tools/, env-knob sweeps): device-timed numbers only, not a bench line
        if rank == 0:
            print(json.dumps({"quick": True, "images_per_sec": BATCH * world * args.steps / (dev_ms / 1e3),
                              "ms_per_step": dev_ms / args.steps, "ms_per_step_hot_l2": hot_ms / args.steps,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("GG_")}}))
        return
    # ---- end to end through Session.run: host batches in, cost scalars out, every step ----
    for i in range(3):
        iteration_e2e(i)
    barrier()
    t0 = time.perf_counter()
    last = None
    for i in range(args.steps):
        last = iteration_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([dev_ms, hot_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    dev_ms, hot_ms, e2e_ms = [float(v) for v in times.cpu()]
    if rank != 0:
        return

    gplan = RT.plans[[k for k in RT.plans if k[0][0] == g.gen_cost.id][0]]
    dplan = RT.plans[[k for k in RT.plans if k[0][0] == g.disc_cost.id][0]]
    launches_per_iter = int(gplan.kernel_launches + dplan.kernel_launches)
    images = BATCH * world * args.steps
    value = images / (dev_ms / 1e3)
    hbm_peak, tf_peak, peak_kind = load_peaks()
    k_ms_hot, k_ms, k_flops, k_backend = time_dominant_kernel(torch, cabi, flush_buf)
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    ncu = load_ncu_traffic()
    cpu = None
    if world == 1:
        ips, _, threads = cpu_oracle_ips(8, 2)
        cpu = {"value": ips, "unit": "images/sec", "cores": threads, "kind": "port",
               "sample": "8 iterations (G step + D step, bs=64) of the same workload on the host cores, PyTorch-CPU fp32"}
    line = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 storage, tf32 tensor-core operands / f32 accumulate on the 5x5 convs",
        "data": "synthetic",
        "config": {"workload": "gmgan_inference_cifar10.py MODE=local_ep bs=64 per GPU, 32x32x3 (BASELINE.json configs[1])",
                   "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                   "l2": "256 MiB memset between timed steps, outside the per-step CUDA-event brackets",
                   "ms_per_step_hot_l2": hot_ms / args.steps, "ms_per_step_p10_p50_p90_rank0": pct,
                   "cuda_graph": bool(RT.use_cuda_graph),
                   "algorithmic_gflop_per_iteration": GF_PER_ITER,
                   "last_costs": [float(last[0]), float(last[1])]},
        "e2e": {"value": images / (e2e_ms / 1e3), "unit": "images/sec",
                "h2d_bytes_per_step": 2 * BATCH * 3072 * 4, "d2h_bytes_per_step": 8},
        "gpu_launches": launches_per_iter * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                     "traffic": (ncu or {}).get("dram_bytes_per_launch"),
                     "kernel": "gg_conv2d_fwd 128x16x16x64->128 5x5 s2, Discriminator.2 on the batched fake+real towers (%s)" %
                               ("tcgen05 tf32" if k_backend else "direct fp32"),
                     "algorithmic_flop_per_launch": k_flops, "kernel_ms": k_ms, "kernel_ms_hot_l2": k_ms_hot,
                     "timing": "CUDA graph of 20 launches between two events, 256 MiB memset between launches (subtracted)",
                     "peak_kind": peak_kind + " dense bf16 burst (kind::tf32 peaks at half of it)",
                     "ncu": (ncu or {}).get("source")},
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="device-timed throughput only (knob sweeps); prints a short JSON")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
