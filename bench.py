#!/usr/bin/env python
"""bench.py — throughput of the Graphical-GAN adversarial training hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path, configs[1] (GMGAN CIFAR-10)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU
    python bench.py --impl reference ...                          # the reference's algorithm on the host cores (oracle port)
    python bench.py --config face|ssgan ...                       # the 64x64 workloads: configs[3] / configs[4]

A "step" is one training iteration of the script's loop (gmgan_inference_cifar10.py:480-494): one generator/extractor step
and one discriminator step, each on its own synthetic batch; images/sec = global batch x steps / time.
`value` is device-timed with inputs resident in HBM; `e2e` goes through Session.run with host numpy batches (uint8 pixels
staged in pinned memory, H2D copy + on-device decode in, cost scalar D2H out, every step).  Rank 0 prints ONE JSON line.

  --config cifar  gmgan_inference_cifar10.py MODE=local_ep, 64 images of 32x32x3 PER GPU (weak scaling; the headline metric)
  --config face   gan_inference_face.py MODE=ali, GLOBAL batch 128 of 64x64x3 sharded over the ranks (strong scaling, configs[3])
  --config ssgan  ssgan_inference_moving_mnist.py MODE=local_ep, LEN 8, GLOBAL batch 32 sequences of 1x64x64 frames sharded
                  over the ranks (strong scaling, configs[4]); unit = frames/sec
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


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------------------
class Cifar(object):
    """BASELINE.json configs[1]"""
    name = "cifar"
    metric = "images/sec GMGAN CIFAR-10 LOCAL_EP bs=64/GPU (1 iteration = G step + D step)"
    unit = "images/sec"
    scaling = "weak"
    gf_per_iter_per_unit = 63.6 / 64          # algorithmic GFLOP per image and iteration
    dom = (128, 16, 16, 64, 128)              # dominant kernel: Discriminator.2 forward on the batched fake+real towers
    dom_name = "gg_conv2d_fwd 128x16x16x64->128 5x5 s2, Discriminator.2 on the batched fake+real towers"

    def __init__(self, world, rank):
        self.per_rank, self.world, self.rank = 64, world, rank
        self.units_per_iter = 64 * world
        self.workload = "gmgan_inference_cifar10.py MODE=local_ep bs=64 per GPU, 32x32x3 (BASELINE.json configs[1])"

    def build(self):
        import gmgan_inference_cifar10 as S
        self.g = S.build_graph(BATCH_SIZE=self.per_rank)
        rs = np.random.RandomState(100 + self.rank)
        self.ring = [rs.randint(0, 256, size=(self.per_rank, 3072)).astype(np.uint8) for _ in range(8)]   # tflib/cifar10.py dtype
        return self.g

    def host_feeds(self, j):
        return {self.g.real_x_int: self.ring[j % 8]}

    def device_feeds(self, torch):
        self.ring_d = [torch.from_numpy(a.astype(np.int32)).cuda() for a in self.ring]
        return lambda j: {self.g.real_x_int: self.ring_d[j % 8]}

    def oracle(self, params, torch):
        from oracle import gmgan_cifar10 as OM
        model = OM.GMGANCifar10(params, dtype=torch.float32)
        return model, (lambda step: OM.synthetic_inputs(self.per_rank, step))


class Face(object):
    """BASELINE.json configs[3]: gan_inference_face.py:36,100-176"""
    name = "face"
    metric = "images/sec gan_inference_face ALI 64x64x3, global bs=128 sharded over the GPUs (1 iteration = G step + D step)"
    unit = "images/sec"
    scaling = "strong"
    gf_per_iter_per_unit = 187.2 / 128
    dom_name = "gg_conv2d_fwd (2B/N)x32x32x32->64 5x5 s2, Discriminator.2 on the batched fake+real towers"

    def __init__(self, world, rank):
        if 128 % world:
            raise SystemExit("--config face: 128 images do not shard over %d ranks" % world)
        self.per_rank, self.world, self.rank = 128 // world, world, rank
        self.units_per_iter = 128
        self.dom = (2 * self.per_rank, 32, 32, 32, 64)
        self.workload = "gan_inference_face.py MODE=ali global bs=128 (%d per GPU), 64x64x3 (BASELINE.json configs[3])" % self.per_rank

    def build(self):
        import gan_inference_face as S
        self.g = S.build_graph(BATCH_SIZE=self.per_rank)
        rs = np.random.RandomState(200 + self.rank)
        self.ring = [rs.randint(0, 256, size=(self.per_rank, 12288)).astype(np.uint8) for _ in range(4)]  # tflib/celebA.py dtype
        return self.g

    def host_feeds(self, j):
        return {self.g.real_x_int: self.ring[j % 4]}

    def device_feeds(self, torch):
        self.ring_d = [torch.from_numpy(a.astype(np.int32)).cuda() for a in self.ring]
        return lambda j: {self.g.real_x_int: self.ring_d[j % 4]}

    def oracle(self, params, torch):
        from oracle import gan_face as OM
        model = OM.GANFace(params, dtype=torch.float32)
        return model, (lambda step: OM.synthetic_inputs(self.per_rank, step))


class SSGAN(object):
    """BASELINE.json configs[4]: ssgan_inference_moving_mnist.py:42,49 with LEN 8, bs 32"""
    name = "ssgan"
    LEN = 8
    metric = "frames/sec ssgan_inference_moving_mnist LOCAL_EP LEN=8 1x64x64, global bs=32 sequences sharded over the GPUs"
    unit = "frames/sec"
    scaling = "strong"
    gf_per_iter_per_unit = None
    dom_name = "gg_conv2d_fwd (2*B*LEN/N)x32x32x32->64 5x5 s2, Discriminator.2 on the batched fake+real frame towers"

    def __init__(self, world, rank):
        if 32 % world:
            raise SystemExit("--config ssgan: 32 sequences do not shard over %d ranks" % world)
        self.per_rank, self.world, self.rank = 32 // world, world, rank
        self.units_per_iter = 32 * self.LEN
        self.dom = (2 * self.per_rank * self.LEN, 32, 32, 32, 64)
        self.workload = ("ssgan_inference_moving_mnist.py MODE=local_ep LEN=8 global bs=32 sequences (%d per GPU), 1x64x64 frames "
                         "(BASELINE.json configs[4])" % self.per_rank)

    def build(self):
        import ssgan_inference_moving_mnist as S
        self.g = S.build_graph(BATCH_SIZE=self.per_rank, LEN=self.LEN)
        rs = np.random.RandomState(300 + self.rank)
        self.ring = [(rs.uniform(0, 1, size=(self.per_rank, self.LEN, 4096)).astype(np.float32),
                      np.eye(10, dtype=np.float32)[rs.randint(0, 10, size=self.per_rank)]) for _ in range(4)]
        return self.g

    def host_feeds(self, j):
        x, y = self.ring[j % 4]
        return {self.g.real_x_unit: x, self.g.real_y: y}

    def device_feeds(self, torch):
        self.ring_d = [(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()) for x, y in self.ring]
        return lambda j: {self.g.real_x_unit: self.ring_d[j % 4][0], self.g.real_y: self.ring_d[j % 4][1]}

    def oracle(self, params, torch):
        from oracle import ssgan_moving_mnist as OM
        model = OM.SSGANMovingMNIST(params, self.per_rank, self.LEN, dtype=torch.float32)
        return model, (lambda step: OM.synthetic_inputs(self.per_rank, self.LEN, step))


WORKLOADS = {"cifar": Cifar, "face": Face, "ssgan": SSGAN}


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


def cpu_oracle_rate(wl, iters, warmup, threads=None):
    """the reference's algorithm (oracle port, PyTorch-CPU fp32, all host cores) on the same workload (this rank's shard)"""
    import torch
    import tflib as lib
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    if not lib._params:
        np.random.seed(1234)
        wl.build()
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    model, inputs = wl.oracle(params, torch)
    step, t0 = 0, None
    for it in range(warmup + iters):
        if it == warmup:
            t0 = time.perf_counter()
        model.gen_step(**inputs(step)); step += 1
        model.disc_step(**inputs(step)); step += 1
    dt = time.perf_counter() - t0
    units = wl.units_per_iter // wl.world
    return units * iters / dt, dt / iters, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.config](1, 0)
    iters = max(1, min(args.steps, 20 if args.config == "cifar" else 4))
    rate, sec_per, threads = cpu_oracle_rate(wl, iters, min(args.warmup, 3 if args.config == "cifar" else 1))
    line = {
        "impl": "reference", "metric": wl.metric, "value": rate, "unit": wl.unit, "n_gpus": args.gpus, "steps": iters,
        "warmup": min(args.warmup, 3), "ms_per_step": sec_per * 1e3, "higher_is_better": True, "scaling": wl.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.workload + ", CPU oracle port of the reference's tflib/TensorFlow path (TensorFlow itself "
                                            "cannot run here)"},
        "cpu_baseline": {"value": rate, "unit": wl.unit, "cores": threads, "kind": "port",
                         "sample": "%d iterations (G step + D step) of the same workload, PyTorch-CPU fp32" % iters},
        "e2e": {"value": rate, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_dominant_kernel(torch, cabi, flush_buf, dom):
    """live CUDA-event timing of the dominant tensor kernel: the tcgen05 conv2d forward at the Discriminator.2 shape of the
    step (fake and real towers batched).  N launches are replayed from a CUDA graph between two events, so the number is the
    kernel's device time (launch gap included), not the host's launch overhead; `cold` puts a 256 MiB memset between
    launches (its own time, measured the same way, subtracted).  The output of the timed launch is then checked against
    the fp32 direct backend on the same buffers."""
    B, H, W, Ci, Co = dom
    k, Ho, Wo = 5, H // 2, W // 2
    x = torch.randn(B, H, W, Ci, device="cuda")
    w = torch.randn(k, k, Ci, Co, device="cuda") * 0.05
    b = torch.randn(Co, device="cuda")
    y = torch.empty(B, Ho, Wo, Co, device="cuda")
    ws = torch.zeros(max(int(cabi.lib.gg_conv2d_workspace(0, B, H, W, Ci, Co, k, 2, Ho, Wo)), 256), dtype=torch.uint8, device="cuda")
    N = 20

    def launch(st, out=y):
        cabi.call("gg_conv2d_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, Ci, Co, k, 2, 1, 1, Ho, Wo,
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
    backend = cabi.lib.gg_last_backend()
    info = cabi.last_tc_info() if backend else {}
    # parity of the timed launch: same buffers through the fp32 direct kernels
    y2 = torch.empty_like(y)
    cabi.call("gg_set_conv_backend", 1)
    try:
        launch(torch.cuda.current_stream().cuda_stream, y2)
    finally:
        cabi.call("gg_set_conv_backend", 0)
    torch.cuda.synchronize()
    err = float((y - y2).abs().max() / y2.abs().max())
    if not err < 1e-3:
        raise SystemExit("bench.py: dominant kernel differs from the fp32 direct backend by %.3e of the tensor scale" % err)
    flops = 2.0 * B * Ho * Wo * Co * Ci * k * k
    return hot * 1e-3, cold * 1e-3, flops, backend, info, err


def measure_tf32_peak(torch):
    """cuBLAS fp32 GEMM with tf32 tensor-core math (the precision class of the conv kernel), 8192^3, best of 5 — the tf32
    denominator BASELINE.md §2 asks for next to the driver-measured bf16 figure"""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a, b = torch.randn(n, n, device="cuda"), torch.randn(n, n, device="cuda")
        (a @ b).sum().item()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/ncu_dominant_kernel.json, regenerated by tools/gpu_round.sh every round); None when absent"""
    path = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
    try:
        with open(path) as f:
            d = json.load(f)
        try:
            d["git_blob"] = subprocess.check_output(["git", "hash-object", path], cwd=ROOT, text=True, stderr=subprocess.DEVNULL).strip()
        except Exception:
            d["git_blob"] = None
        return d
    except Exception:
        return None


def run_ours(args):
    import torch
    from gg import cabi, dist as ggdist
    from gg.executor import RT
    import tensorflow as tf

    rank, world = ggdist.init_from_env()
    if world == 1 and torch.cuda.is_available():
        torch.cuda.set_device(0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    wl = WORKLOADS[args.config](world, rank)
    np.random.seed(1234)                      # identical initial weights on every rank
    g = wl.build()
    sess = tf.Session()
    dev_feeds = wl.device_feeds(torch)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def iteration_async(i):
        RT.run([g.gen_cost, g.gen_train_op], dev_feeds(2 * i), to_host=False)
        return RT.run([g.disc_cost, g.disc_train_op], dev_feeds(2 * i + 1), to_host=False)

    # End-to-end loop = the reference's training loop (gmgan_inference_cifar10.py:483-494): two session.run calls per iteration,
    # host batches in, the cost scalars out.  The session is opened with deferred_fetches=True: every run still enqueues the
    # host->device copy of its batch and the device->host copy of its cost, but the host turns a cost into a float one run
    # LATER (the reference only logs it), so feeding run k+1 overlaps the execution of run k.  GG_E2E_SYNC=1: wait every run.
    e2e_sync = os.environ.get("GG_E2E_SYNC", "0") == "1"
    sess_e2e = tf.Session(deferred_fetches=not e2e_sync)
    pending = []

    def settle(keep):
        while len(pending) > keep:
            float(pending.pop(0))

    def iteration_e2e(i):
        gc, _ = sess_e2e.run([g.gen_cost, g.gen_train_op], feed_dict=wl.host_feeds(2 * i))
        pending.append(gc)
        settle(1)
        dc, _ = sess_e2e.run([g.disc_cost, g.disc_train_op], feed_dict=wl.host_feeds(2 * i + 1))
        pending.append(dc)
        settle(1)
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
            print(json.dumps({"quick": True, "config": wl.name, "units_per_sec": wl.units_per_iter * args.steps / (dev_ms / 1e3),
                              "ms_per_step": dev_ms / args.steps, "ms_per_step_hot_l2": hot_ms / args.steps,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("GG_")}}))
        return
    # ---- end to end through Session.run: host batches in, cost scalars out, every step ----
    for i in range(3):
        iteration_e2e(i)
    barrier()
    plans = list(RT.plans.values())
    h2d0 = sum(getattr(p, "h2d_bytes", 0) for p in plans)
    d2h0 = sum(getattr(p, "d2h_bytes", 0) for p in plans)
    t0 = time.perf_counter()
    last = None
    for i in range(args.steps):
        last = iteration_e2e(i)
    settle(0)                                 # every cost of every run has been read on the host inside the timed region
    barrier()
    e2e_s = time.perf_counter() - t0
    last = (float(last[0]), float(last[1]))
    h2d_per_step = (sum(getattr(p, "h2d_bytes", 0) for p in plans) - h2d0) / float(args.steps)
    d2h_per_step = (sum(getattr(p, "d2h_bytes", 0) for p in plans) - d2h0) / float(args.steps)
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([dev_ms, hot_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    # replica consistency: every rank must hold bit-identical parameters after the timed steps (data-parallel correctness)
    checksum = torch.zeros(1, dtype=torch.float64, device="cuda")
    for t in RT.params.values():
        checksum += t.double().sum()
    cs_all = [checksum.clone() for _ in range(world)]
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_gather(cs_all, checksum)
    dev_ms, hot_ms, e2e_ms = [float(v) for v in times.cpu()]
    cs = [float(c.item()) for c in cs_all]
    if rank != 0:
        return
    if max(cs) - min(cs) != 0.0:
        raise SystemExit("bench.py: data-parallel replicas diverged: parameter checksums %r" % (cs,))

    gplan = RT.plans[[k for k in RT.plans if k[0][0] == g.gen_cost.id][0]]
    dplan = RT.plans[[k for k in RT.plans if k[0][0] == g.disc_cost.id][0]]
    launches_per_iter = int(gplan.kernel_launches + dplan.kernel_launches)
    units = wl.units_per_iter * args.steps
    value = units / (dev_ms / 1e3)
    hbm_peak, tf_peak, peak_kind = load_peaks()
    k_ms_hot, k_ms, k_flops, k_backend, k_info, k_err = time_dominant_kernel(torch, cabi, flush_buf, wl.dom)
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    tf32_peak = measure_tf32_peak(torch)
    ncu = load_ncu_traffic() if wl.name == "cifar" else None
    cpu = None
    if world == 1:
        n_it = 8 if wl.name == "cifar" else 2
        rate, _, threads = cpu_oracle_rate(wl, n_it, 2 if wl.name == "cifar" else 1)
        cpu = {"value": rate, "unit": wl.unit, "cores": threads, "kind": "port",
               "sample": "%d iterations (G step + D step) of the same workload on the host cores, PyTorch-CPU fp32" % n_it}
    line = {
        "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": wl.scaling,
        "vs_baseline": None, "dtype": "f32 storage, tf32 tensor-core operands / f32 accumulate on the 5x5 convs",
        "data": "synthetic",
        "config": {"workload": wl.workload,
                   "global_batch": wl.per_rank * world, "parallelism": "dp%d" % world,
                   "l2": "256 MiB memset between timed steps, outside the per-step CUDA-event brackets",
                   "ms_per_step_hot_l2": hot_ms / args.steps, "ms_per_step_p10_p50_p90_rank0": pct,
                   "cuda_graph": bool(RT.use_cuda_graph), "sync_bn": bool(RT.sync_bn),
                   "algorithmic_gflop_per_iteration": (wl.gf_per_iter_per_unit * wl.units_per_iter) if wl.gf_per_iter_per_unit else None,
                   "last_costs": [float(np.sum(last[0])), float(np.sum(last[1]))],
                   "replica_checksum": {"max_minus_min": max(cs) - min(cs), "value": cs[0], "ranks": len(cs)}},
        "e2e": {"value": units / (e2e_ms / 1e3), "unit": wl.unit,
                "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": int(d2h_per_step),
                "fetch": "every run waits for its cost" if e2e_sync else
                         "deferred by one run: each run enqueues its H2D batch copy and the D2H copy of its cost; the host reads "
                         "the cost while the next run executes (all costs read inside the timed region)"},
        "gpu_launches": launches_per_iter * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                     "traffic": (ncu or {}).get("dram_bytes_per_launch"),
                     "kernel": "%s (%s)" % (wl.dom_name, "tcgen05 tf32" if k_backend else "direct fp32"),
                     "launch_config": k_info,
                     "algorithmic_flop_per_launch": k_flops, "kernel_ms": k_ms, "kernel_ms_hot_l2": k_ms_hot,
                     "kernel_check_max_err_vs_fp32_backend": k_err,
                     "timing": "CUDA graph of 20 launches between two events, 256 MiB memset between launches (subtracted)",
                     "peak_kind": peak_kind + " dense bf16 burst (kind::tf32 peaks at half of it)",
                     "tf32_tflops_measured_cublas": tf32_peak, "frac_of_measured_tf32": achieved / tf32_peak,
                     "ncu": (ncu or {}).get("source"), "ncu_git_blob": (ncu or {}).get("git_blob")},
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        # leave NCCL before the interpreter tears the CUDA graphs down (captured collectives keep the communicator alive)
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cifar", choices=sorted(WORKLOADS))
    ap.add_argument("--quick", action="store_true", help="device-timed throughput only (knob sweeps); prints a short JSON")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
