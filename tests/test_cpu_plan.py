"""CPU: the plan compiler of gg/executor.py run against a host "device" (buffers are host tensors, launches are recorded,
nothing executes): launch-list structure, the static multi-stream schedule, zero-copy row concat / row slices, the
data-parallel launch-list cuts.  No kernel runs here; the numbers are checked on the GPU by tests/test_gpu_*.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


@pytest.fixture()
def cpu_device(monkeypatch):
    from gg import cabi, executor
    monkeypatch.setattr(executor.Runtime, "dev", lambda self: torch.device("cpu"))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    record = []
    real = cabi.call

    def fake(name, *args):
        if name in ("gg_set_tc_max_ctas", "gg_set_tc_stages", "gg_set_pdl"):
            return real(name, *args)
        record.append((name, args))
    monkeypatch.setattr(cabi, "call", fake)
    executor.reset_runtime()
    yield record
    executor.reset_runtime()


def _gmgan(batch=64):
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    return S.build_graph(BATCH_SIZE=batch)


def _plans(g):
    from gg.executor import RT, Plan
    return (Plan(RT, [g.gen_cost, g.gen_train_op], [g.real_x_int]), Plan(RT, [g.disc_cost, g.disc_train_op], [g.real_x_int]))


def _group_deps(plan):
    producer, deps, last_barrier, seen = {}, {}, None, []
    for gi, g in enumerate(plan.groups):
        d = set(producer[o] for o in g["reads"] if o in producer)
        if last_barrier is not None:
            d.add(last_barrier)
        if g["barrier"]:
            d |= set(seen)
            last_barrier = gi
        deps[gi] = d
        producer[g["writes"]] = gi
        seen.append(gi)
    return deps


@pytest.mark.parametrize("n_streams", [1, 2, 6, 12])
def test_schedule_respects_every_dependency(cpu_device, n_streams):
    """issue order is topological; every producer of a group is either earlier on the SAME stream or covered by an event wait
    on a group at or after it on ITS stream; waited events are recorded before the wait is issued"""
    for plan in _plans(_gmgan()):
        order, assign, waits, need_event = plan._schedule_range(range(len(plan.groups)), n_streams)
        assert sorted(order) == list(range(len(plan.groups)))
        pos = {gi: i for i, gi in enumerate(order)}
        deps = _group_deps(plan)
        assert all(0 <= assign[g] < n_streams for g in order)
        for gi in order:
            for w in waits[gi]:
                assert pos[w] < pos[gi] and w in need_event and assign[w] != assign[gi]
            for d in deps[gi]:
                assert pos[d] < pos[gi], "group %d issued before its producer %d" % (gi, d)
                if assign[d] == assign[gi]:
                    continue                                   # stream order
                covered = any(assign[w] == assign[d] and pos[w] >= pos[d] for w in waits[gi])
                # ... or transitively: an earlier group on MY stream already waited for it
                if not covered:
                    mine = [x for x in order[:pos[gi]] if assign[x] == assign[gi]]
                    covered = any(assign[w] == assign[d] and pos[w] >= pos[d] for x in mine for w in waits[x])
                assert covered, "dependency %d -> %d crosses streams without an event" % (d, gi)
        if n_streams > 1:
            assert 200.0 < plan.sched_estimate_us < 2000.0


def test_heft_schedule_is_shorter_than_plan_order(cpu_device, monkeypatch):
    """the simulated makespan of the list schedule is close to the critical path and well below the serial sum"""
    gplan, dplan = _plans(_gmgan())
    for plan in (gplan, dplan):
        plan._schedule_range(range(len(plan.groups)), 6)
        serial = sum(plan._group_cost(g) for g in plan.groups)
        assert plan.sched_estimate_us < 0.6 * serial


def test_sibling_batching_halves_discriminator_launches_and_concat_is_zero_copy(cpu_device, monkeypatch):
    g = _gmgan()
    gplan, dplan = _plans(g)

    def heavy(plan):
        return [n for n in plan.order if n.op in ("conv", "matmul")]
    monkeypatch.setenv("GG_BATCH_SIBLINGS", "0")
    g0 = _gmgan()
    gplan0, dplan0 = _plans(g0)
    assert len(heavy(dplan)) <= len(heavy(dplan0)) - 20 and len(heavy(gplan)) <= len(heavy(gplan0)) - 12
    monkeypatch.delenv("GG_BATCH_SIBLINGS")
    # every batched conv runs on 2x the rows; row concats of the towers' inputs own no copy kernels
    assert any(n.op == "conv" and n.attrs["B"] == 128 for n in dplan.order)
    for plan in (gplan, dplan):
        assert plan.inplace_concat, "no zero-copy concat in the plan"
        for node in plan.order:
            if node.id in plan.inplace_concat:
                base = plan.buf[node.id]
                off = 0
                for inp in node.inputs:
                    piece = plan.buf[inp.id]
                    assert piece.data_ptr() == base.data_ptr() + 4 * off, "concat piece is not placed inside the concat buffer"
                    off += inp.size
            if node.op == "slice" and plan._slice_is_view(node) and node.id not in plan.placed:
                src = plan.buf[node.inputs[0].id]
                inner = int(np.prod(node.inputs[0].shape[node.attrs["axis"] + 1:], dtype=np.int64))
                assert plan.buf[node.id].data_ptr() == src.data_ptr() + 4 * node.attrs["start"] * inner


def test_launch_list_uses_single_launch_batchnorm_and_fused_epilogues(cpu_device, monkeypatch):
    record = cpu_device
    gplan, _ = _plans(_gmgan())
    del record[:]
    for f in gplan.steps:
        f(0)
    names = [n for n, _ in record]
    # every batch norm of the step sits behind a tensor-core conv / deconv / dense launch: that launch's epilogue takes the
    # statistics (gg_conv2d_bnstats) and the batch norm is one element-wise pass (gg_bn_apply) — no moments pass at all
    assert names.count("gg_conv2d_bnstats") == 5 and names.count("gg_bn_apply") == 5 and "gg_bn_fwd_fused" not in names
    assert "gg_bn_bwd_fused" in names and "gg_bn_stats" not in names
    assert names.count("gg_adam_multi") == 1 and names.count("gg_rng_tick") == 1
    assert len(names) < 115, "G-step launch list grew to %d C-ABI calls" % len(names)
    for n, a in record:
        if n == "gg_conv2d_bnstats":
            assert a[5] is not None and a[5] != 0, "statistics buffer missing"
    # without the epilogue statistics: the one-launch cluster batch norm
    monkeypatch.setenv("GG_BN_CONV_STATS", "0")
    gplan0, _ = _plans(_gmgan())
    del record[:]
    for f in gplan0.steps:
        f(0)
    names0 = [n for n, _ in record]
    assert names0.count("gg_bn_fwd_fused") == 5 and "gg_conv2d_bnstats" not in names0 and "gg_bn_apply" not in names0
    monkeypatch.delenv("GG_BN_CONV_STATS")
    del record[:]
    for f in gplan.steps:
        f(0)
    # LeakyReLU / ReLU / tanh never appear as their own launches: they ride conv / dense / BN epilogues
    from gg import cabi
    act_codes = {cabi.UNARY[k] for k in ("relu", "leaky", "tanh", "sigmoid")}
    assert not [a for n, a in record if n == "gg_unary" and a[0] in act_codes]


class _FakeArena(object):
    """host stand-in of gg.dist.PeerArena: bump allocation only"""
    peers = None

    def __init__(self):
        self.off, self.sites = 0, []

    def alloc(self, nbytes):
        off = self.off
        self.off += (int(nbytes) + 127) & ~127
        self.sites.append((off, int(nbytes)))
        return off


def test_data_parallel_plan_buckets_gradients_by_readiness_and_fuses_syncbn(cpu_device, monkeypatch):
    """world 2: (a) every batch norm is ONE launch with the cross-rank exchange inside (gg_bn_*_fused_dp), each call site with
    its own arena region, chained in the same order on every rank; (b) parameter gradients are produced inside the
    optimiser's flat buffer (no pack kernel) in readiness order; (c) one NCCL all-reduce per bucket, each depending only on
    the kernels that produce its gradients, so the first bucket can overlap the rest of the backward pass; (d) the update is
    the barrier behind all buckets and divides by the world size"""
    from gg import dist as ggdist
    arena = _FakeArena()
    monkeypatch.setattr(ggdist, "world_size", lambda: 2)
    monkeypatch.setattr(ggdist, "rank", lambda: 1)
    monkeypatch.setattr(ggdist, "peer_arena", lambda: arena)
    monkeypatch.setenv("GG_DP_BUCKETS", "2")
    g = _gmgan(32)
    gplan, dplan = _plans(g)
    # default: the SyncBN kernels of this model are small grids (<= 148 CTAs) and run UNordered — chaining them would
    # serialise the generator's and the extractor's batch-norm chains
    assert not [grp for grp in gplan.groups if grp.get("ordered") == "peer"]
    monkeypatch.setenv("GG_BN_DP_ORDER", "1")
    g = _gmgan(32)
    gplan, dplan = _plans(g)
    names = [n for n, _ in cpu_device]
    assert "gg_pack_grads" not in names and "gg_bn_stats" not in names
    assert names.count("gg_bn_fwd_fused_dp") == 0          # launches are recorded at capture / run time, not at plan build
    # (a) arena regions: disjoint, one per BN forward / backward call site of the two plans
    assert len(arena.sites) >= 10
    ends = [o + n for o, n in arena.sites]
    assert all(arena.sites[i + 1][0] >= ends[i] for i in range(len(ends) - 1))
    for plan, n_params in ((gplan, None), (dplan, None)):
        coll = [gi for gi, grp in enumerate(plan.groups) if grp["collective"]]
        assert len(coll) == 2, "one NCCL all-reduce per readiness bucket"
        assert all(plan.groups[gi]["ordered"] == "nccl" and not plan.groups[gi]["barrier"] for gi in coll)
        peer = [gi for gi, grp in enumerate(plan.groups) if grp.get("ordered") == "peer"]
        if plan is gplan:
            assert len(peer) >= 8                           # 5 BN layers forward + backward in G / E
        # (b) most gradient bytes are produced in place
        (entries, buckets, total), = plan.bucket_plan.values()
        direct = sum(e["var"].size for e in entries if e["direct"])
        assert direct > 0.95 * sum(e["var"].size for e in entries), (direct, total)
        assert [e["off"] for e in entries] == sorted(e["off"] for e in entries)
        assert [e["ready"] for e in entries] == sorted(e["ready"] for e in entries)
        assert all(e["off"] % 64 == 0 for e in entries)
        for e in entries:
            if e["direct"]:
                buf, (flat,) = plan.buf[e["own"].id], plan.flat.values()
                assert buf.data_ptr() == flat.data_ptr() + 4 * e["off"]
        # (c) dependencies of the first bucket do not include the last backward kernels
        order, assign, waits, _ = plan._schedule_range(range(len(plan.groups)), 6)
        pos = {gi: i for i, gi in enumerate(order)}
        first, last = sorted(coll)[0], sorted(coll)[-1]
        barrier = [gi for gi, grp in enumerate(plan.groups) if grp["barrier"]][-1]
        assert pos[first] < pos[barrier] and pos[last] < pos[barrier]
        deps = _group_deps(plan)

        def closure(gi):
            seen, stack = set(), [gi]
            while stack:
                for d in deps[stack.pop()]:
                    if d not in seen:
                        seen.add(d)
                        stack.append(d)
            return seen
        c_first, c_last = closure(first), closure(last)
        assert len(c_last - c_first) > 5, "the first bucket's all-reduce does not wait for the kernels that only the second needs"
        # peer-exchange kernels keep one global order
        seq = [gi for gi in order if plan.groups[gi].get("ordered") == "peer"]
        assert seq == sorted(seq)


ALL_SCRIPTS = {
    "gmgan_inference_mnist": dict(BATCH_SIZE=4),
    "gmgan_inference_cifar10": dict(BATCH_SIZE=4),
    "gmgan_inference_cifar10:reinforce": dict(BATCH_SIZE=4, MODE_K='REINFORCE'),
    "gmgan_inference_cifar10:local_epce": dict(BATCH_SIZE=4, MODE='local_epce'),
    "gmgan_inference_svhn": dict(BATCH_SIZE=4),
    "gmgan_inference_face": dict(BATCH_SIZE=2, N_COMS=10),
    "gmgan_inference_face:ali": dict(BATCH_SIZE=2, N_COMS=10, MODE='ali'),
    "gan_inference_mnist": dict(BATCH_SIZE=4),
    "gan_inference_mnist:vegan-wgan-gp": dict(BATCH_SIZE=4, MODE='vegan-wgan-gp'),
    "gan_inference_cifar10:wali-gp": dict(BATCH_SIZE=4, MODE='wali-gp'),
    "gan_inference_cifar10:alice": dict(BATCH_SIZE=4, MODE='alice'),
    "gan_inference_svhn:wali": dict(BATCH_SIZE=4, MODE='wali'),
    "gan_inference_face": dict(BATCH_SIZE=2),
    # the discriminator-free VEGAN variants: generator objective only, CRITIC_ITERS = 0
    "gan_inference_mnist:vegan-kl": dict(BATCH_SIZE=4, MODE='vegan-kl', Z_SAMPLES=8),
    "gan_inference_mnist:vegan-jsd": dict(BATCH_SIZE=4, MODE='vegan-jsd', Z_SAMPLES=8),
    "gan_inference_svhn:vegan-ikl": dict(BATCH_SIZE=4, MODE='vegan-ikl', Z_SAMPLES=8),
    "gan_inference_cifar10:vegan-mmd": dict(BATCH_SIZE=4, MODE='vegan-mmd'),
    "ssgan_inference_moving_mnist": dict(BATCH_SIZE=2, LEN=4),
    # the SSGAN scripts' ALI mode: one critic on (clip, all latents) — Conv3D trunk / frames as channels / per-frame codes
    "ssgan_inference_moving_mnist:ali-3dcnn": dict(BATCH_SIZE=2, LEN=4, MODE='ali', ALI_MODE='3dcnn'),
    "ssgan_inference_moving_mnist:ali-3dcnn-bn": dict(BATCH_SIZE=2, LEN=4, MODE='ali', ALI_MODE='3dcnn', BN_FLAG=True),
    "ssgan_inference_moving_mnist:alice-z-concat_z": dict(BATCH_SIZE=2, LEN=4, MODE='alice-z', ALI_MODE='concat_z'),
    "ssgan_inference_moving_mnist:ali-concat_x": dict(BATCH_SIZE=2, LEN=4, MODE='ali', ALI_MODE='concat_x'),
    "ssgan_inference_chairs": dict(BATCH_SIZE=2, LEN=4),
    "ssgan_inference_chairs:local_epce-z": dict(BATCH_SIZE=2, LEN=3, MODE='local_epce-z'),
}


@pytest.mark.parametrize("script", sorted(ALL_SCRIPTS))
def test_every_script_compiles_to_a_launch_list(cpu_device, script):
    """all ten reference scripts (and their non-default modes) go through the plan compiler: every node has a launcher,
    every layout pattern fits the kernels' index limits, both train steps schedule on 6 streams"""
    import importlib
    import tensorflow as tf
    import tflib as lib
    from gg.executor import RT, Plan
    from gg.ops import toposort
    mod = importlib.import_module(script.split(":")[0])
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(2)
    g = mod.build_graph(**ALL_SCRIPTS[script])
    for cost, op in ((g.gen_cost, g.gen_train_op), (g.disc_cost, g.disc_train_op)):
        if op is None:                      # no-discriminator modes have only the generator step
            assert g.CRITIC_ITERS == 0 and cost is None
            continue
        roots = [cost] + [d for d in op.deps if d is not None]
        fed = [n for n in toposort(roots) if n.op == "placeholder"]
        plan = Plan(RT, [cost, op], fed)
        assert len(plan.steps) > 20 and plan.groups
        order, assign, waits, _ = plan._schedule_range(range(len(plan.groups)), 6)
        assert sorted(order) == list(range(len(plan.groups)))
        record = cpu_device
        del record[:]
        for f in plan.steps:
            f(0)
        names = [n for n, _ in record]
        whole = sum(n.startswith(("gg_adam_multi", "gg_rmsprop_multi")) for n in names)
        # one update per step: a single launch, or (Adam, split by gradient readiness) state advance + early + late launch
        assert (whole == 1 and "gg_adam_apply" not in names) or \
               (whole == 0 and names.count("gg_adam_tick") == 1 and names.count("gg_adam_apply") == 2)
        assert any(n in ("gg_conv2d_fwd", "gg_conv2d_dgrad") for n in names)


def test_split_update_waits_for_gradients_and_for_every_reader_of_its_parameters(cpu_device, monkeypatch):
    """The optimiser update is split by gradient readiness (executor._late_vars): the EARLY launch may run while the last
    backward kernels are still in flight, so it must be ordered after (a) the kernels producing the gradients it consumes and
    (b) the last kernel of every node that READS a parameter it overwrites; the LATE launch is the step's barrier."""
    monkeypatch.setenv("GG_SPLIT_UPDATE", "1")        # opt-in: measured slower on the B200 at N=1 (executor._late_vars)
    for plan in _plans(_gmgan()):
        names = {g["writes"]: gi for gi, g in enumerate(plan.groups)}
        early = [gi for w, gi in names.items() if isinstance(w, str) and w.endswith("_early")]
        tick = [gi for w, gi in names.items() if isinstance(w, str) and w.endswith("_tick") and w.startswith("op")]
        assert len(early) == 1 and len(tick) == 1
        early, tick = early[0], tick[0]
        deps = _group_deps(plan)
        closure, stack = set(), [early]
        while stack:
            for d in deps[stack.pop()]:
                if d not in closure:
                    closure.add(d)
                    stack.append(d)
        assert tick in closure and not plan.groups[early]["barrier"]
        barrier = [gi for gi, g in enumerate(plan.groups) if g["barrier"]]
        assert len(barrier) == 1 and early in deps[barrier[0]] and barrier[0] > early
        opt = [o for o in plan._optimizer_ops(plan.fetches)][0]
        pairs = [(v, g) for v, g in zip(opt.attrs["vars"], opt.deps) if g is not None]
        late = plan._late_vars(pairs)
        assert 0 < len(late) < len(pairs)
        late_bytes = sum(v.size for v, _ in pairs if v.id in late)
        assert late_bytes < 0.2 * sum(v.size for v, _ in pairs), "the barrier launch should only hold the last-ready gradients"
        early_vars = set(v.id for v, _ in pairs if v.id not in late)
        # (a) producers of the early gradients
        for v, g in pairs:
            if v.id in early_vars:
                for o in plan._owners(g):
                    if o in names:
                        assert names[o] in closure, "early update does not wait for the gradient of %s" % v.name
        # (b) every group reading an early parameter (first part carries the reads; the node's last part must be in the closure)
        n_readers = 0
        for gi, g in enumerate(plan.groups):
            if gi != early and g["part"][0] == 0 and (g["reads"] & early_vars):
                assert gi + g["part"][1] - 1 in closure, "early update may overwrite a parameter that %r still reads" % (g["writes"],)
                n_readers += 1
        assert n_readers > 10
