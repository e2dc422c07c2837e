"""CPU: the element-wise cluster fusion of gg/fuse.py.  Every cluster program of every model family's training plans is evaluated
with NumPy (float64) on the values the graph interpreter computes for the cluster's inputs, and must reproduce the
interpreter's values of the nodes it replaces; the re-sorted plan must still be a topological order; the ctypes struct handed to
gg_ew_run must say what the description says.  (The kernel itself is checked on the GPU: tests/test_gpu_fusion.py.)"""
import os

import numpy as np
import pytest
import torch

import graph_interp as GI
from graph_interp import Interp
from test_cpu_graph import _families

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def eval_program(desc, vals):
    """NumPy evaluation of a cluster description: {node id: array} -> list of output arrays (output 0 reduced if desc says so)"""
    from gg import cabi
    un = {v: k for k, v in cabi.UNARY.items()}
    bi = {v: k for k, v in cabi.BINARY.items()}
    dims = desc["dims"]
    n = int(np.prod(dims))
    idx = np.arange(n)
    i3 = idx % dims[3]
    r = idx // dims[3]
    i2 = r % dims[2]
    r //= dims[2]
    i1 = r % dims[1]
    i0 = r // dims[1]
    regs = {}
    for k, ld in enumerate(desc["loads"]):
        s = ld["strides"]
        off = i0 * s[0] + i1 * s[1] + i2 * s[2] + i3 * s[3]
        if desc["flat"]:
            assert np.array_equal(off, idx), "flat program with a non-contiguous load"
        arr = np.asarray(vals[ld["node"].id], dtype=np.float64).reshape(-1)
        assert off.max() < arr.size
        regs[k] = torch.from_numpy(arr[off])
    for q in desc["instrs"]:
        if q["kind"] == 0:
            v = GI.UNARY[un[q["op"]]](regs[q["rsrc0"]], q["a"], q["b"])
        else:
            a, b = regs[q["rsrc0"]], regs[q["rsrc1"]]
            v = GI.BINARY[bi[q["op"]]](a, b, q["a"])
        regs[q["rdst"]] = v.clone()
    outs = [regs[rg].numpy() for rg in desc["out_regs"]]
    if desc["reduce"]:
        red = desc["reduce"]["red"]
        assert dims[3] == red
        m = outs[0].reshape(-1, red)
        outs[0] = {1: m.sum(1), 2: m.mean(1), 3: m.max(1)}[desc["reduce"]["op"]]
    return outs


def _feeds_for(nodes, rs):
    feeds = {}
    for n in sorted(nodes, key=lambda n: n.id):
        if n.op == "placeholder" or (n.op == "random" and n.attrs["kind"] != "categorical"):
            if n.dtype.name == "int32":
                feeds[n] = rs.randint(0, 10 if n.size < 4096 else 256, size=tuple(n.shape))
            elif n.op == "placeholder" or n.attrs["kind"] == "uniform":
                feeds[n] = rs.uniform(0.05, 0.95, size=tuple(n.shape))
            else:
                feeds[n] = rs.randn(*n.shape)
        elif n.op == "random":
            feeds[n] = rs.randint(0, n.inputs[0].size, size=tuple(n.shape))
    return feeds


def _check_plan(plan, it):
    from gg import fuse
    pos = {n.id: i for i, n in enumerate(plan.order)}
    # the re-sorted order is topological (and the mask of a fused dgrad launch precedes it)
    for n in plan.order:
        if n.id in plan.fed:
            continue
        for i in n.inputs:
            assert pos[i.id] < pos[n.id], "plan order broken at %s <- %s" % (n, i)
    for gid, (y, _a, _al) in plan.fuse_mask.items():
        assert pos[y.id] < pos[gid]
    n_checked = 0
    for cl in plan.ew_clusters:
        d = cl.desc
        # members of a cluster are contiguous in the plan and every external input comes before them
        ps = sorted(pos[n.id] for n in cl.nodes())
        assert ps == list(range(ps[0], ps[0] + len(ps))), "cluster is not contiguous in the plan"
        for ld in d["loads"]:
            assert pos[ld["node"].id] < ps[0]
            assert plan.buf[ld["node"].id] is not None, "cluster input %s has no buffer" % ld["node"]
        vals = {ld["node"].id: it.eval(ld["node"]).numpy() for ld in d["loads"]}
        outs = eval_program(d, vals)
        assert len(outs) == len(d["out_nodes"])
        for k, (m, got) in enumerate(zip(d["out_nodes"], outs)):
            target = cl.reduce if (d["reduce"] and k == 0) else m
            ref = it.eval(target).numpy().reshape(-1)
            scale = np.abs(ref).max() + 1e-30
            assert got.shape == ref.shape, (target, got.shape, ref.shape)
            assert np.abs(got - ref).max() <= 1e-12 * scale + 1e-300, "cluster output %s differs from the graph's value" % target
            assert plan.buf[target.id] is not None and plan.buf[target.id].numel() >= max(target.size, 1)
        # interior values own no buffer and nobody outside the cluster reads them
        inside = set(n.id for n in cl.nodes())
        for m in d["interior"]:
            if d["reduce"] and m is cl.members[-1]:
                continue
            for c in plan.order:
                if c.id not in inside and c.id not in plan.fed:
                    assert all(i is not m for i in c.inputs), "%s reads the register-only value %s" % (c, m)
        # the struct handed to the library
        st = fuse.to_struct(d, [0x1000 * (k + 1) for k in range(len(d["loads"]))], [0x100000 * (k + 1) for k in range(len(outs))])
        assert st.n_in == len(d["loads"]) and st.n_out == len(outs) and st.n_instr == len(d["instrs"])
        assert list(st.dims) == list(d["dims"]) and st.flat == int(d["flat"])
        for j, q in enumerate(d["instrs"]):
            assert (st.instr[j].dst, st.instr[j].src0, st.instr[j].kind, st.instr[j].op) == (q["rdst"], q["rsrc0"], q["kind"], q["op"])
            assert max(q["rdst"], q["rsrc0"], q["rsrc1"]) < 32
        n_checked += 1
    return n_checked


@pytest.mark.parametrize("family", ["gmgan_cifar10_local_ep", "gmgan_mnist_local_ep", "gmgan_svhn_local_epce", "gmgan_face_local_ep",
                                    "gan_svhn_wali_gp", "gan_face_ali", "gan_mnist_ali_bn_in_critic", "gan_cifar10_wali_gp",
                                    "ssgan_chairs", "gmgan_cifar10_reinforce", "ssgan_moving_mnist"])
def test_cluster_programs_reproduce_the_graph(cpu_device, family):
    import tensorflow as tf
    import tflib as lib
    from gg.executor import RT, Plan
    from gg.ops import toposort
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(11)
    g = _families()[family]()
    total = 0
    for cost, op in ((g.gen_cost, g.gen_train_op), (g.disc_cost, g.disc_train_op)):
        roots = [cost] + [d for d in op.deps if d is not None]
        plan = Plan(RT, [cost, op], [n for n in toposort(roots) if n.op == "placeholder"])
        fed = _feeds_for(plan.order, np.random.RandomState(5))
        it = Interp(fed)
        for n in plan.order:            # evaluate in one fixed order so that aux values (batch statistics) exist
            it.eval(n)
        total += _check_plan(plan, it)
    assert total >= 4, "no element-wise clusters were formed for %s" % family


def test_fusion_cuts_the_launch_list_and_can_be_switched_off(cpu_device, monkeypatch):
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    from gg.executor import RT, Plan

    def steps():
        tf.reset_default_graph()
        lib.delete_all_params()
        np.random.seed(1234)
        g = S.build_graph(BATCH_SIZE=64)
        gp = Plan(RT, [g.gen_cost, g.gen_train_op], [g.real_x_int])
        dp = Plan(RT, [g.disc_cost, g.disc_train_op], [g.real_x_int])
        return len(gp.steps) + len(dp.steps), gp, dp
    fused, gp, dp = steps()
    monkeypatch.setenv("GG_FUSE_EW", "0")
    plain, gp0, dp0 = steps()
    assert not gp0.ew_clusters and not dp0.ew_clusters
    assert fused <= plain - 55, (fused, plain)
    names = [r[0] for r in cpu_device]
    # the Gumbel-softmax chain -log(-log(U + eps) + eps) .. / temperature is one launch in front of the softmax
    chain = [cl for cl in gp.ew_clusters if sum(1 for q in cl.desc["instrs"] if q["kind"] == 0 and q["op"] == 6) >= 2]
    assert chain and len(chain[0].desc["instrs"]) >= 8
    # the four sigmoid-cross-entropy means of LOCAL_EP each run as program + row reduction
    assert sum(1 for cl in gp.ew_clusters if cl.reduce is not None and cl.desc["reduce"]["op"] == 2) >= 4


# ---- launch-list peepholes (gg/executor.py): structure of the compiled gmgan-CIFAR step ------------------------------------
def _cifar_plans(batch=64):
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    from gg.executor import RT, Plan
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=batch)
    return g, Plan(RT, [g.gen_cost, g.gen_train_op], [g.real_x_int]), Plan(RT, [g.disc_cost, g.disc_train_op], [g.real_x_int])


def _calls(plan, record):
    del record[:]
    for f in plan.steps:
        f(0)
    return list(record)


def test_strided_transposes_replace_slice_copy_and_activation_gradient(cpu_device):
    """`tf.reshape(output, [-1, 4*4*4*DIM])` -> `tf.concat([output, z_output], 1)` (gmgan_inference_cifar10.py:292-294): the forward
    transpose writes into the concat's columns, the backward one reads the dense gradient's columns and applies LeakyReLU'"""
    from gg import cabi
    g, gplan, dplan = _cifar_plans()
    for plan in (gplan, dplan):
        calls = _calls(plan, cpu_device)
        ex = [a for n, a in calls if n == "gg_transpose_b2d_ex"]
        assert len(ex) == 2
        fwd = [a for a in ex if a[7] is None]
        bwd = [a for a in ex if a[7] is not None]
        assert len(fwd) == 1 and len(bwd) == 1
        # forward: [128, 4*4, 256] -> [128, 256, 4*4], contiguous in, rows of the [128, 4096 + 512] concat out
        assert fwd[0][2:7] == (128, 16, 256, 4096, 4608)
        # backward: columns 0..4095 of the [128, 4608] gradient in, contiguous out, mask = the conv output (leaky 0.2)
        assert bwd[0][2:7] == (128, 256, 16, 4608, 4096) and bwd[0][8] == cabi.ACT["leaky"] and abs(bwd[0][9] - 0.2) < 1e-7
        info = [i for i in plan.tr_fuse.values() if i["y_concat"] is not None][0]
        store = plan._concat_storage(info["y_concat"].id)
        assert fwd[0][1] == store.data_ptr() + 4 * info["y_off"] and plan.buf[info["y_concat"].id].data_ptr() == store.data_ptr()
        binfo = [i for i in plan.tr_fuse.values() if i["x_node"] is not None][0]
        assert bwd[0][0] == plan.buf[binfo["x_node"].id].data_ptr() + 4 * binfo["x_off"]
        assert bwd[0][7] == plan.buf[binfo["mask"][0].id].data_ptr()
        # the slice in front of the backward transpose owns no launch and no buffer; the concat copies only its other piece
        assert plan.buf[binfo["slice"].id] is None
        cat = info["y_concat"]
        copies = [a for n, a in calls if n == "gg_copy2d" and store.data_ptr() <= a[2] < store.data_ptr() + 4 * cat.size]
        assert len(copies) == len(cat.inputs) - 1


def test_prior_sample_is_a_gather_with_the_noise_added_and_the_tick_comes_last(cpu_device):
    g, gplan, dplan = _cifar_plans()
    for plan in (gplan, dplan):
        calls = _calls(plan, cpu_device)
        names = [n for n, _ in calls]
        assert names.count("gg_gather_rows") == 1
        ga = [a for n, a in calls if n == "gg_gather_rows"][0]
        (mid, (add_node, noise)), = plan.gather_add.items()
        m = [n for n in plan.order if n.id == mid][0]
        idx = m.inputs[0].inputs[0]
        assert ga[0] == plan.buf[idx.id].data_ptr() and ga[1] == plan.buf[m.inputs[1].id].data_ptr()
        assert ga[2] == plan.buf[noise.id].data_ptr() and ga[3] == plan.buf[add_node.id].data_ptr()
        assert ga[4:7] == (64, 128, 30)
        # the launch waits for the index and the noise, not for the one-hot matrix
        pos = {n: i for i, (n, _) in enumerate(calls)}
        assert pos["gg_rng_categorical"] < pos["gg_gather_rows"] and pos["gg_rng_normal"] < pos["gg_gather_rows"]
        grp = [q for q in plan.groups if q.get("node") is m][0]
        assert m.inputs[0].id not in grp["reads"] and idx.id in grp["reads"]
        # every random draw of the step happens before the counter is advanced
        last_rng = max(i for i, n in enumerate(names) if n.startswith("gg_rng_") and n != "gg_rng_tick")
        assert names.count("gg_rng_tick") == 1 and names.index("gg_rng_tick") > last_rng


def test_dense_activation_gradients_and_rank1_products_leave_the_launch_list(cpu_device, monkeypatch):
    g, gplan, dplan = _cifar_plans()
    for plan in (gplan, dplan):
        calls = _calls(plan, cpu_device)
        dense = [a for n, a in calls if n == "gg_conv2d_dgrad_actgrad" and a[7:9] == (1, 1)]
        assert len(dense) >= 2 and all(a[4] != 0 for a in dense)
        # no K = 1 GEMM is left: dy [128,1] x W^T [1,512] is an element-wise multiply inside a cluster
        assert not [a for n, a in calls if n == "gg_gemm" and a[6] == 1]


def test_batchnorm_statistics_buffer_is_shared_between_producer_and_apply(cpu_device):
    g, gplan, dplan = _cifar_plans()
    for plan in (gplan, dplan):
        calls = _calls(plan, cpu_device)
        prod = {a[5]: a for n, a in calls if n == "gg_conv2d_bnstats"}
        app = [a for n, a in calls if n == "gg_bn_apply"]
        assert len(prod) == 5 and len(app) == 5
        from gg import cabi
        for a in app:
            assert a[1] in prod, "gg_bn_apply reads a statistics buffer no launch wrote"
            p = prod[a[1]]
            assert a[2] == cabi.lib.gg_conv2d_stats_tiles(p[0], *p[6:17])            # S = m-tiles of the producer
            assert a[0] == p[4], "the batch norm does not read the producer's output"
            R, C = a[10], a[11]
            assert abs(a[3] - float(R)) < 0.5 and C == (p[10] if p[0] == 0 else p[9])  # channels = Co (fwd) / Ci (dgrad)
        order = [n for n, _ in calls]
        for ptr, p in prod.items():
            i_prod = [i for i, (n, a) in enumerate(calls) if n == "gg_conv2d_bnstats" and a[5] == ptr][0]
            i_app = [i for i, (n, a) in enumerate(calls) if n == "gg_bn_apply" and a[1] == ptr][0]
            assert i_prod < i_app


def test_no_high_priority_group_waits_for_a_low_priority_one(cpu_device, monkeypatch):
    """GG_PRIO_TAIL: on (auto) for the multi-wave face plans, off for the latency-bound batch-64 cifar plans"""
    import tensorflow as tf
    import tflib as lib
    import gan_inference_face as F
    from gg.executor import RT, Plan
    g, gplan, dplan = _cifar_plans()
    for plan in (gplan, dplan):
        plan._schedule_range(range(len(plan.groups)), 8)
        assert plan.prio_tail_mode == "0"
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(2)
    gf = F.build_graph(BATCH_SIZE=128)
    for fetch in ([gf.gen_cost, gf.gen_train_op], [gf.disc_cost, gf.disc_train_op]):
        plan = Plan(RT, fetch, [gf.real_x_int])
        plan._schedule_range(range(len(plan.groups)), 8)
        assert plan.prio_tail_mode == "1"
        producer = {}
        for gi, grp in enumerate(plan.groups):
            for o in grp["reads"]:
                d = producer.get(o)
                if d is not None and not grp["barrier"] and not plan.low_class[gi]:
                    assert not plan.low_class[d], "high-priority group %d waits for low-priority group %d" % (gi, d)
            producer[grp["writes"]] = gi
        assert any(plan.low_class.values()), "nothing is left in the low class"
