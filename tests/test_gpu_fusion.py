"""-m gpu: the fused element-wise program launches (gg_ew_run, gg/fuse.py) are BIT-IDENTICAL to the one-op launches they replace.

(1) hand-written programs through the C-ABI against gg_unary / gg_binary / gg_reduce chains: broadcast strides, int32 loads,
    several outputs, register reuse, the three thread-count tiers of the row reduction;
(2) whole training plans of every model family with GG_FUSE_EW=1 vs GG_FUSE_EW=0 on the same parameters, batches and injected
    noise: costs and every parameter after several iterations must be equal bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


def _prog(dims, loads, instrs, outs, flat, reduce_op=0):
    """loads: [(tensor, strides4, is_int)], instrs: [(kind, op, dst, src0, src1, a, b)], outs: [(tensor, reg)]"""
    from gg import cabi
    p = cabi.EwProgram()
    p.n_in, p.n_out, p.n_instr, p.flat, p.reduce_op = len(loads), len(outs), len(instrs), int(flat), reduce_op
    for i in range(4):
        p.dims[i] = dims[i]
    for k, (t, s, is_int) in enumerate(loads):
        p.inp[k] = t.data_ptr()
        p.in_is_int[k] = int(is_int)
        for i in range(4):
            p.in_stride[k][i] = s[i]
    for k, (t, r) in enumerate(outs):
        p.out[k] = t.data_ptr()
        p.out_reg[k] = r
    for j, (kind, op, dst, s0, s1, a, b) in enumerate(instrs):
        q = p.instr[j]
        q.kind, q.op, q.dst, q.src0, q.src1, q.a, q.b = kind, op, dst, s0, s1, a, b
    return p


def _unary(fn, x, a=0.0, b=0.0):
    from gg import cabi
    y = torch.empty_like(x)
    cabi.call("gg_unary", cabi.UNARY[fn], x.data_ptr(), y.data_ptr(), x.numel(), a, b, cabi.stream_ptr())
    return y


def _binary(fn, a, b, out_shape, sa, sb, alpha=0.0):
    from gg import cabi
    out = torch.empty(out_shape, device="cuda")
    d = list(out_shape)
    d = [1] * (4 - len(d)) + d
    cabi.call("gg_binary", cabi.BINARY[fn], a.data_ptr(), b.data_ptr(), out.data_ptr(), cabi.int4(d), cabi.int4(sa), cabi.int4(sb),
              alpha, cabi.stream_ptr())
    return out


def test_flat_chain_with_int_load_and_two_outputs():
    """the image decode 2*((float(x)/255)-.5) + noise, keeping an intermediate (gmgan_inference_cifar10.py:341-342)"""
    from gg import cabi
    U, B = cabi.UNARY, cabi.BINARY
    n = 64 * 3072 + 3
    xi = torch.randint(0, 256, (n,), dtype=torch.int32, device="cuda")
    noise = torch.rand(n, device="cuda")
    xf = torch.empty(n, device="cuda")
    cabi.call("gg_cast_i32_f32", xi.data_ptr(), xf.data_ptr(), n, 1.0, 0.0, cabi.stream_ptr())
    t1 = _unary("divc", xf, 255.0)
    t2 = _unary("affine", t1, 1.0, -0.5)
    t3 = _unary("affine", t2, 2.0, 0.0)
    t4 = _binary("add", t3, noise, (n,), [0, 0, 0, 1], [0, 0, 0, 1])
    t5 = _unary("tanh", t4)
    o3, o5 = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    p = _prog([1, 1, 1, n], [(xi, [0, 0, 0, 1], True), (noise, [0, 0, 0, 1], False)],
              [(0, U["divc"], 2, 0, 0, 255.0, 0.0), (0, U["affine"], 2, 2, 0, 1.0, -0.5), (0, U["affine"], 0, 2, 0, 2.0, 0.0),
               (1, B["add"], 2, 0, 1, 0.0, 0.0), (0, U["tanh"], 31, 2, 0, 0.0, 0.0)],
              [(o3, 0), (o5, 31)], flat=True)
    cabi.call("gg_ew_run", C.byref(p), cabi.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(o3, t3) and torch.equal(o5, t5)


def test_broadcast_strides_match_gg_binary():
    """(x[B,1,D] - mu[1,K,D])^2 scaled by a per-row weight w[B,K,1]: the mixture-prior distance (gmgan_inference_cifar10.py:155-160)"""
    from gg import cabi
    U, Bn = cabi.UNARY, cabi.BINARY
    Bt, K, D = 64, 30, 128
    x, mu, w = torch.randn(Bt, D, device="cuda"), torch.randn(K, D, device="cuda"), torch.rand(Bt, K, device="cuda")
    diff = _binary("sub", x, mu, (Bt, K, D), [0, D, 0, 1], [0, 0, D, 1])
    sq = _unary("square", diff)
    ref = _binary("mul", sq, w, (Bt, K, D), [0, K * D, D, 1], [0, K, 1, 0])
    o_diff, o = torch.empty(Bt, K, D, device="cuda"), torch.empty(Bt, K, D, device="cuda")
    p = _prog([1, Bt, K, D], [(x, [0, D, 0, 1], False), (mu, [0, 0, D, 1], False), (w, [0, K, 1, 0], False)],
              [(1, Bn["sub"], 3, 0, 1, 0.0, 0.0), (0, U["square"], 4, 3, 0, 0.0, 0.0), (1, Bn["mul"], 0, 4, 2, 0.0, 0.0)],
              [(o_diff, 3), (o, 0)], flat=False)
    cabi.call("gg_ew_run", C.byref(p), cabi.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(o_diff, diff) and torch.equal(o, ref)


@pytest.mark.parametrize("rows,red", [(1, 64), (7, 130), (3, 1500), (1920, 128), (1, 1)])
@pytest.mark.parametrize("fn", ["sum", "mean", "max"])
def test_row_reduction_matches_gg_reduce(rows, red, fn):
    """BCE(x, label) -> mean: tflib/objs/gan_inference.py:85-101; every thread-count tier of gg_reduce's row kernel"""
    from gg import cabi
    x = torch.randn(rows, red, device="cuda") * 3
    e = _unary("bce", x, 1.0)
    ref = torch.empty(rows, device="cuda")
    cabi.call("gg_reduce", cabi.REDUCE[fn], e.data_ptr(), ref.data_ptr(), rows, red, 1, cabi.stream_ptr())
    out, keep = torch.empty(rows, device="cuda"), torch.empty(rows, red, device="cuda")
    p = _prog([1, 1, rows, red], [(x, [0, 0, red, 1], False)], [(0, cabi.UNARY["bce"], 5, 0, 0, 1.0, 0.0)], [(out, 5), (keep, 5)],
              flat=True, reduce_op={"sum": 1, "mean": 2, "max": 3}[fn])
    cabi.call("gg_ew_run", C.byref(p), cabi.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out, ref), (out[:4], ref[:4])
    assert torch.equal(keep, e)


def test_bad_programs_are_refused():
    from gg import cabi
    x = torch.zeros(8, device="cuda")
    p = _prog([1, 1, 1, 8], [(x, [0, 0, 0, 1], False)], [(0, 99, 1, 0, 0, 0.0, 0.0)], [(x, 1)], flat=True)
    with pytest.raises(cabi.GGError):
        cabi.call("gg_ew_run", C.byref(p), cabi.stream_ptr())
    p = _prog([1, 1, 1, 8], [(x, [0, 0, 0, 1], False)], [(0, 0, 40, 0, 0, 0.0, 0.0)], [(x, 1)], flat=True)
    with pytest.raises(cabi.GGError):
        cabi.call("gg_ew_run", C.byref(p), cabi.stream_ptr())


@pytest.mark.parametrize("Bt,R,Cc", [(128, 256, 16), (128, 16, 256), (5, 33, 7), (64, 3, 1024)])
def test_strided_transpose_with_activation_gradient(Bt, R, Cc):
    """gg_transpose_b2d_ex: batches read from / written into the columns of a wider matrix, optional act'(y) on the way out"""
    from gg import cabi
    wide = torch.randn(Bt, R * Cc + 72, device="cuda")
    x = wide[:, 40:40 + R * Cc]
    ref = x.reshape(Bt, R, Cc).transpose(1, 2).contiguous()
    y_fwd = torch.randn(Bt, Cc, R, device="cuda")
    for act, alpha in ((None, 0.0), ("leaky", 0.2), ("relu", 0.0), ("tanh", 0.0)):
        out = torch.full((Bt, Cc * R + 24), 7.0, device="cuda")
        mp = y_fwd.data_ptr() if act else None
        cabi.call("gg_transpose_b2d_ex", wide.data_ptr() + 4 * 40, out.data_ptr() + 4 * 8, Bt, R, Cc, wide.shape[1], out.shape[1], mp,
                  cabi.ACT[act], alpha, cabi.stream_ptr())
        torch.cuda.synchronize()
        want = ref
        if act == "leaky":
            want = torch.where(y_fwd > 0, ref, alpha * ref)
        elif act == "relu":
            want = torch.where(y_fwd > 0, ref, torch.zeros_like(ref))
        elif act == "tanh":
            want = (1.0 - y_fwd * y_fwd) * ref
        got = out[:, 8:8 + R * Cc].reshape(Bt, Cc, R)
        if act == "tanh":
            assert (got - want).abs().max() <= 1e-6 * want.abs().max()
        else:
            assert torch.equal(got, want), act
        assert bool((out[:, :8] == 7.0).all()) and bool((out[:, 8 + R * Cc:] == 7.0).all()), "wrote outside its columns"


def test_gather_rows_matches_one_hot_matmul():
    from gg import cabi
    from gpu_util import gemm
    M, N, depth = 64, 128, 30
    idx = torch.randint(0, depth, (M,), dtype=torch.int32, device="cuda")
    idx[3] = -1
    idx[5] = depth
    table, noise = torch.randn(depth, N, device="cuda"), torch.randn(M, N, device="cuda")
    oh = torch.empty(M, depth, device="cuda")
    cabi.call("gg_one_hot", idx.data_ptr(), oh.data_ptr(), M, depth, cabi.stream_ptr())
    ref = gemm(oh, table, None, M, N, depth)
    out, out2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    cabi.call("gg_gather_rows", idx.data_ptr(), table.data_ptr(), None, out.data_ptr(), M, N, depth, cabi.stream_ptr())
    cabi.call("gg_gather_rows", idx.data_ptr(), table.data_ptr(), noise.data_ptr(), out2.data_ptr(), M, N, depth, cabi.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out, ref) and bool((out[3] == 0).all()) and bool((out[5] == 0).all())
    assert torch.equal(out2, ref + noise)


def test_dense_dgrad_with_fused_activation_gradient():
    """dx = leaky'(y) * (dy W^T) for the Linear layers of the latent discriminators (gmgan_inference_cifar10.py:262-272, backward):
    gg_conv2d_dgrad_actgrad on the 1x1 geometry gg_gemm itself uses == gg_gemm followed by gg_binary(leaky_grad)"""
    from gg import cabi
    from gpu_util import gemm, ws
    for M, N, K in ((128, 512, 512), (128, 4608, 512), (64, 128, 4096)):
        dy, W, y = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.05, torch.randn(M, N, device="cuda")
        if cabi.lib.gg_conv2d_tc_supported(1, M, 1, 1, N, K, 1, 1, 1, 1) != 1:
            continue
        two = gemm(dy, W, None, M, N, K, ta=0, tb=1)
        assert cabi.lib.gg_last_backend() == 1
        two = torch.where(y > 0, two, 0.2 * two)
        one = torch.empty(M, N, device="cuda")
        wsp = ws(cabi.lib.gg_gemm_workspace(M, N, K))
        cabi.call("gg_conv2d_dgrad_actgrad", dy.data_ptr(), W.data_ptr(), one.data_ptr(), y.data_ptr(), cabi.ACT["leaky"], 0.2,
                  M, 1, 1, N, K, 1, 1, 0, 0, 1, 1, wsp.data_ptr(), wsp.numel(), cabi.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(one, two), (M, N, K)


# ---- whole plans --------------------------------------------------------------------------------------------------------
def _builders():
    import gmgan_inference_cifar10 as Cf
    return {
        "gmgan_cifar10": lambda: Cf.build_graph(BATCH_SIZE=16),
        "gmgan_cifar10_reinforce": lambda: Cf.build_graph(BATCH_SIZE=8, MODE_K='REINFORCE'),
        "gmgan_mnist": lambda: __import__("gmgan_inference_mnist").build_graph(BATCH_SIZE=10),
        "gan_svhn_wali_gp": lambda: __import__("gan_inference_svhn").build_graph(MODE='wali-gp', BATCH_SIZE=8),
        "gan_mnist_ali": lambda: __import__("gan_inference_mnist").build_graph(MODE='ali', BATCH_SIZE=10),
        "gan_face_ali": lambda: __import__("gan_inference_face").build_graph(BATCH_SIZE=8),
        "ssgan_moving_mnist": lambda: __import__("ssgan_inference_moving_mnist").build_graph(BATCH_SIZE=4, LEN=4),
        "gmgan_svhn_local_epce": lambda: __import__("gmgan_inference_svhn").build_graph(MODE='local_epce', BATCH_SIZE=8),
    }


def _train(family, fuse, iters=3, bn_stats=False):
    import tensorflow as tf
    import tflib as lib
    from gg.executor import RT
    from gg.ops import toposort
    knobs = ("GG_FUSE_EW", "GG_FUSE_TRANSPOSE", "GG_FUSE_ACTGRAD_DENSE", "GG_RANK1_MUL", "GG_GATHER")
    for k in knobs:                                  # every launch-list fusion of round 2 on, or the one-launch-per-node plan
        os.environ[k] = "1" if fuse else "0"
    os.environ["GG_BN_CONV_STATS"] = "1" if bn_stats else "0"    # (changes the summation order of the moments: not bit-identical)
    try:
        tf.reset_default_graph()
        lib.delete_all_params()
        np.random.seed(31)
        g = _builders()[family]()
        sess = tf.Session()
        steps = ((g.gen_cost, g.gen_train_op), (g.disc_cost, g.disc_train_op))
        leaves = {}
        for cost, op in steps:
            roots = [cost] + [d for d in op.deps if d is not None]
            leaves[id(op)] = sorted((n for n in toposort(roots) if n.op in ("placeholder", "random")), key=lambda n: n.id)
        rs = np.random.RandomState(3)
        costs, n_fused = [], 0
        for it in range(iters):
            for cost, op in steps:
                feeds = {}
                for n in leaves[id(op)]:           # placeholders AND in-graph random draws: both builds see the same numbers
                    if n.dtype.name == "int32":
                        hi = n.inputs[0].size if n.op == "random" else (256 if n.size >= 1024 else 10)
                        feeds[n] = rs.randint(0, hi, size=tuple(n.shape)).astype(np.int32)
                    elif n.op == "random" and n.attrs["kind"] == "normal":
                        feeds[n] = rs.randn(*n.shape).astype(np.float32)
                    else:
                        feeds[n] = rs.uniform(0.05, 0.95, size=tuple(n.shape)).astype(np.float32)
                c, _ = sess.run([cost, op], feed_dict=feeds)
                costs.append(np.asarray(c, dtype=np.float32).copy())
        for plan in RT.plans.values():
            n_fused += len(plan.ew_clusters) + len(plan.tr_fuse) + len(plan.gather_add)
            if bn_stats:
                n_fused += 1000 * len(plan.bn_stats)
        params = {n: RT.get_param(p).copy() for n, p in sorted(lib._params.items())}
        return costs, params, n_fused
    finally:
        for k in knobs + ("GG_BN_CONV_STATS",):
            os.environ.pop(k, None)


@pytest.mark.parametrize("family", ["gmgan_cifar10", "gmgan_cifar10_reinforce", "gmgan_mnist", "gan_svhn_wali_gp", "gan_mnist_ali",
                                    "gan_face_ali", "ssgan_moving_mnist", "gmgan_svhn_local_epce"])
def test_fused_plans_are_bit_identical(family):
    c0, p0, f0 = _train(family, False)
    c1, p1, f1 = _train(family, True)
    assert f0 == 0 and f1 >= 4, (f0, f1)
    assert all(np.isfinite(c).all() for c in c0)
    for a, b in zip(c0, c1):
        assert np.array_equal(a, b), (family, a, b)
    assert set(p0) == set(p1)
    for n in p0:
        assert np.array_equal(p0[n], p1[n]), "%s: parameter %s differs between the fused and the unfused plan" % (family, n)


# ---- batch-norm statistics in the epilogue of the producing conv / deconv / dense launch ----------------------------------
BNSTAT_CASES = [
    # mode, B, H, W, Ci, Co, k, stride      (mode 1: the Deconv2D forward = gg_conv2d_dgrad geometry)
    (0, 64, 16, 16, 64, 128, 5, 2),      # Extractor.2 (gmgan_inference_cifar10.py:173-176), split-K cluster
    (0, 64, 8, 8, 128, 256, 5, 2),       # Extractor.3
    (1, 64, 8, 8, 128, 256, 5, 2),       # Generator.2 deconv 256 -> 128
    (1, 64, 16, 16, 64, 128, 5, 2),      # Generator.3 deconv 128 -> 64
    (0, 64, 1, 1, 128, 4096, 1, 1),      # Generator.1 Linear (rows 64 of a 128-row tile)
    (0, 128, 32, 32, 32, 64, 5, 2),      # face Discriminator-sized, un-split multi-wave launch (two CTAs per SM)
    (1, 128, 16, 16, 64, 128, 5, 2),     # face Generator deconv
    (0, 6, 8, 8, 32, 96, 3, 1),          # ragged batch box, n_tile 32
]


@pytest.mark.parametrize("case", BNSTAT_CASES)
def test_conv_epilogue_batchnorm_statistics(case):
    from gg import cabi
    from gpu_util import ws, geom
    mode, B, H, W, Ci, Co, k, stride = case
    Ho, Wo, pt, pl = geom(H, W, k, stride, "SAME")
    geo = (B, H, W, Ci, Co, k, stride, pt, pl, Ho, Wo)
    T = cabi.lib.gg_conv2d_stats_tiles(mode, *geo)
    assert T > 0, "shape should run on the tensor-core path"
    g = torch.Generator(device="cuda").manual_seed(3)
    w = torch.randn(k, k, Ci, Co, device="cuda", generator=g) * 0.05
    if mode == 0:
        a = torch.randn(B, H, W, Ci, device="cuda", generator=g)
        C, out_shape = Co, (B, Ho, Wo, Co)
    else:
        a = torch.randn(B, Ho, Wo, Co, device="cuda", generator=g)
        C, out_shape = Ci, (B, H, W, Ci)
    bias = torch.randn(C, device="cuda", generator=g)
    gamma, beta = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g)
    ref, out = torch.empty(out_shape, device="cuda"), torch.empty(out_shape, device="cuda")
    wsp = ws(cabi.lib.gg_conv2d_workspace(mode, B, H, W, Ci, Co, k, stride, Ho, Wo))
    name = "gg_conv2d_fwd" if mode == 0 else "gg_conv2d_dgrad"
    cabi.call(name, a.data_ptr(), w.data_ptr(), bias.data_ptr(), ref.data_ptr(), *geo, cabi.ACT[None], 0.0, wsp.data_ptr(), wsp.numel(),
              cabi.stream_ptr())
    assert cabi.lib.gg_last_backend() == 1
    stats = torch.full((T, 2, C), float("nan"), device="cuda")
    cabi.call("gg_conv2d_bnstats", mode, a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), stats.data_ptr(), *geo,
              cabi.ACT[None], 0.0, wsp.data_ptr(), wsp.numel(), cabi.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out, ref), "the statistics epilogue changed the convolution's output"
    assert bool(torch.isfinite(stats).all()), "an (m-tile, channel) slot of the statistics was never written"
    R = out.numel() // C
    x2 = ref.reshape(R, C).double()
    tot = stats.double().sum(0)
    assert (tot[0] - x2.sum(0)).abs().max() <= 1e-5 * x2.abs().sum(0).max()
    assert (tot[1] - (x2 * x2).sum(0)).abs().max() <= 1e-5 * (x2 * x2).sum(0).max()
    # through gg_bn_apply == the one-launch batch norm (and the fp64 formula)
    y1, y2 = torch.empty(R, C, device="cuda"), torch.empty(R, C, device="cuda")
    m1, r1, m2, r2 = (torch.empty(C, device="cuda") for _ in range(4))
    cabi.call("gg_bn_apply", ref.data_ptr(), stats.data_ptr(), T, float(R), gamma.data_ptr(), beta.data_ptr(), 1e-5, y1.data_ptr(),
              m1.data_ptr(), r1.data_ptr(), R, C, cabi.ACT["relu"], 0.0, cabi.stream_ptr())
    mean = x2.mean(0)
    var = x2.var(0, unbiased=False)
    want = torch.clamp((x2 - mean) * torch.rsqrt(var + 1e-5) * gamma.double() + beta.double(), min=0)
    torch.cuda.synchronize()
    assert (y1.double() - want).abs().max() <= 2e-5 * max(1.0, float(want.abs().max()))
    assert (m1.double() - mean).abs().max() <= 1e-5 * max(1.0, float(mean.abs().max()))
    if cabi.lib.gg_bn_fused_supported(R, C) == 1:
        cabi.call("gg_bn_fwd_fused", ref.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, y2.data_ptr(), m2.data_ptr(), r2.data_ptr(),
                  R, C, cabi.ACT["relu"], 0.0, cabi.stream_ptr())
        torch.cuda.synchronize()
        assert (y1 - y2).abs().max() <= 2e-5 * max(1.0, float(y2.abs().max()))
        assert (r1 - r2).abs().max() <= 1e-5 * float(r2.abs().max())


@pytest.mark.parametrize("family", ["gmgan_cifar10", "gan_mnist_ali"])      # (gan_inference_face has no batch norm)
def test_plans_with_epilogue_statistics_track_the_one_launch_batchnorm(family):
    c0, p0, f0 = _train(family, True, iters=2, bn_stats=False)
    c1, p1, f1 = _train(family, True, iters=2, bn_stats=True)
    assert f1 >= 1000, "no batch norm took its statistics from the producing launch"
    for a, b in zip(c0, c1):
        assert np.allclose(a, b, rtol=2e-3, atol=2e-4), (family, a, b)
    # parameters after two Adam steps: an update is lr * m / (sqrt(v) + eps), i.e. it has the size of lr whatever the size of
    # the gradient, so the rounding-level change of the moments shows up relative to lr (2e-4), not to the parameter
    # Adam's first steps move every weight by ~lr * sign(g): an entry whose gradient sits at the rounding-noise level may step the
    # other way (2 * lr = 4e-4 apart per step); everything else must agree closely (the criterion of tests/test_gpu_multi.py)
    worst, frac = 0.0, 0.0
    for n in p0:
        d = np.abs(p0[n] - p1[n])
        worst, frac = max(worst, float(d.max())), max(frac, float((d > 1e-4).mean()))
    print("bn-stats tracking %s: worst |dp| %.2e, worst fraction of entries apart by > 1e-4: %.4f" % (family, worst, frac))
    for n in p0:
        d = np.abs(p0[n] - p1[n])
        assert d.max() <= 1.5e-3 and (d > 1e-4).mean() < 0.03, (n, float(d.max()), float((d > 1e-4).mean()))
