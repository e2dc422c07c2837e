"""-m gpu: every C-ABI kernel against the CPU oracle (oracle/tf_ops.py) on the same seeded inputs.

Tolerances: conv / linear 1e-3 relative (BASELINE.json north_star); pure fp32 elementwise work 1e-5;
index / shape / integer work bit-exact.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_ops as O


def _mods():
    import gpu_util as U
    from gg import cabi
    return U, cabi


CONV_CASES = [
    # B, H, W, Ci, Co, k, stride, padding      (reference call sites)
    (4, 32, 32, 3, 64, 5, 2, 'SAME'),     # Extractor.1 / Discriminator.1  gmgan_inference_cifar10.py:200,276
    (4, 16, 16, 64, 128, 5, 2, 'SAME'),   # Extractor.2 / Discriminator.2  :203,280
    (8, 8, 8, 128, 256, 5, 2, 'SAME'),    # Extractor.3 / Discriminator.3  :208,284
    (3, 28, 28, 1, 64, 5, 2, 'SAME'),     # MNIST first layer, 28 -> 14
    (3, 14, 14, 64, 128, 5, 2, 'SAME'),   # 14 -> 7
    (3, 7, 7, 128, 256, 5, 2, 'SAME'),    # 7 -> 4: pad (2,2)
    (2, 16, 16, 32, 32, 3, 1, 'SAME'),    # 3x3 s1 (synthetic)
    (2, 16, 16, 32, 64, 3, 2, 'SAME'),    # 3x3 s2 (synthetic)
    (2, 7, 7, 16, 8, 4, 1, 'VALID'),      # 4x4 s1 VALID  ssgan_inference_moving_mnist.py:483
    (1, 5, 6, 2, 3, 5, 2, 'SAME'),        # ragged: odd sizes, channel counts not multiples of 4
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("backend", [1, 0])
def test_conv_fwd_dgrad_wgrad(case, backend):
    U, cabi = _mods()
    B, H, W, Ci, Co, k, stride, padding = case
    g = torch.Generator().manual_seed(1234 + H + Ci)
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(k, k, Ci, Co, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
    bias = torch.randn(Co, generator=g, dtype=torch.float64)
    y = O.leaky_relu(O.conv2d(x, w, stride, padding, bias), 0.2)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    ylin = O.conv2d(x, w, stride, padding)
    dx, dw = torch.autograd.grad(ylin, (x, w), gy)

    cabi.call("gg_set_conv_backend", backend)
    try:
        xd, wd, bd = U.dev(U.nhwc(x.detach())), U.dev(w.detach()), U.dev(bias)
        used = []
        yd = U.conv_fwd(xd, wd, bd, stride, padding, act="leaky", alpha=0.2)
        used.append(cabi.lib.gg_last_backend())
        U.assert_close(U.nchw(yd), y, 1e-3, "conv fwd %s" % (case,))
        gyd = U.dev(U.nhwc(gy))
        dxd = U.conv_dgrad(gyd, wd, None, H, W, stride, padding)
        used.append(cabi.lib.gg_last_backend())
        U.assert_close(U.nchw(dxd), dx, 1e-3, "conv dgrad %s" % (case,))
        dwd = U.conv_wgrad(xd, gyd, k, stride, padding)
        used.append(cabi.lib.gg_last_backend())
        U.assert_close(dwd, dw, 1e-3, "conv wgrad %s" % (case,))
        print("case", case, "backend mode", backend, "used (fwd,dgrad,wgrad) [1=tcgen05]:", used)
        if backend == 1:
            assert used == [0, 0, 0]
        elif Ci % 32 == 0 and Co % 32 == 0 and padding == 'SAME' and H % 2 == 0 and B % 2 == 0:
            assert used == [1, 1, 1], "tensor-core path not taken for %s: %s" % (case, used)
    finally:
        cabi.call("gg_set_conv_backend", 0)


DECONV_CASES = [
    # B, Hin, Cin, Cout, k      Deconv2D 5x5 s2 SAME (gmgan_inference_cifar10.py:182,187,192)
    (8, 4, 256, 128, 5),
    (4, 8, 128, 64, 5),
    (4, 16, 64, 3, 5),
    (2, 7, 64, 1, 5),    # MNIST last layer 14 -> 28 is (2,14,64,1); 7 -> 14 here
]


@pytest.mark.parametrize("case", DECONV_CASES)
def test_deconv_forward_matches_conv2d_transpose(case):
    U, cabi = _mods()
    B, Hin, Cin, Cout, k = case
    g = torch.Generator().manual_seed(99 + Hin)
    x = torch.randn(B, Cin, Hin, Hin, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, Cout, Cin, generator=g, dtype=torch.float64) * 0.02  # deconv2d.py:60-69; pre-tanh values O(1)
    bias = torch.randn(Cout, generator=g, dtype=torch.float64)
    y = torch.tanh(O.conv2d_transpose(x, w, 2, 'SAME', bias))
    # Deconv2D forward == conv dgrad with Ci := Cout, Co := Cin on the 2H x 2W grid
    yd = U.conv_dgrad(U.dev(U.nhwc(x)), U.dev(w), U.dev(bias), 2 * Hin, 2 * Hin, 2, 'SAME', act="tanh")
    U.assert_close(U.nchw(yd), y, 1e-3, "deconv fwd %s" % (case,))


GEMM_CASES = [
    (64, 4096, 128, 0, 0),   # Generator.Input  gmgan_inference_cifar10.py:176
    (64, 128, 4096, 0, 0),   # Extractor.Output :224
    (64, 512, 4608, 0, 0),   # Discriminator.zx1 :295
    (64, 1, 512, 0, 0),      # Discriminator.Output :299
    (64, 512, 158, 0, 0),    # Discriminator.HyperInput :257
    (4608, 512, 64, 1, 0),   # dW = X^T dY
    (64, 4608, 512, 0, 1),   # dX = dY W^T
    (50, 33, 77, 0, 0), (33, 50, 77, 1, 1),   # ragged
    # the thin products served by gemm_rowdot_kernel / gemm_smallk_kernel (critic output layer and its gradients, prior lookups)
    (128, 1, 512, 0, 0), (128, 512, 1, 0, 1), (512, 1, 128, 1, 0), (64, 128, 30, 0, 0), (30, 128, 64, 1, 0), (128, 3, 200, 0, 1),
    (7, 5, 3, 1, 1), (19, 2, 1000, 0, 0),
]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm(case):
    U, cabi = _mods()
    M, N, K, ta, tb = case
    g = torch.Generator().manual_seed(7 + M + N)
    A = torch.randn((K, M) if ta else (M, K), generator=g, dtype=torch.float64)
    Bm = torch.randn((N, K) if tb else (K, N), generator=g, dtype=torch.float64)
    bias = torch.randn(N, generator=g, dtype=torch.float64)
    ref = O.leaky_relu((A.t() if ta else A) @ (Bm.t() if tb else Bm) + bias, 0.2)
    out = U.gemm(U.dev(A), U.dev(Bm), U.dev(bias), M, N, K, ta, tb, act="leaky", alpha=0.2)
    U.assert_close(out, ref, 1e-3, "gemm %s" % (case,))


@pytest.mark.parametrize("R,Cc,act", [(64, 4096, "relu"), (64 * 8 * 8, 128, "relu"), (64 * 16 * 16, 64, None),
                                      (64 * 4 * 4, 256, "leaky"), (50, 7, None), (1000, 33, "relu"),
                                      (50 * 14 * 14, 64, "leaky"), (50 * 7 * 7, 128, None), (37, 12, "relu"), (3, 4, None)])
def test_batchnorm_fwd_bwd(R, Cc, act):
    U, cabi = _mods()
    g = torch.Generator().manual_seed(5 + R)
    x = (torch.randn(R, Cc, generator=g, dtype=torch.float64) * 1.7 + 0.3).requires_grad_(True)
    gamma = (torch.rand(Cc, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = torch.randn(Cc, generator=g, dtype=torch.float64).requires_grad_(True)
    y = O.batchnorm(x, gamma, beta, [0])
    if act == "relu":
        y = torch.relu(y)
    elif act == "leaky":
        y = O.leaky_relu(y, 0.2)
    gy = torch.randn(R, Cc, generator=g, dtype=torch.float64)
    dx, dg, db = torch.autograd.grad(y, (x, gamma, beta), gy)

    S = cabi.lib.gg_bn_slices(R, Cc)
    xd, gd, bd, gyd = U.dev(x.detach()), U.dev(gamma.detach()), U.dev(beta.detach()), U.dev(gy)
    part = torch.empty(S, 2, Cc, device="cuda")
    yd = torch.empty(R, Cc, device="cuda")
    mean = torch.empty(Cc, device="cuda")
    rstd = torch.empty(Cc, device="cuda")
    st = cabi.stream_ptr()
    cabi.call("gg_bn_stats", cabi.ptr(xd), cabi.ptr(part), R, Cc, st)
    cabi.call("gg_bn_apply", cabi.ptr(xd), cabi.ptr(part), S, float(R), cabi.ptr(gd), cabi.ptr(bd), 1e-5, cabi.ptr(yd),
              cabi.ptr(mean), cabi.ptr(rstd), R, Cc, cabi.ACT[act], 0.2, st)
    U.assert_close(yd, y, 1e-5, "bn fwd")
    part2 = torch.empty(S, 2, Cc, device="cuda")
    dxd = torch.empty(R, Cc, device="cuda")
    dgd = torch.empty(Cc, device="cuda")
    dbd = torch.empty(Cc, device="cuda")
    cabi.call("gg_bn_bwd_reduce", cabi.ptr(gyd), cabi.ptr(xd), cabi.ptr(yd), cabi.ptr(mean), cabi.ptr(rstd), cabi.ptr(gd),
              cabi.ptr(bd), cabi.ptr(part2), R, Cc, cabi.ACT[act], 0.2, st)
    cabi.call("gg_bn_bwd_apply", cabi.ptr(gyd), cabi.ptr(xd), cabi.ptr(yd), cabi.ptr(mean), cabi.ptr(rstd), cabi.ptr(gd),
              cabi.ptr(bd), cabi.ptr(part2), S, float(R), cabi.ptr(dxd), cabi.ptr(dgd), cabi.ptr(dbd), R, Cc, cabi.ACT[act],
              0.2, st)
    U.assert_close(dxd, dx, 1e-4, "bn dx")
    U.assert_close(dgd, dg, 1e-4, "bn dgamma")
    U.assert_close(dbd, db, 1e-4, "bn dbeta")
    # the single-kernel (thread-block-cluster) forms used by the single-GPU plan: same contract, no partial buffers
    assert cabi.lib.gg_bn_fused_supported(R, Cc) == (1 if Cc % 4 == 0 else 0)
    if Cc % 4 == 0:
        y2, mean2, rstd2 = torch.empty(R, Cc, device="cuda"), torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
        cabi.call("gg_bn_fwd_fused", cabi.ptr(xd), cabi.ptr(gd), cabi.ptr(bd), 1e-5, cabi.ptr(y2), cabi.ptr(mean2),
                  cabi.ptr(rstd2), R, Cc, cabi.ACT[act], 0.2, st)
        U.assert_close(y2, y, 1e-5, "bn fused fwd")
        U.assert_close(mean2, mean, 1e-5, "bn fused mean")
        U.assert_close(rstd2, rstd, 1e-5, "bn fused rstd")
        dx2, dg2, db2 = torch.empty(R, Cc, device="cuda"), torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
        cabi.call("gg_bn_bwd_fused", cabi.ptr(gyd), cabi.ptr(xd), cabi.ptr(y2), cabi.ptr(mean2), cabi.ptr(rstd2), cabi.ptr(gd),
                  cabi.ptr(dx2), cabi.ptr(dg2), cabi.ptr(db2), R, Cc, cabi.ACT[act], 0.2, st)
        U.assert_close(dx2, dx, 1e-4, "bn fused dx")
        U.assert_close(dg2, dg, 1e-4, "bn fused dgamma")
        U.assert_close(db2, db, 1e-4, "bn fused dbeta")


def test_unary_binary_reduce_softmax():
    U, cabi = _mods()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(64, 3072, generator=g)
    xd = U.dev(x)
    st = cabi.stream_ptr()
    refs = {"relu": torch.relu(x), "leaky": O.leaky_relu(x, 0.2), "tanh": torch.tanh(x), "sigmoid": torch.sigmoid(x),
            "exp": torch.exp(x), "square": x * x, "neg": -x, "abs": x.abs(), "affine": 2.0 * x - 1.0,
            "bce": O.sigmoid_cross_entropy_with_logits(x, torch.ones_like(x))}
    for name, ref in refs.items():
        out = torch.empty_like(xd)
        a, b = {"leaky": (0.2, 0), "affine": (2.0, -1.0), "bce": (1.0, 0)}.get(name, (0.0, 0.0))
        cabi.call("gg_unary", cabi.UNARY[name], cabi.ptr(xd), cabi.ptr(out), x.numel(), a, b, st)
        U.assert_close(out, ref, 1e-5, "unary " + name)
    # odd length / unaligned tail
    out = torch.empty(1001, device="cuda")
    cabi.call("gg_unary", cabi.UNARY["square"], cabi.ptr(xd), cabi.ptr(out), 1001, 0.0, 0.0, st)
    U.assert_close(out, x.flatten()[:1001] ** 2, 1e-6, "unary tail")

    # broadcasting: [64,1,128] - [1,30,128]  (HyperExtractor, gmgan_inference_cifar10.py:158)
    z = torch.randn(64, 128, generator=g)
    mu = torch.randn(30, 128, generator=g)
    out = torch.empty(64, 30, 128, device="cuda")
    zd, mud = U.dev(z), U.dev(mu)
    cabi.call("gg_binary", cabi.BINARY["sub"], cabi.ptr(zd), cabi.ptr(mud), cabi.ptr(out),
              cabi.int4([1, 64, 30, 128]), cabi.int4([0, 128, 0, 1]), cabi.int4([0, 0, 128, 1]), 0.0, st)
    U.assert_close(out, z[:, None, :] - mu[None, :, :], 1e-6, "binary bcast")
    # flat
    y = torch.randn(64, 3072, generator=g)
    out = torch.empty_like(xd)
    yd = U.dev(y)
    cabi.call("gg_binary", cabi.BINARY["mul"], cabi.ptr(xd), cabi.ptr(yd), cabi.ptr(out),
              cabi.int4([1, 1, 64, 3072]), cabi.int4([0, 0, 3072, 1]), cabi.int4([0, 0, 3072, 1]), 0.0, st)
    U.assert_close(out, x * y, 1e-6, "binary flat")
    # reduce: sum over last axis, mean over rows, max
    t = torch.randn(64, 30, 128, generator=g)
    td = U.dev(t)
    out = torch.empty(64, 30, device="cuda")
    cabi.call("gg_reduce", cabi.REDUCE["sum"], cabi.ptr(td), cabi.ptr(out), 64 * 30, 128, 1, st)
    U.assert_close(out, t.sum(-1), 1e-5, "reduce sum last")
    out = torch.empty(30, 128, device="cuda")
    cabi.call("gg_reduce", cabi.REDUCE["mean"], cabi.ptr(td), cabi.ptr(out), 1, 64, 30 * 128, st)
    U.assert_close(out, t.mean(0), 1e-5, "reduce mean rows")
    out = torch.empty(64, 128, device="cuda")
    cabi.call("gg_reduce", cabi.REDUCE["max"], cabi.ptr(td), cabi.ptr(out), 64, 30, 128, st)
    U.assert_close(out, t.max(1).values, 0.0 + 1e-7, "reduce max mid")
    # softmax fwd/bwd
    lg = torch.randn(64, 30, generator=g, dtype=torch.float64, requires_grad=True)
    sm = torch.softmax(lg / 0.1, dim=-1)
    gy = torch.randn(64, 30, generator=g, dtype=torch.float64)
    (dl,) = torch.autograd.grad(sm, lg, gy)
    smd = torch.empty(64, 30, device="cuda")
    lgd = U.dev(lg.detach() / 0.1)
    cabi.call("gg_softmax_fwd", cabi.ptr(lgd), cabi.ptr(smd), 64, 30, st)
    U.assert_close(smd, sm, 1e-5, "softmax fwd")
    dld = torch.empty(64, 30, device="cuda")
    gyd = U.dev(gy)
    cabi.call("gg_softmax_bwd", cabi.ptr(smd), cabi.ptr(gyd), cabi.ptr(dld), 64, 30, st)
    U.assert_close(dld / 0.1, dl, 1e-4, "softmax bwd")


def test_layout_index_and_cast_ops_bit_exact():
    U, cabi = _mods()
    g = torch.Generator().manual_seed(11)
    st = cabi.stream_ptr()
    x = torch.randn(6, 5, 7, 3, generator=g)
    xd = U.dev(x)
    out = torch.empty(6, 3, 5, 7, device="cuda")
    cabi.call("gg_transpose_b2d", cabi.ptr(xd), cabi.ptr(out), 6, 35, 3, st)     # NHWC -> NCHW
    assert torch.equal(out.cpu(), x.permute(0, 3, 1, 2).contiguous())
    out2 = torch.empty(7, 6, 3, 5, device="cuda")
    cabi.call("gg_transpose4", cabi.ptr(xd), cabi.ptr(out2), cabi.int4([6, 5, 7, 3]), cabi.int4([2, 0, 3, 1]), st)
    assert torch.equal(out2.cpu(), x.permute(2, 0, 3, 1).contiguous())
    # concat along axis 1 through copy2d
    a, b = torch.randn(64, 4096, generator=g), torch.randn(64, 512, generator=g)
    cat = torch.empty(64, 4608, device="cuda")
    ad, bd = U.dev(a), U.dev(b)
    cabi.call("gg_copy2d", cabi.ptr(ad), 4096, cabi.ptr(cat), 4608, 64, 4096, 0, st)
    cabi.call("gg_copy2d", cabi.ptr(bd), 512, cat.data_ptr() + 4096 * 4, 4608, 64, 512, 0, st)
    assert torch.equal(cat.cpu(), torch.cat([a, b], 1))
    # one_hot / argmax
    idx = torch.randint(0, 30, (64,), generator=g, dtype=torch.int32)
    oh = torch.empty(64, 30, device="cuda")
    idxd = idx.cuda()
    cabi.call("gg_one_hot", cabi.ptr(idxd), cabi.ptr(oh), 64, 30, st)
    assert torch.equal(oh.cpu(), torch.nn.functional.one_hot(idx.long(), 30).float())
    lg = torch.randn(64, 30, generator=g)
    lg[3, 5] = lg[3, 9] = 100.0   # tie -> first index, like tf.argmax
    am = torch.empty(64, dtype=torch.int32, device="cuda")
    lgd = U.dev(lg)
    cabi.call("gg_argmax", cabi.ptr(lgd), cabi.ptr(am), 64, 30, st)
    assert torch.equal(am.cpu().long(), lg.argmax(1)) and int(am[3]) == 5
    # int32 image decode: the host emits cast (a=1,b=0) and the affine chain separately for exact parity
    xi = torch.randint(0, 256, (64, 3072), generator=g, dtype=torch.int32)
    xf = torch.empty(64, 3072, device="cuda")
    xid = xi.cuda()
    cabi.call("gg_cast_i32_f32", cabi.ptr(xid), cabi.ptr(xf), xi.numel(), 1.0, 0.0, st)
    assert torch.equal(xf.cpu(), xi.float())
    back = torch.empty(64, 3072, dtype=torch.int32, device="cuda")
    cabi.call("gg_cast_f32_i32", cabi.ptr(xf), cabi.ptr(back), xi.numel(), st)
    assert torch.equal(back.cpu(), xi)
    # empty inputs are a no-op
    cabi.call("gg_unary", 0, None, None, 0, 0.0, 0.0, st)
    cabi.call("gg_fill", cabi.ptr(xf), 0, 1.0, st)
    # add_n
    ts = [torch.randn(1000, generator=g) for _ in range(3)]
    tds = [U.dev(t) for t in ts]
    arr = (C.c_void_p * 3)(*[t.data_ptr() for t in tds])
    out = torch.empty(1000, device="cuda")
    cabi.call("gg_add_n", arr, 3, cabi.ptr(out), 1000, st)
    U.assert_close(out, ts[0] + ts[1] + ts[2], 1e-6, "add_n")


def test_losses():
    U, cabi = _mods()
    g = torch.Generator().manual_seed(21)
    st = cabi.stream_ptr()
    x = (torch.randn(64, generator=g, dtype=torch.float64) * 3).requires_grad_(True)
    for label in (0.0, 1.0):
        ref = O.bce_mean(x, label)
        (gref,) = torch.autograd.grad(ref, x)
        out = torch.zeros(1, device="cuda")
        xd = U.dev(x.detach())
        cabi.call("gg_bce_mean", cabi.ptr(xd), 64, label, 0.5, cabi.ptr(out), 0, st)
        U.assert_close(out, 0.5 * ref.detach().reshape(1), 1e-5, "bce mean")
        dx = torch.empty(64, device="cuda")
        cabi.call("gg_bce_mean_grad", cabi.ptr(xd), 64, label, 0.5, None, cabi.ptr(dx), 0, st)
        U.assert_close(dx, 0.5 * gref, 1e-5, "bce grad")
    a = torch.randn(64, 3072, generator=g, dtype=torch.float64)
    b = torch.randn(64, 3072, generator=g, dtype=torch.float64)
    for p, name in ((2, 'l2'), (1, 'l1')):
        out = torch.zeros(1, device="cuda")
        ad, bd = U.dev(a), U.dev(b)
        cabi.call("gg_dist_mean", cabi.ptr(ad), cabi.ptr(bd), a.numel(), p, 1.0, cabi.ptr(out), 0, st)
        U.assert_close(out, O.distance(a, b, name).reshape(1), 1e-5, "distance " + name)
    gr = torch.randn(64, 3072, generator=g, dtype=torch.float64) * 0.02
    slopes = torch.empty(64, device="cuda")
    out = torch.zeros(1, device="cuda")
    grd = U.dev(gr)
    cabi.call("gg_gp_slope_penalty", cabi.ptr(grd), 64, 3072, 10.0, cabi.ptr(slopes), cabi.ptr(out), st)
    U.assert_close(out, O.gradient_penalty(gr, 10.0).reshape(1), 1e-5, "gp")


def _adam_tables(cabi, params, grads, ms, vs):
    import struct
    entries = b"".join(struct.pack("<QQQQq", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
                       for p, g, m, v in zip(params, grads, ms, vs))
    chunks = []
    for ti, p in enumerate(params):
        for off in range(0, p.numel(), cabi.GG_ADAM_CHUNK):
            chunks.append(struct.pack("<iiq", ti, 0, off))
    tab = torch.frombuffer(bytearray(entries), dtype=torch.uint8).cuda()
    chk = torch.frombuffer(bytearray(b"".join(chunks)), dtype=torch.uint8).cuda()
    return tab, chk, len(chunks)


def test_adam_matches_tf_form():
    U, cabi = _mods()
    g = torch.Generator().manual_seed(31)
    shapes = [(5, 5, 64, 128), (128,), (4608, 512), (1,), (30, 128)]
    p_ref = [torch.randn(s, generator=g, dtype=torch.float64) for s in shapes]
    opt = O.TFAdam(p_ref, lr=2e-4, beta1=0.5, beta2=0.999)
    pd = [U.dev(p) for p in p_ref]
    md = [torch.zeros_like(p) for p in pd]
    vd = [torch.zeros_like(p) for p in pd]
    gd = [torch.empty_like(p) for p in pd]
    tab, chk, n_chunks = _adam_tables(cabi, pd, gd, md, vd)
    state = torch.zeros(3, dtype=torch.float64, device="cuda")
    for step in range(5):
        grads = [torch.randn(s, generator=g, dtype=torch.float64) * (10.0 ** (step - 3)) for s in shapes]
        for t, gr in zip(gd, grads):
            t.copy_(gr.float())
        opt.step(grads)
        cabi.call("gg_adam_multi", cabi.ptr(tab), cabi.ptr(chk), n_chunks, cabi.ptr(state), 2e-4, 0.5, 0.999, 1e-8, 1.0,
                  cabi.stream_ptr())
    for a, b in zip(pd, p_ref):
        U.assert_close(a, b, 1e-5, "adam params after 5 steps")
    assert int(state.view(torch.int64)[2]) == 5


def test_rng_statistics_and_replay():
    U, cabi = _mods()
    st = cabi.stream_ptr()
    tick = torch.zeros(1, dtype=torch.int64, device="cuda")
    n = 1 << 20
    a = torch.empty(n, device="cuda")
    cabi.call("gg_rng_normal", cabi.ptr(a), n, 0.0, 1.0, 1234, 7, cabi.ptr(tick), st)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1.0) < 5e-3
    b = torch.empty(n, device="cuda")
    cabi.call("gg_rng_normal", cabi.ptr(b), n, 0.0, 1.0, 1234, 7, cabi.ptr(tick), st)
    assert torch.equal(a, b)                       # same (seed, stream, tick) -> same numbers
    cabi.call("gg_rng_tick", cabi.ptr(tick), st)
    cabi.call("gg_rng_normal", cabi.ptr(b), n, 0.0, 1.0, 1234, 7, cabi.ptr(tick), st)
    assert not torch.equal(a, b)                   # a tick advances the stream
    u = torch.empty(n, device="cuda")
    cabi.call("gg_rng_uniform", cabi.ptr(u), n, 0.0, 1.0, 1234, 8, cabi.ptr(tick), st)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0 and abs(float(u.mean()) - 0.5) < 2e-3
    probs = torch.full((30,), 1.0 / 30, device="cuda")
    idx = torch.empty(60000, dtype=torch.int32, device="cuda")
    cabi.call("gg_rng_categorical", cabi.ptr(idx), 60000, cabi.ptr(probs), 30, 1234, 9, cabi.ptr(tick), st)
    cnt = torch.bincount(idx.long().cpu(), minlength=30)
    assert int(idx.min()) >= 0 and int(idx.max()) < 30 and int(cnt.min()) > 1700 and int(cnt.max()) < 2300


@pytest.mark.parametrize("case", [
    # B, H, W, Ci, Co, k, stride: the dgrad launches of the cifar plan that are followed by an activation gradient
    (128, 16, 16, 64, 128, 5, 2), (128, 8, 8, 128, 256, 5, 2), (64, 16, 16, 64, 128, 5, 2), (64, 8, 8, 128, 256, 5, 2),
    (4, 16, 16, 32, 64, 3, 1), (256, 32, 32, 32, 64, 5, 2),
])
@pytest.mark.parametrize("act", ["leaky", "relu"])
def test_dgrad_with_fused_activation_gradient(case, act):
    """gg_conv2d_dgrad_actgrad (dx = act'(y) * dgrad(dy, w), the tensor-core write-out applies the mask) is EXACTLY the separate
    pair gg_conv2d_dgrad -> leaky_grad / relu_grad it replaces in the plan (same accumulation, one extra multiply), and is
    within 1e-3 of the fp64 oracle."""
    U, cabi = _mods()
    B, H, W, Ci, Co, k, stride = case
    g = torch.Generator().manual_seed(31 + B + Ci)
    w = torch.randn(k, k, Ci, Co, generator=g, dtype=torch.float64) * 0.05
    Ho, Wo, pt, pl = U.geom(H, W, k, stride, 'SAME')
    dy = torch.randn(B, Co, Ho, Wo, generator=g, dtype=torch.float64)
    y = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64)
    x = torch.zeros(B, Ci, H, W, dtype=torch.float64, requires_grad=True)
    dx_ref, = torch.autograd.grad(O.conv2d(x, w, stride, 'SAME'), x, dy)
    slope = 0.2 if act == "leaky" else 0.0
    ref = torch.where(y > 0, dx_ref, slope * dx_ref)
    assert cabi.lib.gg_conv2d_tc_supported(1, B, H, W, Ci, Co, k, stride, Ho, Wo) == 1
    dyd, wd, yd = U.dev(U.nhwc(dy)), U.dev(w), U.dev(U.nhwc(y))
    plain = U.conv_dgrad(dyd, wd, None, H, W, stride, 'SAME')
    two_step = torch.where(yd > 0, plain, slope * plain)
    fused = torch.empty_like(plain)
    wsp = U.ws(cabi.lib.gg_conv2d_workspace(1, B, H, W, Ci, Co, k, stride, Ho, Wo))
    cabi.call("gg_conv2d_dgrad_actgrad", cabi.ptr(dyd), cabi.ptr(wd), cabi.ptr(fused), cabi.ptr(yd), cabi.ACT[act], 0.2, B, H, W,
              Ci, Co, k, stride, pt, pl, Ho, Wo, cabi.ptr(wsp), wsp.numel(), cabi.stream_ptr())
    assert cabi.lib.gg_last_backend() == 1
    assert torch.equal(fused, two_step), "fused write-out differs from dgrad + mask: max |d| = %g" % float((fused - two_step).abs().max())
    U.assert_close(U.nchw(fused), ref, 1e-3, "dgrad+actgrad %s %s" % (case, act))
