"""-m gpu: per-op parity at the EXACT launches of the benchmarked step.

The toy shapes of test_gpu_kernels.py never reach the production launch configurations of the tcgen05 convolution:
split-K count, ring depth and CTA cap depend on the tile count.  Here the geometry of every conv / dense node is read
out of the compiled plans of the bench workload itself (gmgan_inference_cifar10.py, MODE=local_ep, bs=64: the sibling-
batched B=128 discriminator launches, the row-sliced dgrads, the M=64/128 dense layers) and of gan_inference_face.py
(64x64x3), and each unique launch is replayed through the C-ABI under the plan's own settings (gg_set_tc_stages(3),
gg_set_tc_max_ctas(148)) on seeded inputs against the fp64 oracle (oracle/tf_ops.py) at the north_star tolerance 1e-3.
"""
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_ops as O

TOL = 1e-3


def _plans_of(script, **kw):
    """build the script's graph, run G and D train ops once, return the compiled plans"""
    import importlib
    import tensorflow as tf
    import tflib as lib
    from gg.executor import RT
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    S = importlib.import_module(script)
    g = S.build_graph(**kw)
    sess = tf.Session()
    rs = np.random.RandomState(0)
    feeds = {}
    for name in ("real_x_int", "real_x"):
        ph = getattr(g, name, None)
        if ph is not None and ph.op == "placeholder":
            if ph.dtype.as_numpy_dtype == np.int32:
                feeds[ph] = rs.randint(0, 256, size=ph.shape).astype(np.int32)
            else:
                feeds[ph] = rs.rand(*ph.shape).astype(np.float32)
            break
    sess.run([g.gen_cost, g.gen_train_op], feed_dict=feeds)
    sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)
    return list(RT.plans.values())


def _unique_launches(plans):
    convs, gemms = {}, {}
    for plan in plans:
        for n in plan.order:
            if n.id in plan.fed:
                continue
            if n.op == "conv":
                a = n.attrs
                key = (a["mode"],) + tuple(a[k] for k in ("B", "H", "W", "Ci", "Co", "k", "stride", "pad_t", "pad_l", "Ho", "Wo")) + \
                      (a["act"] if a["mode"] != "wgrad" else None, len(n.inputs) == 3)
                convs[key] = a.get("alpha", 0.0)
            elif n.op == "matmul":
                a = n.attrs
                M, N = n.shape
                K = n.inputs[0].shape[0] if a["ta"] else n.inputs[0].shape[1]
                gemms[(M, N, K, int(a["ta"]), int(a["tb"]), a["act"], len(n.inputs) == 3)] = a["alpha"]
    return convs, gemms


def _act(t, act, alpha):
    if act == "relu":
        return torch.relu(t)
    if act == "leaky":
        return O.leaky_relu(t, alpha)
    if act == "tanh":
        return torch.tanh(t)
    if act == "sigmoid":
        return torch.sigmoid(t)
    return t


def _check_conv(U, cabi, key, alpha, report):
    mode, B, H, W, Ci, Co, k, stride, pt, pl, Ho, Wo, act, has_bias = key
    g = torch.Generator().manual_seed(zlib.crc32(repr(key).encode()))
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(k, k, Ci, Co, generator=g, dtype=torch.float64) * (1.0 / np.sqrt(k * k * Ci))).requires_grad_(True)
    bias = torch.randn(Co if mode == "fwd" else Ci, generator=g, dtype=torch.float64) if has_bias else None
    dy = torch.randn(B, Co, Ho, Wo, generator=g, dtype=torch.float64)
    ylin = O.conv2d(x, w, stride, 'SAME')
    assert tuple(ylin.shape) == (B, Co, Ho, Wo), (key, tuple(ylin.shape))
    geo = (B, H, W, Ci, Co, k, stride, pt, pl, Ho, Wo)
    xd, wd, dyd = U.dev(U.nhwc(x.detach())), U.dev(w.detach()), U.dev(U.nhwc(dy))
    bd = U.dev(bias) if has_bias else None
    if mode == "fwd":
        ref = _act(ylin + (bias.view(1, -1, 1, 1) if has_bias else 0.0), act, alpha).detach()
        out = torch.empty(B, Ho, Wo, Co, device="cuda")
        wsp = U.ws(cabi.lib.gg_conv2d_workspace(0, *geo[:7], Ho, Wo))
        cabi.call("gg_conv2d_fwd", xd.data_ptr(), wd.data_ptr(), cabi.ptr(bd), out.data_ptr(), *geo, cabi.ACT[act], alpha,
                  wsp.data_ptr(), wsp.numel(), cabi.stream_ptr())
        got = U.nchw(out)
    elif mode == "dgrad":
        dx, = torch.autograd.grad(ylin, x, dy)
        ref = _act(dx + (bias.view(1, -1, 1, 1) if has_bias else 0.0), act, alpha)
        out = torch.empty(B, H, W, Ci, device="cuda")
        wsp = U.ws(cabi.lib.gg_conv2d_workspace(1, *geo[:7], Ho, Wo))
        cabi.call("gg_conv2d_dgrad", dyd.data_ptr(), wd.data_ptr(), cabi.ptr(bd), out.data_ptr(), *geo, cabi.ACT[act], alpha,
                  wsp.data_ptr(), wsp.numel(), cabi.stream_ptr())
        got = U.nchw(out)
    else:
        ref, = torch.autograd.grad(ylin, w, dy)
        out = torch.empty(k, k, Ci, Co, device="cuda")
        wsp = U.ws(cabi.lib.gg_conv2d_wgrad_workspace(*geo[:7], Ho, Wo))
        cabi.call("gg_conv2d_wgrad", xd.data_ptr(), dyd.data_ptr(), out.data_ptr(), *geo, wsp.data_ptr(), wsp.numel(),
                  cabi.stream_ptr())
        got = out
    torch.cuda.synchronize()
    backend = cabi.lib.gg_last_backend()
    info = cabi.last_tc_info() if backend else {}
    U.assert_close(got, ref, TOL, "conv %s" % (key,))
    report.append((key, backend, info.get("tiles"), info.get("splits"), info.get("n_tile"), info.get("stages"), U.rel_err(got, ref)))
    return backend, info


def _check_gemm(U, cabi, key, alpha, report):
    M, N, K, ta, tb, act, has_bias = key
    g = torch.Generator().manual_seed(zlib.crc32(repr(key).encode()))
    A = torch.randn((K, M) if ta else (M, K), generator=g, dtype=torch.float64)
    Bm = torch.randn((N, K) if tb else (K, N), generator=g, dtype=torch.float64) / np.sqrt(K)
    bias = torch.randn(N, generator=g, dtype=torch.float64) if has_bias else None
    ref = (A.t() if ta else A) @ (Bm.t() if tb else Bm)
    ref = _act(ref + (bias if has_bias else 0.0), act, alpha)
    out = U.gemm(U.dev(A), U.dev(Bm), U.dev(bias) if has_bias else None, M, N, K, ta, tb, act=act, alpha=alpha)
    torch.cuda.synchronize()
    backend = cabi.lib.gg_last_backend()
    info = cabi.last_tc_info() if backend else {}
    U.assert_close(out, ref, TOL, "gemm %s" % (key,))
    report.append((key, backend, info.get("tiles"), info.get("splits"), info.get("n_tile"), info.get("stages"), U.rel_err(out, ref)))
    return backend, info


def _run_workload(script, kw, must_have):
    import gpu_util as U
    from gg import cabi
    plans = _plans_of(script, **kw)
    convs, gemms = _unique_launches(plans)
    for need in must_have:
        assert any(all(k[i] == v for i, v in need.items()) for k in convs), "plan has no conv launch matching %s" % (need,)
    # the multi-stream plan's own tensor-core settings (gg/executor.py Plan.__init__)
    cabi.call("gg_set_tc_max_ctas", 148)
    cabi.call("gg_set_tc_stages", 3)
    report = []
    for key, alpha in sorted(convs.items(), key=str):
        backend, info = _check_conv(U, cabi, key, alpha, report)
        mode, B, H, W, Ci, Co = key[:6]
        if Ci % 32 == 0 and Co % 32 == 0:
            assert backend == 1, "tensor-core path not taken for production launch %s" % (key,)
            # un-split launches keep their ring under half an SM (two CTAs per SM); split-K clusters own their SM and run a deeper ring
            stage_bytes = 128 * 32 * 4 + info["n_tile"] * 128
            assert (info["stages"] * stage_bytes <= 113 * 1024 or info["splits"] > 1) and 1 <= info["splits"] <= 8, info
            assert info["tiles"] * info["splits"] <= 2 * 148 or info["splits"] == 1, info
    for key, alpha in sorted(gemms.items(), key=str):
        backend, info = _check_gemm(U, cabi, key, alpha, report)
        M, N, K = key[:3]
        if N % 32 == 0 and K % 32 == 0 and not key[3] and not key[4]:
            assert backend == 1, "tensor-core path not taken for production dense launch %s" % (key,)
    print("\n%s: %d conv + %d dense launches at production geometry (key, tcgen05?, tiles, splits, n_tile, stages, err/scale)" %
          (script, len(convs), len(gemms)))
    for r in report:
        print("  ", r)
    assert max(r[-1] for r in report) < TOL


def test_gmgan_cifar10_bench_launches_match_oracle():
    """BASELINE.json configs[1] at the bench batch: D towers batched to B=128 (3->64, 64->128, 128->256), E at B=64, the
    three Generator deconvolutions (256->128->64->3) as dgrad-mode launches, every dense layer"""
    _run_workload("gmgan_inference_cifar10", dict(BATCH_SIZE=64),
                  must_have=[{0: "fwd", 1: 128, 4: 64, 5: 128}, {0: "fwd", 1: 128, 4: 128, 5: 256}, {0: "fwd", 1: 64, 4: 64, 5: 128},
                             {0: "dgrad", 1: 64, 4: 128, 5: 256}, {0: "wgrad", 1: 128, 4: 64, 5: 128}])


def test_gan_face_64x64_launches_match_oracle():
    """configs[3] geometry (gan_inference_face.py: 64x64x3, DIM 32) at a per-GPU shard of 16 images (bs=128 over 8 GPUs) and
    at the full single-GPU batch for the dominant layers"""
    _run_workload("gan_inference_face", dict(BATCH_SIZE=16), must_have=[{0: "fwd", 4: 32, 5: 64}, {0: "fwd", 4: 64, 5: 128}])


def test_dominant_kernel_matches_fp32_direct_backend():
    """the roofline kernel of bench.py (Discriminator.2 forward on the batched towers) against the fp32 direct backend on the
    same device buffers — the check bench.py itself repeats after timing"""
    import gpu_util as U
    from gg import cabi
    cabi.call("gg_set_tc_max_ctas", 148)
    cabi.call("gg_set_tc_stages", 3)
    g = torch.Generator().manual_seed(5)
    x = U.dev(torch.randn(128, 16, 16, 64, generator=g))
    w = U.dev(torch.randn(5, 5, 64, 128, generator=g) * 0.025)
    b = U.dev(torch.randn(128, generator=g))
    y_tc = U.conv_fwd(x, w, b, 2, 'SAME', act="leaky")
    assert cabi.lib.gg_last_backend() == 1
    cabi.call("gg_set_conv_backend", 1)
    try:
        y_fp32 = U.conv_fwd(x, w, b, 2, 'SAME', act="leaky")
        assert cabi.lib.gg_last_backend() == 0
    finally:
        cabi.call("gg_set_conv_backend", 0)
    U.assert_close(y_tc, y_fp32, TOL, "D.2 batched forward, tcgen05 vs fp32 direct")
