"""-m gpu: the CUDA path (through the C-ABI) against the committed golden vectors in tests/golden/ (no oracle code runs)."""
import os
import struct

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: d[k] for k in d.files}


@pytest.mark.parametrize("name", ["conv_e1", "conv_e2", "conv_e3", "conv_mnist3"])
def test_conv_golden(name):
    import gpu_util as U
    d = _load(name)
    x, w, b, gy = (torch.from_numpy(d[k]) for k in ("x", "w", "b", "gy"))
    xd, wd, bd, gyd = U.dev(U.nhwc(x)), U.dev(w), U.dev(b), U.dev(U.nhwc(gy))
    y = U.conv_fwd(xd, wd, bd, 2, 'SAME')
    U.assert_close(U.nchw(y), torch.from_numpy(d["out_y"]), 1e-3, name + " y")
    dx = U.conv_dgrad(gyd, wd, None, x.shape[2], x.shape[3], 2, 'SAME')
    U.assert_close(U.nchw(dx), torch.from_numpy(d["out_dx"]), 1e-3, name + " dx")
    dw = U.conv_wgrad(xd, gyd, 5, 2, 'SAME')
    U.assert_close(dw, torch.from_numpy(d["out_dw"]), 1e-3, name + " dw")


@pytest.mark.parametrize("name", ["deconv_g2", "deconv_g3", "deconv_g5"])
def test_deconv_golden(name):
    import gpu_util as U
    d = _load(name)
    x, w, b = (torch.from_numpy(d[k]) for k in ("x", "w", "b"))
    H = x.shape[2]
    y = U.conv_dgrad(U.dev(U.nhwc(x)), U.dev(w), U.dev(b), 2 * H, 2 * H, 2, 'SAME')
    U.assert_close(U.nchw(y), torch.from_numpy(d["out_y"]), 1e-3, name)


@pytest.mark.parametrize("name", ["bn_spatial", "bn_dense"])
def test_batchnorm_golden(name):
    import gpu_util as U
    from gg import cabi
    d = _load(name)
    x, gy = torch.from_numpy(d["x"]), torch.from_numpy(d["gy"])
    if x.dim() == 4:
        x2, gy2 = U.nhwc(x).reshape(-1, x.shape[1]), U.nhwc(gy).reshape(-1, x.shape[1])
    else:
        x2, gy2 = x, gy
    R, Cc = x2.shape
    S = cabi.lib.gg_bn_slices(R, Cc)
    xd, gd, bd, gyd = U.dev(x2), U.dev(d["scale"].reshape(-1)), U.dev(d["offset"].reshape(-1)), U.dev(gy2)
    part, part2 = torch.empty(S, 2, Cc, device="cuda"), torch.empty(S, 2, Cc, device="cuda")
    y, dx = torch.empty(R, Cc, device="cuda"), torch.empty(R, Cc, device="cuda")
    mean, rstd, dg, db = (torch.empty(Cc, device="cuda") for _ in range(4))
    st = cabi.stream_ptr()
    cabi.call("gg_bn_stats", xd.data_ptr(), part.data_ptr(), R, Cc, st)
    cabi.call("gg_bn_apply", xd.data_ptr(), part.data_ptr(), S, float(R), gd.data_ptr(), bd.data_ptr(), 1e-5, y.data_ptr(),
              mean.data_ptr(), rstd.data_ptr(), R, Cc, 0, 0.0, st)
    cabi.call("gg_bn_bwd_reduce", gyd.data_ptr(), xd.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gd.data_ptr(), None,
              part2.data_ptr(), R, Cc, 0, 0.0, st)
    cabi.call("gg_bn_bwd_apply", gyd.data_ptr(), xd.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gd.data_ptr(), None,
              part2.data_ptr(), S, float(R), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), R, Cc, 0, 0.0, st)

    def back(t):
        return U.nchw(t.reshape(x.shape[0], x.shape[2], x.shape[3], Cc)) if x.dim() == 4 else t
    U.assert_close(back(y), torch.from_numpy(d["out_y"]), 1e-5, name + " y")
    U.assert_close(back(dx), torch.from_numpy(d["out_dx"]), 1e-4, name + " dx")
    U.assert_close(dg, torch.from_numpy(d["out_dscale"]).reshape(-1), 1e-4, name + " dscale")
    U.assert_close(db, torch.from_numpy(d["out_doffset"]).reshape(-1), 1e-4, name + " doffset")
    # single-kernel forms
    y2, dx2 = torch.empty(R, Cc, device="cuda"), torch.empty(R, Cc, device="cuda")
    mean2, rstd2, dg2, db2 = (torch.empty(Cc, device="cuda") for _ in range(4))
    cabi.call("gg_bn_fwd_fused", xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), 1e-5, y2.data_ptr(), mean2.data_ptr(),
              rstd2.data_ptr(), R, Cc, 0, 0.0, st)
    cabi.call("gg_bn_bwd_fused", gyd.data_ptr(), xd.data_ptr(), y2.data_ptr(), mean2.data_ptr(), rstd2.data_ptr(), gd.data_ptr(),
              dx2.data_ptr(), dg2.data_ptr(), db2.data_ptr(), R, Cc, 0, 0.0, st)
    U.assert_close(back(y2), torch.from_numpy(d["out_y"]), 1e-5, name + " fused y")
    U.assert_close(back(dx2), torch.from_numpy(d["out_dx"]), 1e-4, name + " fused dx")
    U.assert_close(dg2, torch.from_numpy(d["out_dscale"]).reshape(-1), 1e-4, name + " fused dscale")
    U.assert_close(db2, torch.from_numpy(d["out_doffset"]).reshape(-1), 1e-4, name + " fused doffset")


def test_adam_golden():
    import gpu_util as U
    from gg import cabi
    d = _load("adam_5steps")
    p = U.dev(d["p"])
    g, m, v = torch.empty_like(p), torch.zeros_like(p), torch.zeros_like(p)
    tab = torch.frombuffer(bytearray(struct.pack("<QQQQq", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())),
                           dtype=torch.uint8).cuda()
    chk = torch.frombuffer(bytearray(struct.pack("<iiq", 0, 0, 0)), dtype=torch.uint8).cuda()
    state = torch.zeros(3, dtype=torch.float64, device="cuda")
    for gr in d["grads"]:
        g.copy_(torch.from_numpy(gr))
        cabi.call("gg_adam_multi", tab.data_ptr(), chk.data_ptr(), 1, state.data_ptr(), float(d["lr"]), float(d["beta1"]),
                  float(d["beta2"]), 1e-8, 1.0, cabi.stream_ptr())
    U.assert_close(p, torch.from_numpy(d["out_p"]), 1e-5, "adam golden")


def test_losses_golden():
    import gpu_util as U
    from gg import cabi
    d = _load("losses")
    st = cabi.stream_ptr()
    df, dr = U.dev(d["disc_fake"]), U.dev(d["disc_real"])
    gen, disc = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    for i in range(2):                                  # local_ep over two (fake, real) logit pairs, then /2
        for out, lf, lr in ((gen, 1.0, 0.0), (disc, 0.0, 1.0)):
            cabi.call("gg_bce_mean", df[i].data_ptr(), 64, lf, 0.5, out.data_ptr(), 1, st)
            cabi.call("gg_bce_mean", dr[i].data_ptr(), 64, lr, 0.5, out.data_ptr(), 1, st)
    U.assert_close(gen, torch.from_numpy(d["out_gen"]).reshape(1), 1e-5, "local_ep gen")
    U.assert_close(disc, torch.from_numpy(d["out_disc"]).reshape(1), 1e-5, "local_ep disc")
    a, b2 = U.dev(d["a"]), U.dev(d["b2"])
    l2 = torch.zeros(1, device="cuda")
    cabi.call("gg_dist_mean", a.data_ptr(), b2.data_ptr(), a.numel(), 2, 1.0, l2.data_ptr(), 0, st)
    U.assert_close(l2, torch.from_numpy(d["out_l2"]).reshape(1), 1e-5, "l2")
    g = a * 0.05
    slopes, gp = torch.empty(64, device="cuda"), torch.zeros(1, device="cuda")
    cabi.call("gg_gp_slope_penalty", g.data_ptr(), 64, 128, 10.0, slopes.data_ptr(), gp.data_ptr(), st)
    U.assert_close(gp, torch.from_numpy(d["out_gp"]).reshape(1), 1e-5, "gp")
