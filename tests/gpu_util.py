"""Helpers for the -m gpu parity tests: call the C-ABI on torch device buffers."""
import ctypes as C

import numpy as np
import torch

from gg import cabi
from oracle import tf_ops as O

DEV = "cuda"


def dev(a, dtype=torch.float32):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(a)
    return a.to(dtype).contiguous().to(DEV)


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def geom(H, W, k, stride, padding):
    Ho, Wo, pt, pb, pl, pr = O.conv_geometry(H, W, k, stride, padding)
    return Ho, Wo, pt, pl


def ws(nbytes):
    n = max(int(nbytes), 256)
    return torch.zeros(n, dtype=torch.uint8, device=DEV)   # zero once: arrival tickets of the split-K conv kernels


def conv_fwd(x_nhwc, w, bias, stride, padding, act=None, alpha=0.2):
    B, H, W, Ci = x_nhwc.shape
    k, _, _, Co = w.shape
    Ho, Wo, pt, pl = geom(H, W, k, stride, padding)
    y = torch.empty(B, Ho, Wo, Co, device=DEV)
    wsp = ws(cabi.lib.gg_conv2d_workspace(0, B, H, W, Ci, Co, k, stride, Ho, Wo))
    cabi.call("gg_conv2d_fwd", cabi.ptr(x_nhwc), cabi.ptr(w), cabi.ptr(bias), cabi.ptr(y), B, H, W, Ci, Co, k, stride, pt, pl,
              Ho, Wo, cabi.ACT[act], alpha, cabi.ptr(wsp), wsp.numel(), cabi.stream_ptr())
    return y


def conv_dgrad(dy_nhwc, w, bias, H, W, stride, padding, act=None, alpha=0.2):
    B, Ho, Wo, Co = dy_nhwc.shape
    k, _, Ci, _ = w.shape
    Ho2, Wo2, pt, pl = geom(H, W, k, stride, padding)
    assert (Ho2, Wo2) == (Ho, Wo)
    dx = torch.empty(B, H, W, Ci, device=DEV)
    wsp = ws(cabi.lib.gg_conv2d_workspace(1, B, H, W, Ci, Co, k, stride, Ho, Wo))
    cabi.call("gg_conv2d_dgrad", cabi.ptr(dy_nhwc), cabi.ptr(w), cabi.ptr(bias), cabi.ptr(dx), B, H, W, Ci, Co, k, stride, pt,
              pl, Ho, Wo, cabi.ACT[act], alpha, cabi.ptr(wsp), wsp.numel(), cabi.stream_ptr())
    return dx


def conv_wgrad(x_nhwc, dy_nhwc, k, stride, padding):
    B, H, W, Ci = x_nhwc.shape
    _, Ho, Wo, Co = dy_nhwc.shape
    Ho2, Wo2, pt, pl = geom(H, W, k, stride, padding)
    assert (Ho2, Wo2) == (Ho, Wo)
    dw = torch.empty(k, k, Ci, Co, device=DEV)
    need = cabi.lib.gg_conv2d_wgrad_workspace(B, H, W, Ci, Co, k, stride, Ho, Wo)
    wsp = ws(need)
    cabi.call("gg_conv2d_wgrad", cabi.ptr(x_nhwc), cabi.ptr(dy_nhwc), cabi.ptr(dw), B, H, W, Ci, Co, k, stride, pt, pl, Ho, Wo,
              cabi.ptr(wsp), wsp.numel(), cabi.stream_ptr())
    return dw


def gemm(A, Bm, bias, M, N, K, ta=0, tb=0, act=None, alpha=0.2):
    Cm = torch.empty(M, N, device=DEV)
    need = cabi.lib.gg_gemm_workspace(M, N, K)
    wsp = ws(need)
    cabi.call("gg_gemm", cabi.ptr(A), cabi.ptr(Bm), cabi.ptr(bias), cabi.ptr(Cm), M, N, K, ta, tb, cabi.ACT[act], alpha,
              cabi.ptr(wsp), wsp.numel(), cabi.stream_ptr())
    return Cm


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, rtol=1e-3, what=""):
    """|a-b| <= rtol * (|b| + max|b|): the 1e-3 relative fp32 tolerance of BASELINE.json's north_star, with the
    tensor's own scale as the absolute floor (element-wise relative error is undefined at zero crossings)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bound = rtol * (b.abs() + b.abs().max())
    bad = (a - b).abs() > bound
    assert not bool(bad.any()), "%s: %d/%d elements out of tolerance, max rel-to-scale err %.3e" % (
        what, int(bad.sum()), bad.numel(), rel_err(a, b))
