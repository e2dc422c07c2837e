"""-m gpu: hardware-behaviour probes the tcgen05 implicit-GEMM convolution relies on (csrc/gg_probe.cu).

A report with the measured numbers is appended to gpurun_out/probes.txt when that directory exists.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _log(msg):
    d = os.path.join(ROOT, "gpurun_out")
    print(msg)
    if os.path.isdir(d):
        with open(os.path.join(d, "probes.txt"), "a") as f:
            f.write(msg + "\n")


def _deswizzle128(raw, rows):
    """raw: smem image of `rows` rows x 32 fp32 written with the 128B swizzle -> logical [rows,32]."""
    raw = raw.reshape(rows, 8, 4)            # 8 chunks of 16 bytes per 128-byte row
    out = np.empty_like(raw)
    for r in range(rows):
        for c in range(8):
            out[r, c] = raw[r, c ^ (r % 8)]
    return out.reshape(rows, 32)


@pytest.mark.parametrize("swz", [0, 1])
@pytest.mark.parametrize("h0,w0", [(-1, -1), (1, 1), (3, -1), (9, 11)])
def test_tma_strided_box_gathers_conv_tap(swz, h0, w0):
    import gpu_util as U
    from gg import cabi
    B, H, W, Cc = 3, 16, 16, 64
    rs = np.random.RandomState(0)
    x = rs.randn(B, H, W, Cc).astype(np.float32)
    hb, wb, b, c0 = 8, 8, 1, 32
    out = torch.empty(hb * wb * 32, device="cuda")
    xd = U.dev(x)
    cabi.call("gg_probe_tma_strided", cabi.ptr(xd), B, H, W, Cc, b, h0, w0, c0, hb, wb, swz, cabi.ptr(out),
              cabi.stream_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    got = _deswizzle128(got, hb * wb) if swz else got.reshape(hb * wb, 32)
    exp = np.zeros((hb, wb, 32), np.float32)
    for i in range(hb):
        for j in range(wb):
            h, w = h0 + 2 * i, w0 + 2 * j
            if 0 <= h < H and 0 <= w < W:
                exp[i, j] = x[b, h, w, c0:c0 + 32]
    exp = exp.reshape(hb * wb, 32)
    nbad = int((got != exp).sum())
    _log("tma_strided swz=%d h0=%d w0=%d mismatches=%d first_row_got=%s first_row_exp=%s" %
         (swz, h0, w0, nbad, got[0, :4], exp[0, :4]))
    assert nbad == 0


def _round_tf32(a, mode):
    """fp32 -> tf32 (10 explicit mantissa bits) by truncation ('trunc') or round-to-nearest-away ('rna')."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    if mode == 'rna':
        u = u + 0x1000
    u = (u & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("N,K", [(128, 64), (64, 32), (256, 96)])
@pytest.mark.parametrize("cv", [0, 1])
def test_umma_tf32_operand_layouts(a_mn, b_mn, N, K, cv):
    import gpu_util as U
    from gg import cabi
    rs = np.random.RandomState(N + K + a_mn * 2 + b_mn)
    A = rs.randn(128, K).astype(np.float32)
    Bm = rs.randn(K, N).astype(np.float32)
    A_st = np.ascontiguousarray(A.T) if a_mn else A          # MN-major A is stored [K,128]
    B_st = Bm if b_mn else np.ascontiguousarray(Bm.T)        # K-major B is stored [N,K]
    D = torch.full((128, N), float("nan"), device="cuda")
    Ad, Bd = U.dev(A_st), U.dev(B_st)    # keep the device buffers alive across the call
    cabi.call("gg_probe_umma_tf32", cabi.ptr(Ad), cabi.ptr(Bd), cabi.ptr(D), N, K, a_mn, b_mn, cv,
              cabi.stream_ptr())
    torch.cuda.synchronize()
    got = D.cpu().numpy().astype(np.float64)
    exact = A.astype(np.float64) @ Bm.astype(np.float64)
    scale = np.abs(exact).max()
    errs = {}
    for mode in ("trunc", "rna"):
        ref = _round_tf32(A, mode).astype(np.float64) @ _round_tf32(Bm, mode).astype(np.float64)
        errs[mode] = float(np.abs(got - ref).max() / scale)
    errs["exact"] = float(np.abs(got - exact).max() / scale)
    bias = float(((got - exact) * np.sign(exact)).mean() / np.abs(exact).mean())
    _log("umma_tf32 a_mn=%d b_mn=%d N=%d K=%d tma_cvt=%d err_vs_trunc=%.2e err_vs_rna=%.2e err_vs_exact=%.2e signed_bias=%.2e" %
         (a_mn, b_mn, N, K, cv, errs["trunc"], errs["rna"], errs["exact"], bias))
    assert errs["exact"] < 2e-3, errs
