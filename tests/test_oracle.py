"""CPU: pins the oracle restatement (oracle/tf_ops.py) against independent loop implementations of the documented
TensorFlow semantics, internal identities, and the committed golden vectors (tests/golden/*.npz, written by
tests/golden/make_golden.py).  The reference has no tests or fixtures of its own: parity is unpinned
by the reference, these identities are what anchors the oracle."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import tf_ops as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _conv2d_loops(x, w, stride, padding):
    """direct NCHW cross-correlation with TF SAME/VALID padding arithmetic, plain loops (small cases only)"""
    B, Ci, H, W = x.shape
    k, _, _, Co = w.shape
    if padding == 'SAME':
        Ho, Wo = -(-H // stride), -(-W // stride)
        pt = max((Ho - 1) * stride + k - H, 0) // 2
        pl = max((Wo - 1) * stride + k - W, 0) // 2
    else:
        Ho, Wo, pt, pl = (H - k) // stride + 1, (W - k) // stride + 1, 0, 0
    y = np.zeros((B, Co, Ho, Wo))
    for b in range(B):
        for ho in range(Ho):
            for wo in range(Wo):
                for r in range(k):
                    for s in range(k):
                        hi, wi = ho * stride + r - pt, wo * stride + s - pl
                        if 0 <= hi < H and 0 <= wi < W:
                            y[b, :, ho, wo] += x[b, :, hi, wi] @ w[r, s]
    return y


@pytest.mark.parametrize("H,W,k,stride,padding", [(8, 8, 5, 2, 'SAME'), (7, 7, 5, 2, 'SAME'), (6, 5, 3, 1, 'SAME'),
                                                  (6, 6, 3, 2, 'SAME'), (7, 7, 4, 1, 'VALID'), (4, 4, 5, 2, 'SAME')])
def test_conv2d_matches_loop_implementation(H, W, k, stride, padding):
    rs = np.random.RandomState(0)
    x, w = rs.randn(2, 3, H, W), rs.randn(k, k, 3, 4)
    got = O.conv2d(torch.tensor(x), torch.tensor(w), stride, padding).numpy()
    np.testing.assert_allclose(got, _conv2d_loops(x, w, stride, padding), rtol=1e-12, atol=1e-12)


def test_same_padding_is_asymmetric_for_stride_2():
    assert O.same_padding(32, 5, 2) == (16, 1, 2)      # (out, before, after): NOT the symmetric (2,2)
    assert O.same_padding(7, 5, 2) == (4, 2, 2)
    assert O.same_padding(28, 5, 2) == (14, 1, 2)
    assert O.same_padding(16, 3, 1) == (16, 1, 1)


@pytest.mark.parametrize("H,k,stride", [(4, 5, 2), (8, 5, 2), (7, 5, 2), (8, 3, 2)])
def test_conv2d_transpose_is_input_gradient_of_conv2d(H, k, stride):
    g = torch.Generator().manual_seed(1)
    Cin, Cout = 5, 3                                   # deconv: Cin -> Cout, filter (k,k,Cout,Cin)
    x = torch.randn(2, Cin, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(k, k, Cout, Cin, generator=g, dtype=torch.float64)
    y = O.conv2d_transpose(x, w, stride, 'SAME')
    assert y.shape == (2, Cout, stride * H, stride * H)
    z = torch.zeros(2, Cout, stride * H, stride * H, dtype=torch.float64, requires_grad=True)
    fwd = O.conv2d(z, w, stride, 'SAME')               # the mirrored conv Cout -> Cin
    (dz,) = torch.autograd.grad(fwd, z, x)
    np.testing.assert_allclose(y.numpy(), dz.numpy(), rtol=1e-12, atol=1e-12)


def test_batchnorm_definition():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(6, 4, 3, 3, generator=g, dtype=torch.float64)
    sc, of = torch.rand(4, generator=g, dtype=torch.float64) + .5, torch.randn(4, generator=g, dtype=torch.float64)
    y = O.batchnorm(x, sc, of, [0, 2, 3]).numpy()
    xn = x.numpy()
    for c in range(4):
        m, v = xn[:, c].mean(), xn[:, c].var()         # biased variance
        np.testing.assert_allclose(y[:, c], (xn[:, c] - m) / math.sqrt(v + 1e-5) * sc[c].item() + of[c].item(), rtol=1e-10)
    x2 = torch.randn(5, 7, generator=g, dtype=torch.float64)
    y2 = O.batchnorm(x2, torch.ones(1, 7, dtype=torch.float64), torch.zeros(1, 7, dtype=torch.float64), [0]).numpy()
    np.testing.assert_allclose(y2.mean(0), 0, atol=1e-12)


def test_sigmoid_cross_entropy_and_local_ep():
    x = torch.tensor([-30.0, -1.0, 0.0, 2.0, 40.0], dtype=torch.float64)
    for z in (0.0, 1.0):
        ref = -(z * torch.log(torch.sigmoid(x)) + (1 - z) * torch.log(1 - torch.sigmoid(x)))
        got = O.sigmoid_cross_entropy_with_logits(x, torch.full_like(x, z))
        np.testing.assert_allclose(got[1:4].numpy(), ref[1:4].numpy(), rtol=1e-10)
        assert torch.isfinite(got).all()               # stable form does not overflow at |x| = 40
    df, dr = [torch.tensor([0.3, -0.2]), torch.tensor([1.0, 2.0])], [torch.tensor([0.1, 0.4]), torch.tensor([-1.0, 0.5])]
    gen, disc = O.local_ep_costs(df, dr)
    manual_gen = sum(O.bce_mean(a, 1.0) + O.bce_mean(b, 0.0) for a, b in zip(df, dr)) / 2
    assert abs(float(gen) - float(manual_gen)) < 1e-7 and float(disc) > 0


def test_tf_adam_scalar_hand_computation():
    p = torch.tensor([1.0], dtype=torch.float64)
    opt = O.TFAdam([p], lr=0.1, beta1=0.5, beta2=0.9, eps=1e-8)
    m = v = 0.0
    ref = 1.0
    for t, g in enumerate([0.5, -0.25, 2.0], start=1):
        opt.step([torch.tensor([g], dtype=torch.float64)])
        m = 0.5 * m + 0.5 * g
        v = 0.9 * v + 0.1 * g * g
        lr_t = 0.1 * math.sqrt(1 - 0.9 ** t) / (1 - 0.5 ** t)
        ref -= lr_t * m / (math.sqrt(v) + 1e-8)        # epsilon OUTSIDE the bias correction (TensorFlow form)
        assert abs(p.item() - ref) < 1e-12


def test_initialisers_follow_reference_formulas():
    rs = np.random.RandomState(3)
    w = O.init_conv2d(rs, 64, 128, 5, stride=2)
    bound = math.sqrt(4. / (64 * 25 + 128 * 25 / 4)) * math.sqrt(3)
    assert w.shape == (5, 5, 64, 128) and w.dtype == np.float32 and np.abs(w).max() <= bound
    wd = O.init_deconv2d(rs, 256, 128, 5)
    assert wd.shape == (5, 5, 128, 256)
    wl = O.init_linear(rs, 128, 4096)
    assert wl.shape == (128, 4096) and np.abs(wl).max() <= math.sqrt(2. / (128 + 4096)) * math.sqrt(3)


GOLDEN = sorted(f for f in os.listdir(GOLD) if f.endswith(".npz") and not f.startswith("trajectory")) if os.path.isdir(GOLD) else []


@pytest.mark.parametrize("fname", GOLDEN)
def test_oracle_reproduces_committed_golden_vectors(fname):
    from tests.golden import make_golden as MG
    d = np.load(os.path.join(GOLD, fname))
    out = MG.compute(str(d["kind"]), {k: d[k] for k in d.files})
    for k, v in out.items():
        np.testing.assert_allclose(v, d["out_" + k], rtol=1e-5, atol=1e-6, err_msg="%s:%s" % (fname, k))


def test_oracle_reproduces_first_iterations_of_committed_trajectory():
    """tests/golden/trajectory100.npz (make_trajectory.py): re-run the first 3 iterations of the fp64 oracle from the
    tflib-initialised weights and compare cost curve and the iteration-1/2 fixed-noise samples with the committed file"""
    import torch
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    from oracle import gmgan_cifar10 as OM
    gold = np.load(os.path.join(GOLD, "trajectory100.npz"))
    B, n_keep = int(gold["batch"]), int(gold["n_keep"])
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=B)
    oracle = OM.GMGANCifar10({n: p.attrs["init"] for n, p in lib._params.items()}, dtype=torch.float64)
    k1h, noise = g.np_fixed_k.astype(np.float32), g.np_fixed_noise
    step = 0
    for it in range(2):
        if it > 0:
            gc, _ = oracle.gen_step(**OM.synthetic_inputs(B, step)); step += 1
            assert abs(gc - gold["gen_costs"][it]) < 1e-9
        dc, _ = oracle.disc_step(**OM.synthetic_inputs(B, step)); step += 1
        assert abs(dc - gold["disc_costs"][it]) < 1e-9
        s = oracle.sample(k1h, noise).numpy().astype(np.float32)[:n_keep]
        assert np.abs(s - gold["samples"][it]).max() < 1e-6
    assert list(gold["checkpoints"]) == [1, 2, 5, 10, 20, 50, 100] and np.isfinite(gold["disc_costs"]).all()
