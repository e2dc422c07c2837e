"""-m gpu: the WGAN-GP rows (gan_inference_svhn.py 'wali-gp' and 'vegan-wgan-gp'): costs, gradient penalty and the
second-order discriminator gradients through the C-ABI path vs the fp64 oracle (torch double backward)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


def _fresh(mode, batch):
    import tensorflow as tf
    import tflib as lib
    import gan_inference_svhn as S
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(4321)
    return tf, lib, S.build_graph(MODE=mode, BATCH_SIZE=batch)


def _params(lib):
    from gg.executor import RT
    return {n: RT.get_param(p).copy() for n, p in lib._params.items()}


def _check(got, ref, tol_l2, what):
    got, ref = np.asarray(got, np.float64), ref.detach().numpy()
    l2 = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30)
    assert got.shape == ref.shape and l2 < tol_l2, "%s: rel-L2 %.3e" % (what, l2)
    return l2


@pytest.mark.parametrize("backend", [1, 0])
def test_wali_gp_double_backward(backend):
    from oracle.gan_svhn import GanSvhn
    from gg import cabi
    B = 32
    tf, lib, g = _fresh('wali-gp', B)
    cabi.call("gg_set_conv_backend", backend)
    try:
        sess = tf.Session()
        oracle = GanSvhn(_params(lib), 'wali-gp')
        rs = np.random.RandomState(5)
        x = rs.randint(0, 256, size=(B, 3072)).astype(np.int32)
        pz = rs.randn(B, 128).astype(np.float32)
        al = rs.uniform(0, 1, size=(B, 1)).astype(np.float32)
        feeds = {g.real_x_int: x, g.p_z: pz, g.alpha: al}
        dparams = g.disc_params
        gparams = [p for p in g.gen_params + g.ext_params]
        dgrads = tf.gradients(g.disc_cost, dparams)
        ggrads = tf.gradients(g.gen_cost, gparams)
        out = sess.run([g.gen_cost, g.disc_cost, g.gradient_penalty] + dgrads + ggrads, feed_dict=feeds)
        ref_gen, ref_disc, ref_gp = oracle.wali_gp(x, pz, al)
        tol_c = 2e-3 if backend == 1 else 2e-2
        assert abs(out[2] - float(ref_gp)) <= tol_c * max(abs(float(ref_gp)), 1e-3), (out[2], float(ref_gp))
        assert abs(out[1] - float(ref_disc)) <= tol_c * max(abs(float(ref_disc)), 1.0)
        assert abs(out[0] - float(ref_gen)) <= tol_c * max(abs(float(ref_gen)), 1.0)
        rd = oracle.grads(ref_disc, [p.name for p in dparams])
        rg = oracle.grads(ref_gen, [p.name for p in gparams])
        tol = 5e-3 if backend == 1 else 8e-2
        worst = {}
        for p, got in zip(dparams, out[3:3 + len(dparams)]):
            worst[p.name] = _check(got, rd[p.name], tol, "disc grad " + p.name)
        for p, got in zip(gparams, out[3 + len(dparams):]):
            worst[p.name] = _check(got, rg[p.name], tol, "gen grad " + p.name)
        print("wali-gp backend", backend, "gp", out[2], "worst rel-L2:", sorted(worst.items(), key=lambda kv: -kv[1])[:4])
        # and the train ops run (5 critic steps per generator step, Adam(1e-4, .5, .9))
        for _ in range(2):
            c, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)
            assert np.isfinite(c)
        c, _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict=feeds)
        assert np.isfinite(c)
    finally:
        cabi.call("gg_set_conv_backend", 0)


def test_vegan_wgan_gp_mlp_critic():
    from oracle.gan_svhn import GanSvhn
    B = 32
    tf, lib, g = _fresh('vegan-wgan-gp', B)
    sess = tf.Session()
    oracle = GanSvhn(_params(lib), 'vegan-wgan-gp')
    rs = np.random.RandomState(6)
    x = rs.randint(0, 256, size=(B, 3072)).astype(np.int32)
    pz = rs.randn(B, 8).astype(np.float32)
    al = rs.uniform(0, 1, size=(B, 1)).astype(np.float32)
    widths = [8, 1024, 512, 256]
    stds = [.3, .5, .5, .5]
    noises = [[(rs.randn(B, w) * s).astype(np.float32) for w, s in zip(widths, stds)] for _ in range(3)]
    assert len(g.noise_layers) == 12                 # three critic calls (real, fake, interpolates) x four noise layers
    feeds = {g.real_x_int: x, g.p_z: pz, g.alpha: al}
    for call in range(3):
        for i in range(4):
            feeds[g.noise_layers[call * 4 + i]] = noises[call][i]
    dgrads = tf.gradients(g.disc_cost, g.disc_params)
    out = sess.run([g.gen_cost, g.disc_cost, g.gradient_penalty] + dgrads, feed_dict=feeds)
    ref_gen, ref_disc, ref_gp = oracle.vegan_wgan_gp(x, pz, al, noises)
    assert abs(out[2] - float(ref_gp)) <= 5e-3 * abs(float(ref_gp))
    assert abs(out[1] - float(ref_disc)) <= 5e-3 * max(abs(float(ref_disc)), 1.0)
    assert abs(out[0] - float(ref_gen)) <= 2e-2 * max(abs(float(ref_gen)), 1.0)
    rd = oracle.grads(ref_disc, [p.name for p in g.disc_params])
    for p, got in zip(g.disc_params, out[3:]):
        _check(got, rd[p.name], 6e-2, "disc grad " + p.name)     # tf32 dense layers, second order (measured 2.3e-2)


@pytest.mark.parametrize("mode", ["ali", "alice", "vegan", "wali"])
def test_other_modes_train_without_nans(mode):
    B = 16
    tf, lib, g = _fresh(mode, B)
    sess = tf.Session()
    rs = np.random.RandomState(7)
    for it in range(2):
        x = rs.randint(0, 256, size=(B, 3072)).astype(np.int32)
        dc, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: x})
        if mode == 'wali':
            sess.run(g.clip_disc_weights)
            from gg.executor import RT
            w = RT.get_param([p for p in g.disc_params if p.name == 'Discriminator.zx1.W'][0])
            assert np.abs(w).max() <= 0.01 + 1e-7                                   # weight clipping, gan_inference.py:16-24
        gc, _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: x})
        assert np.isfinite(dc) and np.isfinite(gc)
