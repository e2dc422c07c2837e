"""-m gpu: the other model families of the hot path run end to end on the CUDA kernels — SSGAN moving-MNIST (state-space
latent, B*LEN frame folding, weighted_local_epce; BASELINE.json configs[4] at seq-len 8 / bs 32) and ALI on 64x64 faces
(configs[3] per-rank shape) — and their objectives agree with the oracle's algebra evaluated on the fetched logits."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


def test_ssgan_moving_mnist_len8_bs32():
    import tensorflow as tf
    import tflib as lib
    import ssgan_inference_moving_mnist as S
    from oracle import tf_ops as O
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(11)
    B, LEN = 32, 8
    g = S.build_graph(BATCH_SIZE=B, LEN=LEN)
    assert len(g.disc_fake) == LEN + 1 and abs(g.ratio.sum() - 1.0) < 1e-12 and g.fake_x.shape == (B, LEN, 4096)
    sess = tf.Session()
    rs = np.random.RandomState(3)
    x = rs.uniform(0, 1, size=(B, LEN, 4096)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rs.randint(0, 10, size=B)]
    feeds = {g.real_x_unit: x, g.real_y: y, g.p_z_l_0: rs.randn(B, 8).astype(np.float32), g.p_z_g: rs.randn(B, 128).astype(np.float32),
             g.epsilons[0]: rs.randn(B, 8).astype(np.float32), g.p_y_idx: rs.randint(0, 10, size=B).astype(np.int32)}
    out = sess.run([g.gen_cost, g.disc_cost] + g.disc_fake + g.disc_real, feed_dict=feeds)
    df = [torch.from_numpy(np.asarray(v, np.float64)) for v in out[2:2 + LEN + 1]]
    dr = [torch.from_numpy(np.asarray(v, np.float64)) for v in out[2 + LEN + 1:]]
    assert df[-1].shape == (B * LEN,) and df[0].shape == (B,)          # frame discriminator sees B*LEN images
    ref_gen, ref_disc = O.weighted_local_epce_costs(df, dr, g.ratio)
    assert abs(out[0] - float(ref_gen)) < 1e-4 * max(1.0, abs(float(ref_gen)))
    assert abs(out[1] - float(ref_disc)) < 1e-4 * max(1.0, abs(float(ref_disc)))
    # the shared-epsilon recurrence: the same latent trajectory for the same (z_0, epsilon)
    a = sess.run(g.p_z_l, feed_dict=feeds)
    b = sess.run(g.p_z_l, feed_dict=feeds)
    assert a.shape == (B, LEN, 8) and np.array_equal(a, b) and np.array_equal(a[:, 0, :], feeds[g.p_z_l_0])
    costs = []
    for it in range(3):
        dc, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)
        gc, _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict=feeds)
        costs.append((float(dc), float(gc)))
    assert np.isfinite(costs).all() and costs[2][0] < costs[0][0]        # D improves on a fixed batch


def test_face_ali_64x64_shard():
    import tensorflow as tf
    import tflib as lib
    import gan_inference_face as S
    from oracle import tf_ops as O
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(12)
    B = 16                                                               # the 8-GPU shard of the reference's bs=128
    g = S.build_graph(BATCH_SIZE=B)
    sess = tf.Session()
    rs = np.random.RandomState(4)
    xi = rs.randint(0, 256, size=(B, 12288)).astype(np.int32)
    dq = rs.uniform(0, 1. / 128, size=(B, 12288)).astype(np.float32)
    feeds = {g.real_x_int: xi, g.dequant: dq, g.p_z: rs.randn(B, 128).astype(np.float32)}
    rx, gc, dc, df, dr = sess.run([g.real_x, g.gen_cost, g.disc_cost, g.disc_fake, g.disc_real], feed_dict=feeds)
    ref_x = (2 * ((xi.astype(np.float32) / np.float32(256.)) - np.float32(.5))).astype(np.float32) + dq
    assert np.array_equal(rx, ref_x)                                     # int -> float decode + dequantisation: bit exact
    ref_gen, ref_disc = O.ali_costs(torch.from_numpy(df.astype(np.float64)), torch.from_numpy(dr.astype(np.float64)))
    # the costs come from the sibling-batched tower, the fetched logits from the script's own two towers: same weights and
    # inputs, different tf32 split-K partitions (3e-4 per conv) -> 1e-4 on the mean, like the SSGAN check above
    assert abs(gc - float(ref_gen)) < 1e-4 * max(1, abs(float(ref_gen))) and abs(dc - float(ref_disc)) < 1e-4 * max(1, abs(float(ref_disc)))
    for it in range(2):
        d, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)
        c, _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict=feeds)
        assert np.isfinite(d) and np.isfinite(c)
    s = sess.run(g.fake_x, feed_dict=feeds)
    assert s.shape == (B, 12288) and np.abs(s).max() <= 1.0


def test_checkpoint_save_restore_roundtrip(tmp_path):
    import tensorflow as tf
    import tflib as lib
    import gan_inference_face as S
    from gg.executor import RT
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(13)
    g = S.build_graph(BATCH_SIZE=8)
    sess = tf.Session()
    rs = np.random.RandomState(5)
    feeds = {g.real_x_int: rs.randint(0, 256, size=(8, 12288)).astype(np.int32), g.dequant: np.zeros((8, 12288), np.float32),
             g.p_z: rs.randn(8, 128).astype(np.float32)}
    sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)
    saver = tf.train.Saver()
    path = saver.save(sess, str(tmp_path / "model.ckpt"))
    w = lib._params['Discriminator.zx1.W']
    before = RT.get_param(w).copy()
    c1, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)
    assert not np.array_equal(before, RT.get_param(w))
    saver.restore(sess, path)
    assert np.array_equal(before, RT.get_param(w))
    c2, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds)     # same params AND same Adam state -> same step
    assert c1 == c2


def test_gmgan_mnist_local_ep_bs50_matches_oracle():
    """BASELINE.json configs[0] (gmgan_inference_mnist.py MODE='local_ep', BATCH_SIZE=50, 1x28x28): costs and
    every parameter gradient of the D step and the G step against the fp64 oracle (oracle/gmgan_mnist.py) — exercises the
    28 -> 14 -> 7 -> 4 convolutions (pad (2,2) on the last), the 8x8 -> 7x7 crop between deconvolutions, Cin = Cout = 1
    layers and a batch of 50 (no power of two anywhere) — then two training iterations."""
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_mnist as S
    from gg import cabi
    from gg.executor import RT
    from oracle import gmgan_mnist as OM
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(21)
    B = 50
    g = S.build_graph(BATCH_SIZE=B)
    sess = tf.Session()
    params = {name: RT.get_param(p).copy() for name, p in lib._params.items()}
    oracle = OM.GMGANMnist(params, dtype=torch.float64)
    inp = OM.synthetic_inputs(B, 0)
    feeds = {g.real_x: inp["real_x"], g.hyper_p_z: inp["hyper_p_z"], g.hyper_p_k_idx: inp["k_idx"], g.gumbel_uniforms[0]: inp["U"]}
    cabi.call("gg_set_conv_backend", 1)           # fp32 kernels: the tight end-to-end bound (see test_gpu_gmgan_step.py)
    try:
        for which, cost, plist, fn in (("disc", g.disc_cost, g.disc_params, oracle.disc_step),
                                       ("gen", g.gen_cost, g.gen_params + g.ext_params, oracle.gen_step)):
            plist = [p for p in plist if 'moving_' not in p.name]
            grads = tf.gradients(cost, plist)
            keep = [(p, gr) for p, gr in zip(plist, grads) if gr is not None]
            out = sess.run([cost] + [gr for _, gr in keep], feed_dict=feeds)
            ref_cost, ref_grads = fn(apply=False, **inp)
            assert abs(float(out[0]) - ref_cost) <= 1e-4 * max(1.0, abs(ref_cost)), (which, float(out[0]), ref_cost)
            gmax = max(float(v.abs().max()) for v in ref_grads.values() if v is not None)
            for (p, _), got in zip(keep, out[1:]):
                ref = ref_grads[p.name].numpy()
                assert got.shape == ref.shape
                if np.abs(ref).max() < 1e-7 * gmax:
                    assert np.abs(got).max() < 1e-4 * gmax, p.name
                    continue
                l2 = float(np.linalg.norm(got.astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-30))
                assert l2 < 5e-3, "%s grad of %s: rel-L2 %.3e" % (which, p.name, l2)
    finally:
        cabi.call("gg_set_conv_backend", 0)
    # default (tensor-core) path: two training iterations against the oracle's costs
    step = 0
    for it in range(2):
        i2 = OM.synthetic_inputs(B, 10 + step); step += 1
        f2 = {g.real_x: i2["real_x"], g.hyper_p_z: i2["hyper_p_z"], g.hyper_p_k_idx: i2["k_idx"], g.gumbel_uniforms[0]: i2["U"]}
        dc, _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=f2)
        rc, _ = oracle.disc_step(**i2)
        assert abs(float(dc) - rc) <= 5e-3 * max(1.0, abs(rc)), (it, float(dc), rc)
        i3 = OM.synthetic_inputs(B, 10 + step); step += 1
        f3 = {g.real_x: i3["real_x"], g.hyper_p_z: i3["hyper_p_z"], g.hyper_p_k_idx: i3["k_idx"], g.gumbel_uniforms[0]: i3["U"]}
        gc, _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict=f3)
        rg, _ = oracle.gen_step(**i3)
        assert abs(float(gc) - rg) <= 5e-3 * max(1.0, abs(rg)), (it, float(gc), rg)
    s = sess.run(g.fake_x, feed_dict=f3)
    assert s.shape == (B, 784) and s.min() >= 0.0 and s.max() <= 1.0
