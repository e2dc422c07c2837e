"""CPU: host-logic fixes of round 2 against the host "device" of test_cpu_plan.py — checkpoint restore before
the first Session.run, TF's RMSProp slot initialisation, per-rank RNG keys under data parallelism, Python-2 integer
division in the initialiser fan computation (draw-for-draw init compatibility with the reference)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))

from test_cpu_plan import cpu_device, _gmgan   # noqa: F401  (fixture)


def test_saver_restore_before_first_run_restores_params_and_optimizer_state(cpu_device, tmp_path):
    import tensorflow as tf
    import tflib as lib
    from gg import executor
    g = _gmgan(8)
    RT = executor.RT
    plan = executor.Plan(RT, [g.disc_cost, g.disc_train_op], [g.real_x_int])
    assert RT.slots and RT.opt_state
    # pretend a few steps happened
    for k, (m, v) in RT.slots.items():
        m.fill_(0.25 + k[0]); v.fill_(0.5)
    for k, st in RT.opt_state.items():
        st.copy_(torch.tensor([0.125, 0.99, 7.0], dtype=torch.float64))
    w = lib._params['Discriminator.zx1.W']
    RT.param_buffer(w).fill_(3.0)
    n_params_saved = len(lib._params)
    path = tf.train.Saver().save(None, str(tmp_path / "ck.pt"))
    saved = torch.load(path)
    assert len(saved["params"]) == n_params_saved          # every registry variable, touched by a plan or not
    assert any("moving_mean" in k for k in saved["params"])
    slot_keys = set(saved["slots"])

    g2 = _gmgan(8)                                           # fresh process stand-in: new runtime, nothing compiled yet
    RT2 = executor.RT
    assert not RT2.params and not RT2.slots
    tf.train.Saver().restore(None, path)
    w2 = lib._params['Discriminator.zx1.W']
    assert float(RT2.param_buffer(w2)[0]) == 3.0            # parameters land although no plan exists yet
    assert set(RT2.pending_restore["slots"]) == slot_keys
    plan2 = executor.Plan(RT2, [g2.disc_cost, g2.disc_train_op], [g2.real_x_int])
    assert RT2.slots and not RT2.pending_restore["slots"]
    for k, (m, v) in RT2.slots.items():
        assert float(m[0]) == 0.25 + k[0] and float(v[0]) == 0.5
    for st in RT2.opt_state.values():
        assert st.tolist() == [0.125, 0.99, 7.0]


def test_saver_restore_raises_on_mismatched_graph(cpu_device, tmp_path):
    import tensorflow as tf
    import tflib as lib
    _gmgan(8)
    path = tf.train.Saver().save(None, str(tmp_path / "ck.pt"))
    _gmgan(8)
    lib.param('Extra.W', np.zeros((2, 2), np.float32))
    with pytest.raises(KeyError):
        tf.train.Saver().restore(None, path)


def test_rmsprop_rms_slot_starts_at_one_like_tensorflow(cpu_device):
    import tensorflow as tf
    from gg import executor
    tf.reset_default_graph()
    w = tf.Variable(np.full((4, 3), 0.5, np.float32), name="w")
    x = tf.placeholder(tf.float32, shape=[2, 4])
    loss = tf.reduce_mean(tf.square(tf.matmul(x, w)))
    op = tf.train.RMSPropOptimizer(learning_rate=5e-5).minimize(loss, var_list=[w])
    executor.Plan(executor.RT, [op], [x])
    (ms, mom), = executor.RT.slots.values()
    assert torch.all(ms == 1.0) and torch.all(mom == 0.0)
    # TF form, first step from ms0 = 1: ms = 0.9 + 0.1 g^2 ; p -= lr g / sqrt(ms + 1e-10)  (~ lr*g, NOT 3.16 lr sign(g))
    g_ = 0.3
    step_tf = 5e-5 * g_ / np.sqrt(0.9 + 0.1 * g_ * g_ + 1e-10)
    step_zero_init = 5e-5 * g_ / np.sqrt(0.1 * g_ * g_ + 1e-10)
    assert abs(step_tf / (5e-5 * g_) - 1.0) < 0.06 and step_zero_init / step_tf > 3.0


def test_adam_slots_still_start_at_zero(cpu_device):
    from gg import executor
    g = _gmgan(8)
    executor.Plan(executor.RT, [g.disc_cost, g.disc_train_op], [g.real_x_int])
    for m, v in executor.RT.slots.values():
        assert not m.any() and not v.any()


def test_rng_key_differs_per_rank(monkeypatch):
    from gg import executor, dist
    rt = executor.Runtime()
    monkeypatch.setattr(dist, "rank", lambda: 0)
    s0 = rt.rng_seed()
    monkeypatch.setattr(dist, "rank", lambda: 3)
    s3 = rt.rng_seed()
    assert s0 == rt.seed                 # single-GPU streams unchanged
    assert s3 != s0 and (s3 >> 32) == 3 and (s3 & 0xFFFFFFFF) == (s0 & 0xFFFFFFFF)


def test_fan_uses_python2_integer_division():
    """reference conv2d.py:63 / deconv2d.py:52 run under Python 2: `output_dim*filter_size**2/(stride**2)` truncates"""
    import tensorflow as tf
    import tflib as lib
    import tflib.ops.conv2d
    import tflib.ops.deconv2d
    tf.reset_default_graph()
    lib.delete_all_params()
    x = tf.placeholder(tf.float32, shape=[2, 4, 8, 8])
    np.random.seed(7)
    lib.ops.conv2d.Conv2D('c', 4, 3, 5, x, stride=2)              # fan_out = 3*25 // 4 = 18 (true division: 18.75)
    w = lib._params['c.Filters'].attrs["init"]
    np.random.seed(7)
    stdev = np.sqrt(4. / (4 * 25 + (3 * 25) // 4))
    ref = np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=(5, 5, 4, 3)).astype('float32')
    assert np.array_equal(w, ref)
    y = tf.placeholder(tf.float32, shape=[2, 3, 4, 4])
    np.random.seed(9)
    lib.ops.deconv2d.Deconv2D('d', 3, 2, 5, y)                    # fan_in = 3*25 // 4 = 18
    w = lib._params['d.Filters'].attrs["init"]
    np.random.seed(9)
    stdev = np.sqrt(4. / ((3 * 25) // 4 + 2 * 25))
    ref = np.random.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=(5, 5, 2, 3)).astype('float32')
    assert np.array_equal(w, ref)
