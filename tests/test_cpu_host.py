"""CPU (no GPU): the C-ABI library loads and exports every symbol include/gg_b200.h declares; host-side graph logic
(shape inference, TF SAME geometry, build-time fusion / transpose sinking, symbolic gradients incl. second order,
plan pruning order); the tflib drop-in surface (names, registry sharing, initialiser replay)."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


def test_cabi_exports_every_declared_symbol():
    from gg import cabi
    hdr = open(os.path.join(ROOT, "include", "gg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(cabi.lib, name), "libgg_b200.so does not export %s" % name
        assert name in cabi.SIGNATURES, "cabi.py has no ctypes signature for %s" % name
    assert cabi.lib.gg_version() >= 100
    assert cabi.lib.gg_bn_slices(4096, 128) >= 1                 # pure host helpers are callable without a GPU
    assert cabi.lib.gg_conv2d_workspace(0, 64, 16, 16, 64, 128, 5, 2, 8, 8) >= 256


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tensorflow as tf
    from gg.cabi import GGError
    tf.reset_default_graph()
    x = tf.constant(np.ones((2, 2), np.float32))
    with pytest.raises(GGError):
        tf.Session().run(x * 2.0)


def test_same_geometry_and_conv_nodes():
    from gg import layers
    g = layers.conv_geometry(64, 32, 32, 3, 64, 5, 2, "SAME")
    assert (g["Ho"], g["Wo"], g["pad_t"], g["pad_l"]) == (16, 16, 1, 1)
    g = layers.conv_geometry(50, 7, 7, 128, 256, 5, 2, "SAME")
    assert (g["Ho"], g["pad_t"]) == (4, 2)
    g = layers.conv_geometry(2, 7, 7, 16, 8, 4, 1, "VALID")
    assert (g["Ho"], g["Wo"], g["pad_t"]) == (4, 4, 0)
    with pytest.raises(ValueError):
        layers.conv_geometry(1, 8, 8, 1, 1, 3, 1, "FULL")


def _fresh():
    import tensorflow as tf
    import tflib as lib
    tf.reset_default_graph()
    lib.delete_all_params()
    return tf, lib


def test_build_time_fusion_keeps_image_layers_in_nhwc():
    tf, lib = _fresh()
    import tflib.ops.conv2d, tflib.ops.batchnorm, tflib.ops.deconv2d, tflib.ops.linear
    x = tf.placeholder(tf.float32, shape=[8, 3 * 32 * 32])
    h = tf.reshape(x, [-1, 3, 32, 32])
    h = lib.ops.conv2d.Conv2D('D.1', 3, 64, 5, h, stride=2)
    h = tf.maximum(0.2 * h, h)                                       # LeakyReLU as the scripts write it
    h = lib.ops.conv2d.Conv2D('D.2', 64, 128, 5, h, stride=2)
    h = lib.ops.batchnorm.Batchnorm('D.BN2', [0, 2, 3], h)
    h = tf.nn.relu(h)
    assert h.shape == (8, 128, 8, 8)                                 # NCHW surface
    assert h.op == "transpose" and h.attrs["perm"] == (0, 3, 1, 2)   # ... as a view of an NHWC tensor
    bn = h.inputs[0]
    assert bn.op == "bn" and bn.attrs["act"] == "relu"               # activation folded into batch norm
    conv2 = bn.inputs[0]
    assert conv2.op == "conv" and conv2.attrs["mode"] == "fwd" and len(conv2.inputs) == 3   # bias fused
    conv1 = conv2.inputs[0]
    assert conv1.op == "conv" and conv1.attrs["act"] == "leaky" and abs(conv1.attrs["alpha"] - 0.2) < 1e-7
    flat = tf.reshape(h, [-1, 128 * 8 * 8])
    assert flat.inputs[0].op == "transpose"                          # flattening materialises NCHW order (C,H,W)


def test_registry_shares_parameters_and_replays_initialiser_draws():
    tf, lib = _fresh()
    import tflib.ops.linear
    np.random.seed(7)
    a = np.random.uniform(-1, 1, size=(3, 4))                        # what one Linear(3,4) call consumes
    np.random.seed(7)
    x = tf.placeholder(tf.float32, shape=[2, 3])
    y1 = lib.ops.linear.Linear('G.L', 3, 4, x)
    y2 = lib.ops.linear.Linear('G.L', 3, 4, x)                       # second call: same variables, but it DRAWS again
    after = np.random.uniform(-1, 1, size=(3, 4))
    np.random.seed(7)
    np.random.uniform(-1, 1, size=(3, 4)); np.random.uniform(-1, 1, size=(3, 4))
    assert np.array_equal(after, np.random.uniform(-1, 1, size=(3, 4)))
    assert y1.inputs[1] is y2.inputs[1] and len(lib.params_with_name('G.L')) == 2
    stdev = np.sqrt(2. / 7) * np.sqrt(3)
    assert np.allclose(lib._params['G.L.W'].attrs["init"], (a * stdev).astype('float32'))
    assert [p.name for p in lib.params_with_name('.b')] == ['G.L.b']
    with pytest.raises(Exception):
        lib.ops.linear.Linear('G.bad', 3, 4, x, initialization='nope')
    import tflib.ops.deconv2d
    with pytest.raises(Exception):
        lib.ops.deconv2d.Deconv2D('G.d', 4, 4, 5, tf.placeholder(tf.float32, shape=[1, 4, 4, 4]), mask_type=('a', 1))


def test_symbolic_gradients_first_and_second_order_shapes():
    tf, lib = _fresh()
    import tflib.ops.conv2d, tflib.ops.linear
    x = tf.placeholder(tf.float32, shape=[4, 3 * 8 * 8])
    h = tf.reshape(x, [-1, 3, 8, 8])
    h = lib.ops.conv2d.Conv2D('Discriminator.1', 3, 32, 5, h, stride=2)
    h = tf.maximum(0.2 * h, h)
    h = tf.reshape(h, [-1, 32 * 4 * 4])
    out = tf.reshape(lib.ops.linear.Linear('Discriminator.Out', 32 * 4 * 4, 1, h), [-1])
    params = lib.params_with_name('Discriminator')
    grads = tf.gradients(tf.reduce_mean(out), params)
    assert all(g is not None and tuple(g.shape) == tuple(p.shape) for g, p in zip(grads, params))
    # WGAN-GP: gradient w.r.t. the input, then differentiate its norm w.r.t. the weights (gan_inference_svhn.py:351-354)
    gx = tf.gradients(out, [x])[0]
    assert tuple(gx.shape) == (4, 192)
    slopes = tf.sqrt(tf.reduce_sum(tf.square(gx), reduction_indices=[1]))
    gp = 10. * tf.reduce_mean((slopes - 1.) ** 2)
    g2 = tf.gradients(gp, params)
    by = dict(zip([p.name for p in params], g2))
    assert by['Discriminator.1.Filters'] is not None and tuple(by['Discriminator.1.Filters'].shape) == (5, 5, 3, 32)
    assert by['Discriminator.Out.W'] is not None
    assert by['Discriminator.1.Biases'] is None and by['Discriminator.Out.b'] is None    # biases do not enter d(out)/dx


def test_gmgan_graph_builds_with_reference_parameter_inventory():
    tf, lib = _fresh()
    import gmgan_inference_cifar10 as S
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=64)
    names = set(lib._params)
    for n in ['Generator.Input.W', 'Generator.BN1.scale', 'Generator.2.Filters', 'Generator.5.Biases', 'Generator.Hyper.Mu',
              'Extractor.1.Filters', 'Extractor.BN3.offset', 'Extractor.Output.W', 'Discriminator.HyperInput.W',
              'Discriminator.3.Filters', 'Discriminator.zx1.W', 'Discriminator.Output.b', 'Generator.BN2.moving_mean']:
        assert n in names, n
    assert tuple(lib._params['Generator.2.Filters'].shape) == (5, 5, 128, 256)      # (k,k,Cout,Cin)  deconv2d.py:60-69
    assert tuple(lib._params['Discriminator.zx1.W'].shape) == (4608, 512)
    n_g = sum(p.size for p in g.gen_params + g.ext_params if 'moving' not in p.name)
    n_d = sum(p.size for p in g.disc_params)
    assert 3.0e6 < n_g < 3.3e6 and 3.9e6 < n_d < 4.2e6                              # 3.12 M / 4.06 M
    assert g.gen_cost.shape == () and g.fake_x.shape == (64, 3072)
    # the train op carries one gradient per trainable variable; BN moving_* get none (like TF: `None` grads are skipped)
    deps = dict(zip([v.name for v in g.gen_train_op.attrs["vars"]], g.gen_train_op.deps))
    assert deps['Generator.BN2.moving_mean'] is None and deps['Generator.Hyper.Mu'] is not None
    # local_ep prunes rec_x: the reconstruction path is not reachable from the generator step's fetches
    from gg.ops import toposort
    reach = {n.id for n in toposort([g.gen_cost] + [d for d in g.gen_train_op.deps if d is not None])}
    # (the [B,3072] reshapes collapse into the consumers' own reshapes: compare their NCHW sources)
    rec_src, fake_src = g.rec_x.inputs[0].inputs[0], g.fake_x.inputs[0].inputs[0]     # the tanh-fused Generator.5 kernels
    assert fake_src.op == 'conv' and fake_src.attrs['mode'] == 'dgrad' and fake_src.attrs['act'] == 'tanh'
    assert rec_src.id not in reach and fake_src.id in reach
    # ... and the discriminator reads the generator's NHWC output directly: the NCHW view of fake_x is never materialised
    assert g.fake_x.inputs[0].id not in reach


def test_deferred_value_waits_once_and_acts_like_numpy():
    """executor.Deferred (Session(deferred_fetches=True)): the host waits for the copy's event on first use only, and the
    value then behaves like the numpy result TensorFlow's session.run returns"""
    import torch
    from gg.executor import Deferred

    class Ev(object):
        def __init__(self): self.waits = 0
        def synchronize(self): self.waits += 1
        def query(self): return True
    ev = Ev()
    d = Deferred(torch.tensor([2.5]), ev, ())
    assert ev.waits == 0 and d.done()
    assert float(d) == 2.5 and d + 1 == 3.5 and 2 * d == 5.0 and d > 2 and "%.2f" % d == "2.50" and ev.waits == 1
    assert np.mean([d, d]) == 2.5 and np.asarray(d).shape == ()
    m = Deferred(torch.arange(6, dtype=torch.float32), Ev(), (2, 3))
    assert m.shape == (2, 3) and m[1][2] == 5.0 and np.array_equal(np.asarray(m), np.arange(6, dtype=np.float32).reshape(2, 3))
