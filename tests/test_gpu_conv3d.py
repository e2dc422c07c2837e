"""-m gpu: tflib.ops.conv3d.Conv3D (tflib/ops/conv3d.py:6-51) on the CUDA kernels — output and the gradients w.r.t. input,
filter and bias against torch.nn.functional.conv3d (fp64) with TensorFlow's asymmetric SAME padding, at the layer shapes of the
reference's 3dcnn critic (ssgan_inference_moving_mnist.py:363-385, DIM = 32) and at 1e-3 of the tensor scale (north_star)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


def _same(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


@pytest.mark.parametrize("geom", [
    # N, L, H, W, Ci, Co, fl, k, s, sl
    (4, 4, 64, 64, 1, 32, 4, 4, 2, 2),      # Discriminator.1 (:363)
    (4, 2, 32, 32, 32, 64, 4, 4, 2, 1),     # Discriminator.2, LEN 4 (:368)
    (2, 8, 32, 32, 32, 64, 4, 4, 2, 2),     # Discriminator.2, LEN 16 (:370)
    (4, 2, 16, 16, 64, 128, 4, 4, 2, 2),    # Discriminator.3 (:376)
    (4, 1, 8, 8, 128, 256, 4, 4, 2, 1),     # Discriminator.4, LEN 4 (:383)
    (2, 3, 6, 7, 3, 8, 3, 3, 1, 2),         # ragged
])
def test_conv3d_matches_torch(geom):
    import tensorflow as tf
    import tflib as lib
    import tflib.ops.conv3d
    from gg.executor import RT
    N, L, H, W, Ci, Co, fl, k, s, sl = geom
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(3)
    x = tf.placeholder(tf.float32, shape=[N, L, H, W, Ci])
    y = lib.ops.conv3d.Conv3D('C3', fl, Ci, Co, k, x, stride=s, stride_len=sl)
    w, b = lib.params_with_name('C3.Filters')[0], lib.params_with_name('C3.Biases')[0]
    rs = np.random.RandomState(4)
    xv = rs.randn(N, L, H, W, Ci).astype(np.float32)
    gy = rs.randn(*y.shape).astype(np.float32)
    loss = tf.reduce_sum(y * tf.constant(gy))
    gx, gw, gb = tf.gradients(loss, [x, w, b])
    sess = tf.Session()
    wv, bv = RT.get_param(w).copy(), RT.get_param(b).copy()
    out = sess.run([y, gx, gw, gb], feed_dict={x: xv})
    tx, tw, tb = (torch.tensor(v.astype(np.float64), requires_grad=True) for v in (xv, wv, bv))
    pd, ph, pw = _same(L, fl, sl), _same(H, k, s), _same(W, k, s)
    xin = F.pad(tx.permute(0, 4, 1, 2, 3), (pw[0], pw[1], ph[0], ph[1], pd[0], pd[1]))
    ref = F.conv3d(xin, tw.permute(4, 3, 0, 1, 2), stride=(sl, s, s)).permute(0, 2, 3, 4, 1) + tb
    rx, rw, rb = torch.autograd.grad((ref * torch.tensor(gy.astype(np.float64))).sum(), [tx, tw, tb])
    for got, want, name in zip(out, (ref.detach(), rx, rw, rb), ("y", "dx", "dw", "db")):
        want = want.numpy().reshape(got.shape)
        err = np.abs(got - want).max() / (np.abs(want).max() + 1e-30)
        assert err < 1e-3, "Conv3D %s %s: %.3e of scale" % (geom, name, err)
