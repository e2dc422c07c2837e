"""-m gpu: Session(deferred_fetches=True) — the training loop of gmgan_inference_cifar10.py:483-494 with the host running one
run ahead of the device (bench.py's end-to-end loop) — produces bit-identical costs and parameters to the loop that waits
for every run: the staging rings (pinned feed slots, pinned fetch slots) never hand a buffer back while a copy is in flight."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200", "scripts"))


def _train(deferred, iters=12, B=16):
    import tensorflow as tf
    import tflib as lib
    from gg.executor import RT, Deferred
    import gmgan_inference_cifar10 as S
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(77)
    g = S.build_graph(BATCH_SIZE=B)
    sess = tf.Session(deferred_fetches=deferred)
    from oracle import gmgan_cifar10 as OM
    costs = []
    step = 0
    for i in range(iters):
        for cost, op in ((g.gen_cost, g.gen_train_op), (g.disc_cost, g.disc_train_op)):
            inp = OM.synthetic_inputs(B, step)                                # every random input fed: the two loops see the
            step += 1                                                         # same batches AND the same noise
            feeds = {g.real_x_int: inp["real_x_int"].astype(np.uint8), g.hyper_p_z: inp["hyper_p_z"],
                     g.hyper_p_k_idx: inp["k_idx"], g.gumbel_uniforms[0]: inp["U"]}
            c, _ = sess.run([cost, op], feed_dict=feeds)
            if deferred:
                assert isinstance(c, Deferred)
            costs.append(c)                                                  # read only after the loop (> ring depth runs)
    vals = np.array([float(c) for c in costs])
    params = {n: RT.get_param(p).copy() for n, p in sorted(lib._params.items())}
    return vals, params


def test_deferred_fetches_match_synchronous_loop():
    v0, p0 = _train(False)
    v1, p1 = _train(True)
    assert np.isfinite(v0).all()
    assert np.array_equal(v0, v1), (v0, v1)
    for n in p0:
        assert np.array_equal(p0[n], p1[n]), n


def test_deferred_value_behaves_like_numpy():
    import tensorflow as tf
    import tflib as lib
    from gg.executor import Deferred
    tf.reset_default_graph()
    lib.delete_all_params()
    a = tf.placeholder(tf.float32, shape=[4, 3])
    s = tf.reduce_sum(a * 2.0)
    sess = tf.Session(deferred_fetches=True)
    x = np.arange(12, dtype=np.float32).reshape(4, 3)
    out, arr = sess.run([s, a * 2.0], feed_dict={a: x})
    assert isinstance(out, Deferred) and isinstance(arr, Deferred)
    assert float(out) == 132.0 and out + 1 == 133.0 and "%.1f" % out == "132.0" and out > 100
    assert arr.shape == (4, 3) and np.array_equal(np.asarray(arr), 2 * x) and np.mean([out, out]) == 132.0
