"""CPU: the face (ALI, 64x64x3) and SSGAN (moving MNIST) graphs evaluated by the float64 graph interpreter against the
new oracles' torch autograd (oracle/gan_face.py, oracle/ssgan_moving_mnist.py): costs and every parameter gradient the
plan compiler would launch.  Two independent restatements of the reference models (tf-surface script port vs functional
torch) must agree to rounding — this is what pins the shared-epsilon recurrence, the B*LEN folding and the weighted
objective on the host side; the -m gpu tests then hold the CUDA path to the same oracles."""
import numpy as np
import pytest
import torch

from graph_interp import Interp


def _grads_of(train_op):
    return {v.name: d for v, d in zip(train_op.attrs["vars"], train_op.deps) if d is not None}


def _check(g, it, model, inp, tol=1e-8):
    for cost, op, fn in ((g.gen_cost, g.gen_train_op, model.gen_step), (g.disc_cost, g.disc_train_op, model.disc_step)):
        ref_cost, ref_grads = fn(apply=False, **inp)
        assert abs(float(it.run(cost)) - ref_cost) < 1e-9 * max(1.0, abs(ref_cost)), (float(it.run(cost)), ref_cost)
        grads = _grads_of(op)
        assert set(grads) == set(k for k, v in ref_grads.items() if v is not None)
        for name, node in grads.items():
            got, ref = it.run(node), ref_grads[name].numpy()
            scale = np.abs(ref).max() + 1e-30
            assert np.abs(got - ref.reshape(got.shape)).max() <= tol * scale + 1e-13, name


def test_gan_face_graph_matches_oracle():
    import tensorflow as tf
    import tflib as lib
    import gan_inference_face as S
    from oracle import gan_face as OM
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(11)
    B = 3
    g = S.build_graph(BATCH_SIZE=B, DIM_G=8, DIM_D=8)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    model = OM.GANFace(params, dtype=torch.float64, dim_g=8, dim_d=8)
    inp = OM.synthetic_inputs(B, 0)
    it = Interp({g.real_x_int: inp["real_x_int"], g.dequant: inp["dequant"], g.p_z: inp["p_z"]})
    _check(g, it, model, inp)


@pytest.mark.parametrize("mode", ["local_ep", "local_epce-z"])
def test_ssgan_graph_matches_oracle(mode):
    import tensorflow as tf
    import tflib as lib
    import ssgan_inference_moving_mnist as S
    from oracle import ssgan_moving_mnist as OM
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(12)
    B, LEN = 2, 4
    g = S.build_graph(MODE=mode, BATCH_SIZE=B, LEN=LEN, DIM=8)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    model = OM.SSGANMovingMNIST(params, B, LEN, mode=mode, dtype=torch.float64, dim=8)
    assert np.allclose(model.ratio, g.ratio)
    inp = OM.synthetic_inputs(B, LEN, 0)
    assert len(g.epsilons) == 1                       # ONE epsilon node shared by the LEN-1 unrolled transitions (:137-139)
    feeds = {g.real_x_unit: inp["real_x_unit"], g.real_y: inp["real_y"], g.p_z_l_0: inp["p_z_l_0"], g.epsilons[0]: inp["epsilon"],
             g.p_z_g: inp["p_z_g"], g.p_y_idx: inp["p_y_idx"]}
    it = Interp(feeds)
    _check(g, it, model, inp)


@pytest.mark.parametrize("mode", ["local_ep", "local_epce", "ali", "alice", "vegan"])
def test_gmgan_cifar10_objective_modes_match_oracle(mode):
    """the MODE matrix of gmgan_inference_cifar10.py:355-410 (objectives gan_inference.py:47-223): script port vs oracle"""
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    from oracle import gmgan_cifar10 as OM
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(5)
    B = 4
    g = S.build_graph(MODE=mode, BATCH_SIZE=B, DIM=8, N_COMS=6)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    model = OM.GMGANCifar10(params, dtype=torch.float64, dim=8, n_coms=6, mode=mode)
    inp = OM.synthetic_inputs(B, 0, n_coms=6, dim_latent=g.DIM_LATENT)
    it = Interp({g.real_x_int: inp["real_x_int"], g.hyper_p_z: inp["hyper_p_z"], g.hyper_p_k_idx: inp["k_idx"],
                 g.gumbel_uniforms[0]: inp["U"]})
    _check(g, it, model, inp)


@pytest.mark.parametrize("spatial", [False, True])
def test_batchnorm_second_order_gradient_matches_torch_double_backward(spatial):
    """WGAN-GP through a critic WITH batch norm (gan_inference_mnist.py:225-230,346-357): d/dtheta of
    mean((||dD/dx||_2 - 1)^2) needs the gradient OF the batch-norm gradient.  gg/ops.py::_grad_bn_grad vs torch autograd
    (create_graph=True) on a two-layer critic, dense [B, C] and spatial NHWC statistics."""
    import tensorflow as tf
    import tflib as lib
    import tflib.ops.linear
    import tflib.ops.conv2d
    import tflib.ops.batchnorm
    from oracle import tf_ops as O
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(3)
    rs = np.random.RandomState(4)
    B = 6
    if spatial:
        xin = tf.placeholder(tf.float32, shape=[B, 2 * 8 * 8])
        h = lib.ops.conv2d.Conv2D('Discriminator.1', 2, 4, 5, tf.reshape(xin, [B, 2, 8, 8]), stride=2)
        h = lib.ops.batchnorm.Batchnorm('Discriminator.BN1', [0, 2, 3], h)
        h = tf.maximum(0.2 * h, h)
        h = tf.reshape(h, [B, 4 * 4 * 4])
        out = lib.ops.linear.Linear('Discriminator.Out', 64, 1, h)
    else:
        xin = tf.placeholder(tf.float32, shape=[B, 10])
        h = lib.ops.linear.Linear('Discriminator.1', 10, 7, xin)
        h = lib.ops.batchnorm.Batchnorm('Discriminator.BN1', [0], h)
        h = tf.maximum(0.2 * h, h)
        out = lib.ops.linear.Linear('Discriminator.Out', 7, 1, h)
    grad = tf.gradients(tf.reshape(out, [-1]), [xin])[0]
    slopes = tf.sqrt(tf.reduce_sum(tf.square(grad), reduction_indices=[1]))
    pen = tf.reduce_mean((slopes - 1.) ** 2)
    plist = [p for p in lib.params_with_name('Discriminator') if 'moving_' not in p.name]
    # make the BN affine non-trivial
    for p in plist:
        if 'scale' in p.name or 'offset' in p.name:
            p.attrs["init"] = (p.attrs["init"] + rs.uniform(0.3, 0.9, size=p.attrs["init"].shape)).astype(np.float32)
    gs = tf.gradients(pen, plist)
    x = rs.randn(*xin.shape)
    it = Interp({xin: x})
    P = {p.name: torch.tensor(p.attrs["init"], dtype=torch.float64, requires_grad=True) for p in plist}
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    if spatial:
        th = O.conv2d(tx.reshape(B, 2, 8, 8), P['Discriminator.1.Filters'], 2, 'SAME', P['Discriminator.1.Biases'])
        th = O.leaky_relu(O.batchnorm(th, P['Discriminator.BN1.scale'], P['Discriminator.BN1.offset'], [0, 2, 3])).reshape(B, 64)
    else:
        th = O.linear(tx, P['Discriminator.1.W'], P['Discriminator.1.b'])
        th = O.leaky_relu(O.batchnorm(th, P['Discriminator.BN1.scale'], P['Discriminator.BN1.offset'], [0]))
    tout = O.linear(th, P['Discriminator.Out.W'], P['Discriminator.Out.b']).reshape(-1)
    tg, = torch.autograd.grad(tout.sum(), tx, create_graph=True)
    tpen = ((torch.sqrt((tg ** 2).sum(1)) - 1.) ** 2).mean()
    refs = torch.autograd.grad(tpen, [P[p.name] for p in plist], allow_unused=True)
    assert abs(float(it.run(pen)) - float(tpen.detach())) < 1e-10
    checked = 0
    for p, gnode, r in zip(plist, gs, refs):
        if r is None or gnode is None:
            assert (r is None or float(r.abs().max()) < 1e-12) and (gnode is None or np.abs(it.run(gnode)).max() < 1e-10), p.name
            continue
        got, ref = it.run(gnode), r.numpy()
        assert np.abs(got - ref.reshape(got.shape)).max() <= 1e-8 * (np.abs(ref).max() + 1e-30) + 1e-12, p.name
        checked += 1
    assert checked >= 4
