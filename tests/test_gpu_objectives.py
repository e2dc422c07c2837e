"""-m gpu: costs and EVERY parameter gradient of the remaining model families / objectives through the tflib + tf surface and
the C-ABI, against the CPU oracles:

  * gan_inference_face.py (ALI, 64x64x3)                       vs oracle/gan_face.py
  * ssgan_inference_moving_mnist.py (LEN 8; shared-epsilon recurrence, B*LEN folding, weighted_local_epce)
                                                               vs oracle/ssgan_moving_mnist.py
  * gmgan_inference_cifar10.py MODE local_epce / alice / ali / vegan (gan_inference.py:47-223)
                                                               vs oracle/gmgan_cifar10.py(mode=...)
  * local_ep_dynamic (gan_inference.py:246-304) on a small critic stack  vs tf_ops.local_ep_dynamic_costs + autograd
  * one step of each remaining script port (gan_inference_mnist with batch norm inside the critic, gan_inference_cifar10,
    ssgan_inference_chairs, gmgan_inference_face) vs the float64 graph interpreter (tests/graph_interp.py) fed the same
    injected noise — the port-vs-reference side of those graphs is pinned on the CPU by test_cpu_graph.py.

Tolerances are the whole-backward-pass bounds of test_gpu_gmgan_step.py (5e-3 rel-L2 with the fp32 kernels, 8e-2 with tf32
tensor-core operands; per-op 1e-3 is held by test_gpu_kernels.py / test_gpu_production_shapes.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {1: (5e-3, 2e-2), 0: (8e-2, 3e-1)}     # backend -> (rel-L2, max-rel-to-scale)


def _reset(seed):
    import tensorflow as tf
    import tflib as lib
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(seed)
    return tf, lib


def _params(lib):
    from gg.executor import RT
    return {name: RT.get_param(p).copy() for name, p in lib._params.items()}


def _compare_step(tf, sess, g, feeds, oracle, inp, backend, names_of):
    tol_l2, tol_max = TOL[backend]
    worst = {}
    for which, cost, plist, fn in (("disc", g.disc_cost, names_of["disc"], oracle.disc_step),
                                   ("gen", g.gen_cost, names_of["gen"], oracle.gen_step)):
        plist = [p for p in plist if 'moving_' not in p.name]
        grads = tf.gradients(cost, plist)
        keep = [(p, gr) for p, gr in zip(plist, grads) if gr is not None]
        out = sess.run([cost] + [gr for _, gr in keep], feed_dict=feeds)
        ref_cost, ref_grads = fn(apply=False, **inp)
        assert abs(float(out[0]) - ref_cost) <= 2e-3 * max(abs(ref_cost), 1e-2), (which, float(out[0]), ref_cost)
        assert len(keep) == len([k for k, v in ref_grads.items() if v is not None]), which
        gmax = max(float(v.abs().max()) for v in ref_grads.values() if v is not None)
        for (p, _), got in zip(keep, out[1:]):
            ref = ref_grads[p.name].numpy()
            assert got.shape == ref.shape, p.name
            if np.abs(ref).max() < 1e-7 * gmax:
                assert np.abs(got).max() < 1e-4 * gmax, (p.name, np.abs(got).max(), gmax)
                continue
            l2 = float(np.linalg.norm(got.astype(np.float64) - ref) / (np.linalg.norm(ref) + 1e-30))
            mx = float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30))
            worst[(which, p.name)] = (mx, l2)
    for key, (mx, l2) in worst.items():
        assert l2 < tol_l2 and mx < tol_max, "%s: max-rel %.3e rel-L2 %.3e (backend %d)" % (key, mx, l2, backend)
    top = sorted(worst.items(), key=lambda kv: -kv[1][1])[:4]
    print("backend", backend, "worst gradient errors (max-rel, rel-L2):", top)


@pytest.mark.parametrize("backend", [1, 0])
def test_gan_face_step_gradients_match_oracle(backend):
    """BASELINE.json configs[3] at the per-GPU shard of the 8-GPU run (128 / 8 = 16 images of 64x64x3)"""
    from gg import cabi
    from oracle import gan_face as OM
    import gan_inference_face as S
    tf, lib = _reset(31)
    B = 16
    g = S.build_graph(BATCH_SIZE=B)
    cabi.call("gg_set_conv_backend", backend)
    try:
        sess = tf.Session()
        oracle = OM.GANFace(_params(lib), dtype=torch.float64)
        inp = OM.synthetic_inputs(B, 0)
        feeds = {g.real_x_int: inp["real_x_int"], g.dequant: inp["dequant"], g.p_z: inp["p_z"]}
        _compare_step(tf, sess, g, feeds, oracle, inp, backend, {"disc": g.disc_params, "gen": g.gen_params + g.ext_params})
    finally:
        cabi.call("gg_set_conv_backend", 0)


@pytest.mark.parametrize("mode,backend", [("local_ep", 1), ("local_ep", 0), ("local_epce-z", 0)])
def test_ssgan_moving_mnist_step_gradients_match_oracle(mode, backend):
    """BASELINE.json configs[4] geometry: LEN 8, 1x64x64 frames, per-GPU shard of the 4-GPU run (32 / 4 = 8 sequences ->
    64 frames folded into the batch); one shared epsilon through the 7 unrolled transitions"""
    from gg import cabi
    from oracle import ssgan_moving_mnist as OM
    import ssgan_inference_moving_mnist as S
    tf, lib = _reset(32)
    B, LEN = 8, 8
    g = S.build_graph(MODE=mode, BATCH_SIZE=B, LEN=LEN)
    cabi.call("gg_set_conv_backend", backend)
    try:
        sess = tf.Session()
        oracle = OM.SSGANMovingMNIST(_params(lib), B, LEN, mode=mode, dtype=torch.float64)
        inp = OM.synthetic_inputs(B, LEN, 0)
        assert len(g.epsilons) == 1
        feeds = {g.real_x_unit: inp["real_x_unit"], g.real_y: inp["real_y"], g.p_z_l_0: inp["p_z_l_0"],
                 g.epsilons[0]: inp["epsilon"], g.p_z_g: inp["p_z_g"], g.p_y_idx: inp["p_y_idx"]}
        _compare_step(tf, sess, g, feeds, oracle, inp, backend, {"disc": g.disc_params, "gen": g.gen_params + g.ext_params})
    finally:
        cabi.call("gg_set_conv_backend", 0)


@pytest.mark.parametrize("mode", ["local_epce", "alice", "ali", "vegan"])
def test_gmgan_cifar10_objective_modes_match_oracle(mode):
    """the other rows of the MODE matrix (gmgan_inference_cifar10.py:355-410) at bs=64 on the tensor-core path"""
    from oracle import gmgan_cifar10 as OM
    import gmgan_inference_cifar10 as S
    tf, lib = _reset(33)
    B = 64
    g = S.build_graph(MODE=mode, BATCH_SIZE=B)
    sess = tf.Session()
    oracle = OM.GMGANCifar10(_params(lib), dtype=torch.float64, mode=mode)
    inp = OM.synthetic_inputs(B, 0, dim_latent=g.DIM_LATENT)
    feeds = {g.real_x_int: inp["real_x_int"], g.hyper_p_z: inp["hyper_p_z"], g.hyper_p_k_idx: inp["k_idx"],
             g.gumbel_uniforms[0]: inp["U"]}
    _compare_step(tf, sess, g, feeds, oracle, inp, 0, {"disc": g.disc_params, "gen": g.gen_params + g.ext_params})


def test_local_ep_dynamic_matches_oracle():
    """tflib/objs/gan_inference.py:246-304 (imported by no script): two latent-pair critics + one (x, z) critic"""
    import tflib.ops.linear
    import tflib.objs.gan_inference
    import tflib.utils.distance
    from oracle import tf_ops as O
    tf, lib = _reset(34)
    B, DZ, DX = 32, 16, 64
    rs = np.random.RandomState(0)

    def LeakyReLU(x):
        return tf.maximum(0.2 * x, x)

    def critic(name, x, n_in):
        h = LeakyReLU(lib.ops.linear.Linear('Discriminator.%s.1' % name, n_in, 128, x))
        return tf.reshape(lib.ops.linear.Linear('Discriminator.%s.2' % name, 128, 1, h), [-1])

    z0 = tf.placeholder(tf.float32, shape=[B, DZ])
    xr = tf.placeholder(tf.float32, shape=[B, DX])
    z1 = lib.ops.linear.Linear('Generator.T', DZ, DZ, z0)                # "dynamics"
    xf = tf.tanh(lib.ops.linear.Linear('Generator.X', DZ, DX, z1))       # fake x
    q0 = lib.ops.linear.Linear('Extractor.Z0', DX, DZ, xr)
    q1 = lib.ops.linear.Linear('Extractor.Z1', DX, DZ, xr)
    fake_zz = [critic('zz', tf.concat([z0, z1], 1), 2 * DZ), critic('zz2', tf.concat([z1, z0], 1), 2 * DZ)]
    real_zz = [critic('zz', tf.concat([q0, q1], 1), 2 * DZ), critic('zz2', tf.concat([q1, q0], 1), 2 * DZ)]
    fake_xz, real_xz = critic('xz', tf.concat([xf, z1], 1), DX + DZ), critic('xz', tf.concat([xr, q1], 1), DX + DZ)
    rec = lib.utils.distance.distance(xr, tf.tanh(lib.ops.linear.Linear('Generator.X', DZ, DX, q1)), 'l2')
    gen_params = lib.params_with_name('Generator') + lib.params_with_name('Extractor')
    disc_params = lib.params_with_name('Discriminator')
    gen_cost, disc_cost, gop, dop = lib.objs.gan_inference.local_ep_dynamic(fake_zz, real_zz, fake_xz, real_xz, gen_params,
                                                                             disc_params, rec_penalty=rec)
    sess = tf.Session()
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in _params(lib).items()}
    a_z0, a_xr = rs.randn(B, DZ).astype(np.float32), rs.randn(B, DX).astype(np.float32)
    t_z0, t_xr = torch.tensor(a_z0, dtype=torch.float64), torch.tensor(a_xr, dtype=torch.float64)
    lin = lambda n, x: O.linear(x, P[n + '.W'], P[n + '.b'])
    tcritic = lambda n, x: lin('Discriminator.%s.2' % n, O.leaky_relu(lin('Discriminator.%s.1' % n, x))).reshape(-1)
    t_z1 = lin('Generator.T', t_z0)
    t_xf = torch.tanh(lin('Generator.X', t_z1))
    t_q0, t_q1 = lin('Extractor.Z0', t_xr), lin('Extractor.Z1', t_xr)
    t_rec = O.distance(t_xr, torch.tanh(lin('Generator.X', t_q1)), 'l2')
    rg, rd = O.local_ep_dynamic_costs([tcritic('zz', torch.cat([t_z0, t_z1], 1)), tcritic('zz2', torch.cat([t_z1, t_z0], 1))],
                                      [tcritic('zz', torch.cat([t_q0, t_q1], 1)), tcritic('zz2', torch.cat([t_q1, t_q0], 1))],
                                      tcritic('xz', torch.cat([t_xf, t_z1], 1)), tcritic('xz', torch.cat([t_xr, t_q1], 1)), t_rec)
    feeds = {z0: a_z0, xr: a_xr}
    for cost, plist, ref in ((gen_cost, gen_params, rg), (disc_cost, disc_params, rd)):
        grads = tf.gradients(cost, plist)
        out = sess.run([cost] + grads, feed_dict=feeds)
        assert abs(float(out[0]) - float(ref.detach())) <= 1e-3 * abs(float(ref.detach()))
        refs = torch.autograd.grad(ref, [P[p.name] for p in plist], retain_graph=True)
        for p, got, r in zip(plist, out[1:], refs):
            r = r.numpy()
            assert np.abs(got - r).max() <= 2e-3 * (np.abs(r).max() + 1e-30), p.name


def _interp_feeds(nodes, rs):
    feeds = {}
    for n in sorted(nodes, key=lambda n: n.id):
        if n.op == "placeholder" or (n.op == "random" and n.attrs["kind"] != "categorical"):
            if n.dtype.name == "int32":
                feeds[n] = rs.randint(0, 10 if n.size < 4096 else 256, size=tuple(n.shape)).astype(np.int32)
            elif n.op == "placeholder" or n.attrs["kind"] == "uniform":
                feeds[n] = rs.uniform(0.05, 0.95, size=tuple(n.shape)).astype(np.float32)
            else:
                feeds[n] = rs.randn(*n.shape).astype(np.float32)
        elif n.op == "random":
            feeds[n] = rs.randint(0, n.inputs[0].size, size=tuple(n.shape)).astype(np.int32)
    return feeds


FAMILIES = {
    # the four ports VERDICT r1 lists as having no -m gpu test, at reference-like sizes
    "gan_inference_mnist_ali_bn_in_critic": ("gan_inference_mnist", dict(MODE='ali', BATCH_SIZE=50)),
    "gan_inference_cifar10_ali": ("gan_inference_cifar10", dict(MODE='ali', BATCH_SIZE=32)),
    "ssgan_inference_chairs": ("ssgan_inference_chairs", dict(BATCH_SIZE=4, LEN=4)),
    "gmgan_inference_face": ("gmgan_inference_face", dict(BATCH_SIZE=16)),
    "gan_inference_mnist_wali_gp_bn_double_backward": ("gan_inference_mnist", dict(MODE='wali-gp', BATCH_SIZE=20)),
    # the discriminator-free VEGAN objectives (objs.kl_aggregated / objs.mmd behind MODE 'vegan-kl|ikl|jsd|mmd'): one step only
    "gan_inference_mnist_vegan_kl": ("gan_inference_mnist", dict(MODE='vegan-kl', BATCH_SIZE=50)),
    "gan_inference_mnist_vegan_jsd": ("gan_inference_mnist", dict(MODE='vegan-jsd', BATCH_SIZE=50)),
    "gan_inference_svhn_vegan_ikl": ("gan_inference_svhn", dict(MODE='vegan-ikl', BATCH_SIZE=32)),
    "gan_inference_cifar10_vegan_mmd": ("gan_inference_cifar10", dict(MODE='vegan-mmd', BATCH_SIZE=32)),
    # the SSGAN ALI critics: tflib.ops.conv3d.Conv3D (depth taps folded into one 2-D convolution) and the concat_z variant
    "ssgan_moving_mnist_ali_3dcnn": ("ssgan_inference_moving_mnist", dict(MODE='ali', ALI_MODE='3dcnn', BATCH_SIZE=4, LEN=4)),
    "ssgan_moving_mnist_alice_z_concat_z": ("ssgan_inference_moving_mnist", dict(MODE='alice-z', ALI_MODE='concat_z', BATCH_SIZE=4, LEN=4)),
}


@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_script_port_step_matches_fp64_interpreter(family):
    """costs + every parameter gradient of both train ops on the GPU (fp32 direct kernels: the comparison is then limited
    by fp32 rounding, 5e-3 rel-L2) against the float64 evaluation of the same graph"""
    import importlib
    from gg import cabi
    from gg.ops import toposort
    from graph_interp import Interp
    script, kw = FAMILIES[family]
    tf, lib = _reset(41)
    g = importlib.import_module(script).build_graph(**kw)
    grads = {}
    for tag, op in (("gen", g.gen_train_op), ("disc", g.disc_train_op)):
        if op is None:
            continue
        for v, d in zip(op.attrs["vars"], op.deps):
            if d is not None:
                grads[(tag, v.name)] = d
    cost_nodes = [c for c in (g.gen_cost, g.disc_cost) if c is not None]
    roots = cost_nodes + list(grads.values())
    feeds = _interp_feeds(toposort(roots), np.random.RandomState(99))
    it = Interp({k: np.asarray(v, np.float64) if v.dtype != np.int32 else v for k, v in feeds.items()})
    cabi.call("gg_set_conv_backend", 1)
    try:
        sess = tf.Session()
        keys = list(grads)
        out = sess.run(cost_nodes + [grads[k] for k in keys], feed_dict=feeds)
    finally:
        cabi.call("gg_set_conv_backend", 0)
    nc = len(cost_nodes)
    for got, node in zip(out[:nc], cost_nodes):
        ref = float(np.sum(it.run(node)))
        assert abs(float(np.sum(got)) - ref) <= 2e-3 * max(abs(ref), 1e-2), (family, float(np.sum(got)), ref)
    refs = {k: it.run(grads[k]) for k in keys}
    gmax = max(np.abs(r).max() for r in refs.values())
    worst = (0.0, None)
    for k, got in zip(keys, out[nc:]):
        ref = refs[k].reshape(got.shape)
        if np.abs(ref).max() < 1e-7 * gmax:
            assert np.abs(got).max() < 1e-4 * gmax, k
            continue
        l2 = float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30))
        worst = max(worst, (l2, k))
        # 5e-3 for the plain stacks; the mixture models put a softmax at temperature 0.1 (x10 on every logit error) and
        # batch norm between the loss and the first extractor layer: fp32 rounding reaches 2e-2 there (64x64 inputs)
        tol = 3e-2 if family.startswith("gmgan") else 5e-3
        assert l2 < tol, "%s %s: rel-L2 %.3e" % (family, k, l2)
    print(family, "worst rel-L2", worst)
