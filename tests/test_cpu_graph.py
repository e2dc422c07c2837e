"""CPU: the host-side graph logic (shape inference, build-time fusion, symbolic gradients of gg/ops.py, the rewrites of
gg/rewrite.py) evaluated by the float64 graph interpreter (tests/graph_interp.py) against the oracle's torch autograd.
No kernels run here; what is pinned is that the launch list the plan compiler emits COMPUTES the reference's step."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from graph_interp import Interp


def _build(batch=8, **kw):
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=batch, **kw)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    return g, lib, params


def _feeds(g, inp):
    return {g.real_x_int: inp["real_x_int"], g.hyper_p_z: inp["hyper_p_z"], g.hyper_p_k_idx: inp["k_idx"],
            g.gumbel_uniforms[0]: inp["U"]}


def _grads_of(train_op):
    return {v.name: d for v, d in zip(train_op.attrs["vars"], train_op.deps) if d is not None}


@pytest.mark.parametrize("batch", [8])
def test_gmgan_graph_gradients_match_oracle_autograd(batch):
    from oracle import gmgan_cifar10 as OM
    g, lib, params = _build(batch)
    model = OM.GMGANCifar10(params, dtype=torch.float64)
    inp = OM.synthetic_inputs(batch, 0)
    it = Interp(_feeds(g, inp))
    for cost, op, fn in ((g.gen_cost, g.gen_train_op, model.gen_step), (g.disc_cost, g.disc_train_op, model.disc_step)):
        ref_cost, ref_grads = fn(apply=False, **inp)
        assert abs(float(it.run(cost)) - ref_cost) < 1e-9 * max(1.0, abs(ref_cost))
        grads = _grads_of(op)
        assert set(grads) == set(k for k, v in ref_grads.items() if v is not None)
        for name, node in grads.items():
            got, ref = it.run(node), ref_grads[name].numpy()
            scale = np.abs(ref).max() + 1e-30
            # (a bias in front of batch norm has an exactly-zero gradient: absolute floor)
            assert np.abs(got - ref.reshape(got.shape)).max() <= 1e-8 * scale + 1e-13, name


# ---- sibling batching (gg/rewrite.py): the rewritten graph computes the same costs and parameter gradients ------------
def _all_nodes(roots):
    from gg.ops import toposort
    return toposort([r for r in roots if r is not None])


def _eval_family(build, batching, seed):
    """build the family's graph with sibling batching on/off; evaluate costs + every parameter gradient of both train ops
    with all placeholders and random draws fed from one seeded stream (in creation order, identical for both builds)"""
    import os
    import tensorflow as tf
    import tflib as lib
    os.environ["GG_BATCH_SIBLINGS"] = "1" if batching else "0"
    try:
        tf.reset_default_graph()
        lib.delete_all_params()
        np.random.seed(seed)
        g = build()
    finally:
        os.environ.pop("GG_BATCH_SIBLINGS", None)
    grads = {}
    for tag, op in (("gen", g.gen_train_op), ("disc", g.disc_train_op)):
        for v, d in zip(op.attrs["vars"], op.deps):
            if d is not None:
                grads[(tag, v.name)] = d
    roots = [g.gen_cost, g.disc_cost] + list(grads.values())
    nodes = _all_nodes(roots)
    rs = np.random.RandomState(99)
    feeds = {}
    for n in sorted(nodes, key=lambda n: n.id):
        if n.op == "placeholder" or (n.op == "random" and n.attrs["kind"] != "categorical"):
            if n.dtype.name == "int32":
                feeds[n] = rs.randint(0, 10 if n.size < 4096 else 256, size=tuple(n.shape))
            elif n.op == "placeholder" or n.attrs["kind"] == "uniform":
                feeds[n] = rs.uniform(0.05, 0.95, size=tuple(n.shape))
            else:
                feeds[n] = rs.randn(*n.shape)
        elif n.op == "random":
            feeds[n] = rs.randint(0, n.inputs[0].size, size=tuple(n.shape))
    it = Interp(feeds)
    n_heavy = sum(1 for n in nodes if n.op in ("conv", "matmul"))
    vals = {k: it.run(v) for k, v in grads.items()}
    # (gen_cost is a [B] vector in REINFORCE mode — scalar + per-sample score function, as in the reference; minimise sums it)
    return float(np.sum(it.run(g.gen_cost))), float(np.sum(it.run(g.disc_cost))), vals, n_heavy


def _families():
    import gmgan_inference_cifar10 as C
    import gan_inference_svhn as V
    import gan_inference_face as F
    import ssgan_inference_moving_mnist as M
    import gmgan_inference_mnist as N
    return {
        "gmgan_mnist_local_ep": lambda: N.build_graph(BATCH_SIZE=3),
        "gmgan_cifar10_reinforce": lambda: C.build_graph(BATCH_SIZE=3, MODE_K='REINFORCE'),
        "gan_mnist_ali_bn_in_critic": lambda: __import__("gan_inference_mnist").build_graph(MODE='ali', BATCH_SIZE=3),
        "gan_cifar10_wali_gp": lambda: __import__("gan_inference_cifar10").build_graph(MODE='wali-gp', BATCH_SIZE=2, DIM=16),
        "ssgan_chairs": lambda: __import__("ssgan_inference_chairs").build_graph(BATCH_SIZE=2, LEN=3, DIM=8),
        "gmgan_face_local_ep": lambda: __import__("gmgan_inference_face").build_graph(BATCH_SIZE=2, N_COMS=5, DIM_G=8, DIM_D=8),
        "gmgan_svhn_local_epce": lambda: __import__("gmgan_inference_svhn").build_graph(MODE='local_epce', BATCH_SIZE=2),
        "gmgan_cifar10_local_ep": lambda: C.build_graph(BATCH_SIZE=4),
        "gan_svhn_wali_gp": lambda: V.build_graph(MODE='wali-gp', BATCH_SIZE=4),
        "gan_face_ali": lambda: F.build_graph(BATCH_SIZE=2),
        "ssgan_moving_mnist": lambda: M.build_graph(BATCH_SIZE=2, LEN=3),
    }


@pytest.mark.parametrize("family", ["gmgan_cifar10_local_ep", "gmgan_mnist_local_ep", "gmgan_svhn_local_epce", "gmgan_face_local_ep",
                                    "gan_svhn_wali_gp", "gan_face_ali", "gan_mnist_ali_bn_in_critic", "gan_cifar10_wali_gp", "ssgan_chairs",
                                    "gmgan_cifar10_reinforce",
                                    "ssgan_moving_mnist"])
def test_sibling_batching_preserves_costs_and_gradients(family):
    build = _families()[family]
    g0, d0, v0, h0 = _eval_family(build, False, 7)
    g1, d1, v1, h1 = _eval_family(build, True, 7)
    assert h1 < h0, "batching should remove conv/dense launches (%d -> %d)" % (h0, h1)
    assert abs(g0 - g1) < 1e-9 * max(1, abs(g0)) and abs(d0 - d1) < 1e-9 * max(1, abs(d0))
    assert set(v0) == set(v1)
    for k in v0:
        scale = np.abs(v0[k]).max() + 1e-30
        assert np.abs(v0[k] - v1[k]).max() <= 1e-8 * scale + 1e-13, k


def test_gmgan_mnist_graph_matches_oracle_autograd():
    """configs[0] on the CPU: the compiled MNIST graph (crop between deconvolutions sunk into NHWC, 7 -> 4 conv with (2,2)
    padding, single-channel first / last layers) against oracle/gmgan_mnist.py, costs and all gradients, batch 5"""
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_mnist as S
    from oracle import gmgan_mnist as OM
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(5)
    g = S.build_graph(BATCH_SIZE=5)
    params = {n: p.attrs["init"] for n, p in lib._params.items()}
    model = OM.GMGANMnist(params, dtype=torch.float64)
    inp = OM.synthetic_inputs(5, 0)
    it = Interp({g.real_x: inp["real_x"], g.hyper_p_z: inp["hyper_p_z"], g.hyper_p_k_idx: inp["k_idx"],
                 g.gumbel_uniforms[0]: inp["U"]})
    for cost, op, fn in ((g.gen_cost, g.gen_train_op, model.gen_step), (g.disc_cost, g.disc_train_op, model.disc_step)):
        ref_cost, ref_grads = fn(apply=False, **inp)
        assert abs(float(it.run(cost)) - ref_cost) < 1e-9 * max(1.0, abs(ref_cost))
        for name, node in _grads_of(op).items():
            got, ref = it.run(node), ref_grads[name].numpy()
            assert np.abs(got - ref.reshape(got.shape)).max() <= 1e-8 * (np.abs(ref).max() + 1e-30) + 1e-13, name


def test_objs_kl_matches_closed_forms_and_autograd():
    """tflib/objs/kl.py:5-14 through the graph IR: values and gradients w.r.t. the posterior parameters vs torch"""
    import tensorflow as tf
    import tflib as lib
    import tflib.objs.kl
    from gg import ops as O
    tf.reset_default_graph()
    lib.delete_all_params()
    rs = np.random.RandomState(3)
    B, D = 6, 5
    arrs = {k: rs.randn(B, D) for k in ("qm", "pm", "x", "mu")}
    arrs.update({k: rs.uniform(0.5, 1.5, size=(B, D)) for k in ("qs", "ps", "std")})
    ph = {k: tf.placeholder(tf.float32, shape=[B, D]) for k in arrs}
    kl = lib.objs.kl.kl_q_p_diagonal_gaussian(ph["qm"], ph["qs"], ph["pm"], ph["ps"])
    nll = lib.objs.kl.neg_log_likelihood_diagnoal_gaussian(ph["x"], ph["mu"], ph["std"])
    g_qm, g_qs = O.gradients(kl, [ph["qm"], ph["qs"]])
    g_mu, = O.gradients(nll, [ph["mu"]])
    it = Interp({ph[k]: v for k, v in arrs.items()})
    t = {k: torch.tensor(v, requires_grad=True) for k, v in arrs.items()}
    ref_kl = (0.5 * (torch.log(t["ps"] ** 2 / t["qs"] ** 2) + ((t["pm"] - t["qm"]) ** 2 + t["qs"] ** 2) / t["ps"] ** 2 - 1)).sum(1).mean()
    ref_nll = (0.5 * (((t["x"] - t["mu"]) / t["std"]) ** 2 + np.log(2 * np.pi) + 2 * torch.log(t["std"]))).sum(1).mean()
    r_qm, r_qs = torch.autograd.grad(ref_kl, [t["qm"], t["qs"]])
    r_mu, = torch.autograd.grad(ref_nll, [t["mu"]])
    assert abs(float(it.run(kl)) - float(ref_kl.detach())) < 1e-12 and abs(float(it.run(nll)) - float(ref_nll.detach())) < 1e-12
    for got, ref in ((g_qm, r_qm), (g_qs, r_qs), (g_mu, r_mu)):
        assert np.abs(it.run(got) - ref.numpy()).max() < 1e-12


def test_objs_kl_aggregated_matches_numpy_estimators():
    """tflib/objs/kl_aggregated.py:17-72 through the graph IR with the Monte-Carlo draws injected: ikl and the two
    mixture log-likelihoods against direct NumPy evaluations (logsumexp form)"""
    import tensorflow as tf
    import tflib as lib
    import tflib.objs.kl_aggregated as KA
    from gg.ops import toposort
    tf.reset_default_graph()
    lib.delete_all_params()
    rs = np.random.RandomState(8)
    nx, nz, dz = 5, 7, 3
    mu, std = rs.randn(nx, dz), rs.uniform(0.5, 1.5, size=(nx, dz))
    pm, ps = np.zeros((nz, dz)), np.ones((nz, dz))
    z = rs.randn(nz, dz)
    t = {k: tf.placeholder(tf.float32, shape=list(v.shape)) for k, v in (("mu", mu), ("std", std), ("pm", pm), ("ps", ps), ("z", z))}
    lq = KA.log_likelihood_mixture_gaussian(t["z"], t["mu"], t["std"])
    lm = KA.log_likelihood_mixture_mixture_gaussian(t["z"], t["mu"], t["std"], t["pm"], t["ps"], nx)
    ikl = KA.ikl_q_aggregated_p_diagonal_gaussian(t["mu"], t["std"], t["pm"], t["ps"], nz, dz)
    feeds = {t["mu"]: mu, t["std"]: std, t["pm"]: pm, t["ps"]: ps, t["z"]: z}
    for n in toposort([ikl]):
        if n.op == "random":
            feeds[n] = z                                  # the estimator's own z ~ p draw
    it = Interp(feeds)

    def ll(x, m, s):
        return (-.5 * (((x - m) / s) ** 2 + np.log(2 * np.pi) + 2 * np.log(s))).sum(-1)
    mat = ll(z[:, None, :], mu[None], std[None])                                           # [nz, nx]
    ref_lq = np.log(np.exp(mat).mean(1))
    ref_lp = ll(z, pm, ps)
    ref_lm = np.log(np.concatenate([np.exp(mat), np.tile(np.exp(ref_lp)[:, None], (1, nx))], 1).mean(1))
    assert np.abs(it.run(lq) - ref_lq).max() < 1e-10 and np.abs(it.run(lm) - ref_lm).max() < 1e-10
    assert abs(float(it.run(ikl)) - float((ref_lp - ref_lq).mean())) < 1e-10


def test_objs_mmd_matches_numpy_and_autograd():
    """tflib/objs/mmd.py through the graph IR: biased and unbiased mixed-RBF MMD^2 and d/dX vs torch"""
    import tensorflow as tf
    import tflib as lib
    import tflib.objs.mmd as MM
    from gg import ops as O
    tf.reset_default_graph()
    lib.delete_all_params()
    rs = np.random.RandomState(4)
    X, Y = rs.randn(6, 3), rs.randn(5, 3) * 1.3
    tx, ty = tf.placeholder(tf.float32, shape=[6, 3]), tf.placeholder(tf.float32, shape=[5, 3])
    biased = MM.mix_rbf_mmd2(tx, ty)
    unbiased = MM.mix_rbf_mmd2(tx, ty, biased=False)
    gx, = O.gradients(biased, [tx])
    it = Interp({tx: X, ty: Y})
    a, b = torch.tensor(X, requires_grad=True), torch.tensor(Y)

    def kern(u, v):
        d2 = ((u[:, None, :] - v[None, :, :]) ** 2).sum(-1)
        return sum(torch.exp(-d2 / (2 * s ** 2)) for s in MM.SIGMAS)
    kxx, kxy, kyy = kern(a, a), kern(a, b), kern(b, b)
    ref_b = kxx.sum() / 36 + kyy.sum() / 25 - 2 * kxy.sum() / 30
    d = float(len(MM.SIGMAS))
    ref_u = (kxx.sum() - 6 * d) / 30 + (kyy.sum() - 5 * d) / 20 - 2 * kxy.sum() / 30
    rg, = torch.autograd.grad(ref_b, [a])
    assert abs(float(it.run(biased)) - float(ref_b.detach())) < 1e-10
    assert abs(float(it.run(unbiased)) - float(ref_u.detach())) < 1e-10
    assert np.abs(it.run(gx) - rg.numpy()).max() < 1e-10


def _tf_same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


@pytest.mark.parametrize("geom", [(2, 4, 8, 8, 3, 5, 4, 4, 2, 2), (2, 4, 8, 8, 4, 6, 4, 4, 2, 1), (1, 16, 8, 8, 2, 4, 4, 4, 2, 2),
                                  (2, 5, 6, 7, 3, 4, 3, 3, 1, 2), (1, 3, 5, 5, 2, 3, 2, 3, 2, 3)])
def test_conv3d_matches_torch_conv3d_with_tf_same_padding(geom):
    """tflib.ops.conv3d.Conv3D (tflib/ops/conv3d.py:6-51; NDHWC, filter (fl,k,k,Ci,Co), strides [1,sl,s,s,1], SAME) through the
    graph IR — depth taps folded into channels + one 2-D convolution — against torch.nn.functional.conv3d with TensorFlow's
    asymmetric SAME padding: the output and the gradients w.r.t. input, filter and bias."""
    import tensorflow as tf
    import tflib as lib
    import tflib.ops.conv3d
    from gg import ops as O
    N, L, H, W, Ci, Co, fl, k, s, sl = geom
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(5)
    x = tf.placeholder(tf.float32, shape=[N, L, H, W, Ci])
    y = lib.ops.conv3d.Conv3D('C3', fl, Ci, Co, k, x, stride=s, stride_len=sl)
    w, b = lib.params_with_name('C3.Filters')[0], lib.params_with_name('C3.Biases')[0]
    assert tuple(w.shape) == (fl, k, k, Ci, Co) and tuple(b.shape) == (1, 1, 1, 1, Co)
    rs = np.random.RandomState(2)
    xv, wv, bv = rs.randn(N, L, H, W, Ci), rs.randn(fl, k, k, Ci, Co), rs.randn(1, 1, 1, 1, Co)
    gy = rs.randn(*y.shape)
    loss = tf.reduce_sum(y * tf.constant(gy.astype(np.float32)))
    gx, gw, gb = O.gradients(loss, [x, w, b])
    it = Interp({x: xv}, params={w.id: wv, b.id: bv})
    tx, tw, tb = (torch.tensor(v, requires_grad=True) for v in (xv, wv, bv))
    pd, ph, pw = _tf_same_pad(L, fl, sl), _tf_same_pad(H, k, s), _tf_same_pad(W, k, s)
    xin = F.pad(tx.permute(0, 4, 1, 2, 3), (pw[0], pw[1], ph[0], ph[1], pd[0], pd[1]))
    ref = F.conv3d(xin, tw.permute(4, 3, 0, 1, 2), stride=(sl, s, s)).permute(0, 2, 3, 4, 1) + tb
    assert tuple(ref.shape) == tuple(y.shape)
    rx, rw, rb = torch.autograd.grad((ref * torch.tensor(gy.astype(np.float32).astype(np.float64))).sum(), [tx, tw, tb])
    for got, want, name in ((y, ref.detach(), "y"), (gx, rx, "dx"), (gw, rw, "dw"), (gb, rb, "db")):
        g, r = it.run(got), want.numpy()
        assert np.abs(g - r.reshape(g.shape)).max() <= 1e-10 * (np.abs(r).max() + 1e-30) + 1e-12, (name, geom)


def test_ssgan_3dcnn_critic_trunk_builds_and_reduces_to_4x4():
    """the four Conv3D layers + Batchnorm([0,1,2,3]) of the reference's 3dcnn critic (ssgan_inference_moving_mnist.py:356-389),
    LEN = 4 and LEN = 16: the clip collapses to one 4x4x(8 DIM) map, and the batch norm over every axis but the channel one is
    the fused [rows, C] kernel's graph op"""
    import tensorflow as tf
    import tflib as lib
    import tflib.ops.conv3d
    import tflib.ops.batchnorm
    DIM, B = 4, 2
    for LEN in (4, 16):
        tf.reset_default_graph()
        lib.delete_all_params()
        np.random.seed(1)
        x = tf.placeholder(tf.float32, shape=[B, LEN, 4096])
        out = tf.transpose(tf.reshape(x, [-1, LEN, 1, 64, 64]), [0, 1, 3, 4, 2])                   # NLHWC (:355-356)
        out = lib.ops.conv3d.Conv3D('D.1', 4, 1, DIM, 4, out, stride=2, stride_len=2)
        out = lib.ops.conv3d.Conv3D('D.2', 4, DIM, 2 * DIM, 4, out, stride=2, stride_len=1 if LEN == 4 else 2)
        out = lib.ops.batchnorm.Batchnorm('D.BN2', [0, 1, 2, 3], out)
        out = lib.ops.conv3d.Conv3D('D.3', 4, 2 * DIM, 4 * DIM, 4, out, stride=2, stride_len=2)
        out = lib.ops.conv3d.Conv3D('D.4', 4, 4 * DIM, 8 * DIM, 4, out, stride=2, stride_len=1 if LEN == 4 else 2)
        assert tuple(out.shape) == (B, 1, 4, 4, 8 * DIM)
        flat = tf.reshape(out, [B, 4 * 4 * 8 * DIM])
        rs = np.random.RandomState(0)
        val = Interp({x: rs.rand(B, LEN, 4096)}).run(flat)
        assert val.shape == (B, 4 * 4 * 8 * DIM) and np.isfinite(val).all() and np.abs(val).max() > 0
        bn = [n for n in O_toposort([flat]) if n.op == "bn"]
        assert len(bn) == 1 and len(bn[0].shape) == 2 and bn[0].shape[1] == 2 * DIM


def O_toposort(roots):
    from gg.ops import toposort
    return toposort(roots)
