"""CPU: the host-side graph logic (shape inference, build-time fusion, symbolic gradients of gg/ops.py, the rewrites of
gg/rewrite.py) evaluated by the float64 graph interpreter (tests/graph_interp.py) against the oracle's torch autograd.
No kernels run here; what is pinned is that the launch list the plan compiler emits COMPUTES the reference's step."""
import numpy as np
import pytest
import torch

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
