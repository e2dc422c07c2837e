"""-m gpu: generator samples at iteration 100 against the reference trajectory (north_star: "generator samples within stated
tolerance of the reference at iteration 100").

The golden (tests/golden/trajectory100.npz, made by tests/golden/make_trajectory.py) holds the first 100 iterations of
gmgan_inference_cifar10.py (local_ep, bs 64, loop :480-494) from the tflib-initialised weights with injected per-step noise,
evaluated twice by the CPU oracle: in float64 (the arbiter) and in float32 (the reference's own precision).

What the measurement shows (table printed by the test, committed as profiles/trajectory100_r2.txt): GAN training is a
chaotic map.  Adam's first steps are ~lr*sign(g) and every ReLU / LeakyReLU mask flip is a discontinuity, so two
implementations of the SAME algorithm separate exponentially: the float32 CPU oracle itself leaves the float64 oracle at
1e-3 after 2 iterations, ~5e-2 after 20, and is fully decorrelated (rel-L2 ~ 1, samples saturating the tanh) after 50.
A point-wise tolerance at iteration 100 therefore cannot be met by ANY fp32 implementation — including TensorFlow run
twice with different thread counts.  The stated tolerances are:
  * iterations 1..20, point-wise: sample rel-L2 vs the fp64 oracle <= 3x the fp32 oracle's own drift at that checkpoint,
    plus the per-op rounding floor of the backend propagated through the generator (2e-5 fp32 kernels, 3e-3 tf32 kernels);
  * iterations 50 and 100, where point-wise comparison is meaningless: the costs stay finite and inside the oracle's
    envelope, and the sample STATISTICS (mean, standard deviation over all 300 x 3072 fixed-noise sample values) are
    within 0.15 of the fp64 oracle's — the fp32 oracle's own statistics differ from it by a comparable amount."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FLOOR = {1: 2e-5, 0: 3e-3}       # backend -> rounding floor of one generator pass (fp32 direct / tf32 tensor cores)


def _run(backend):
    import tensorflow as tf
    import tflib as lib
    import gmgan_inference_cifar10 as S
    from gg import cabi
    from oracle import gmgan_cifar10 as OM
    gold = np.load(os.path.join(HERE, "golden", "trajectory100.npz"))
    B, n_keep = int(gold["batch"]), int(gold["n_keep"])
    checkpoints = [int(c) for c in gold["checkpoints"]]
    tf.reset_default_graph()
    lib.delete_all_params()
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=B)
    cabi.call("gg_set_conv_backend", backend)
    rows = []
    try:
        sess = tf.Session()
        feeds = lambda inp: {g.real_x_int: inp["real_x_int"], g.hyper_p_z: inp["hyper_p_z"], g.hyper_p_k_idx: inp["k_idx"],
                             g.gumbel_uniforms[0]: inp["U"]}
        step = 0
        gen_costs, disc_costs = np.full(len(gold["disc_costs"]), np.nan), np.zeros(len(gold["disc_costs"]))
        for it in range(max(checkpoints)):
            if it > 0:
                gen_costs[it], _ = sess.run([g.gen_cost, g.gen_train_op], feed_dict=feeds(OM.synthetic_inputs(B, step))); step += 1
            disc_costs[it], _ = sess.run([g.disc_cost, g.disc_train_op], feed_dict=feeds(OM.synthetic_inputs(B, step))); step += 1
            if it + 1 in checkpoints:
                i = checkpoints.index(it + 1)
                full = sess.run(g.fixed_noise_samples)
                s, ref, ref32 = full[:n_keep], gold["samples"][i], gold["samples32"][i]
                l2 = float(np.linalg.norm(s - ref) / np.linalg.norm(ref))
                l2_32 = float(np.linalg.norm(ref32 - ref) / np.linalg.norm(ref))
                rows.append(dict(it=it + 1, l2=l2, oracle32_l2=l2_32, mean=float(full.mean()), std=float(full.std()),
                                 ref_mean=float(gold["stats"][i][0]), ref_std=float(gold["stats"][i][1]),
                                 o32_mean=float(gold["stats32"][i][0]), o32_std=float(gold["stats32"][i][1]),
                                 d_gen=0.0 if it == 0 else abs(gen_costs[it] - gold["gen_costs"][it]),
                                 d_disc=abs(disc_costs[it] - gold["disc_costs"][it])))
    finally:
        cabi.call("gg_set_conv_backend", 0)
    return rows, gen_costs, disc_costs, gold


@pytest.mark.parametrize("backend", [1, 0])
def test_generator_samples_through_iteration_100(backend):
    rows, gen_costs, disc_costs, gold = _run(backend)
    name = "fp32 direct" if backend == 1 else "tf32 tcgen05"
    print("\n%s kernels vs fp64 oracle (and the fp32 CPU oracle vs the same arbiter)" % name)
    print("   iter | sample rel-L2 | fp32-oracle rel-L2 | mean / std (ours) | mean / std (fp64 oracle) | mean / std (fp32 oracle) | |d gen| | |d disc|")
    for r in rows:
        print("   %4d |  %.3e   |     %.3e      | %+.4f / %.4f  |    %+.4f / %.4f     |    %+.4f / %.4f     | %.2e | %.2e" %
              (r["it"], r["l2"], r["oracle32_l2"], r["mean"], r["std"], r["ref_mean"], r["ref_std"], r["o32_mean"], r["o32_std"],
               r["d_gen"], r["d_disc"]))
    assert rows[-1]["it"] == 100
    for r in rows:
        if r["it"] <= 20:
            bound = 3.0 * r["oracle32_l2"] + FLOOR[backend] * (1 + r["it"])
            assert r["l2"] <= bound, "iteration %d: sample rel-L2 %.3e > stated bound %.3e" % (r["it"], r["l2"], bound)
        else:
            assert abs(r["mean"] - r["ref_mean"]) < 0.15 and abs(r["std"] - r["ref_std"]) < 0.15, r
    # the point of the second half of the docstring: the reference's own precision decorrelates too
    assert rows[-1]["oracle32_l2"] > 0.3
    assert np.isfinite(gen_costs[1:]).all() and np.isfinite(disc_costs).all()
    lo, hi = np.nanmin(gold["gen_costs"]), np.nanmax(gold["gen_costs"])
    assert np.nanmin(gen_costs) > lo - 0.5 * (hi - lo) and np.nanmax(gen_costs) < hi + 0.5 * (hi - lo)
    assert disc_costs.max() < gold["disc_costs"].max() * 2 + 1.0
