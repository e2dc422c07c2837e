"""tflib.objs.gan_inference — drop-in for tflib/objs/gan_inference.py (same function names, signatures, return
tuples).  Each function builds the adversarial objective from discriminator logits and returns symbolic costs plus
runnable train ops; running a train op executes forward, backward and ONE fused multi-tensor Adam launch
(TensorFlow's ApplyAdam arithmetic, epsilon outside the bias correction).

Reference lines: wali :4-26, wali_gp :28-45, ali :47-79, local_ep :81-119, local_epce :121-159, alice :161-192,
vegan :194-223, vegan_wgan_gp :225-244, local_ep_dynamic :246-304, weighted_local_epce :307-358.
"""
import tensorflow as tf

import tflib as lib
from gg.rewrite import batch_pairs as _siblings   # D(fake) and D(real) share weights: one batched application (gg/rewrite.py)


def _bce(logits, label):
    """mean sigmoid cross-entropy against a constant 0/1 label (tf.ones_like / tf.zeros_like in the reference)"""
    labels = tf.ones_like(logits) if label else tf.zeros_like(logits)
    return tf.reduce_mean(tf.nn.sigmoid_cross_entropy_with_logits(logits=logits, labels=labels))


def _gen_term(disc_fake, disc_real):
    return _bce(disc_fake, 1) + _bce(disc_real, 0)


def _disc_term(disc_fake, disc_real):
    return _bce(disc_fake, 0) + _bce(disc_real, 1)


def _adam_ops(gen_cost, disc_cost, gen_params, disc_params, **kw):
    gen_train_op = tf.train.AdamOptimizer(**kw).minimize(gen_cost, var_list=gen_params)
    disc_train_op = tf.train.AdamOptimizer(**kw).minimize(disc_cost, var_list=disc_params)
    return gen_train_op, disc_train_op


def wali(disc_fake, disc_real, gen_params, disc_params, lr=5e-5):
    disc_fake, disc_real = _siblings(disc_fake, disc_real)
    gen_cost = -tf.reduce_mean(disc_fake) - tf.reduce_mean(disc_real)     # sic: the reference negates both (:5)
    disc_cost = tf.reduce_mean(disc_fake) - tf.reduce_mean(disc_real)
    gen_train_op = tf.train.RMSPropOptimizer(learning_rate=lr).minimize(gen_cost, var_list=gen_params)
    disc_train_op = tf.train.RMSPropOptimizer(learning_rate=lr).minimize(disc_cost, var_list=disc_params)
    clip_ops = [tf.assign(var, tf.clip_by_value(var, -.01, .01)) for var in lib.params_with_name('Discriminator')]
    clip_disc_weights = tf.group(*clip_ops)
    return gen_cost, disc_cost, clip_disc_weights, gen_train_op, disc_train_op, clip_ops


def wali_gp(disc_fake, disc_real, gradient_penalty, gen_params, disc_params, lr=1e-4):
    disc_fake, disc_real = _siblings(disc_fake, disc_real)
    gen_cost = -tf.reduce_mean(disc_fake) + tf.reduce_mean(disc_real)
    disc_cost = tf.reduce_mean(disc_fake) - tf.reduce_mean(disc_real)
    disc_cost += gradient_penalty
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=0.5, beta2=0.9)
    return (gen_cost, disc_cost) + ops


def ali(disc_fake, disc_real, gen_params, disc_params, lr=2e-4, beta1=0.5, beta2=0.999, s_f=None):
    disc_fake, disc_real = _siblings(disc_fake, disc_real)
    gen_cost = _gen_term(disc_fake, disc_real)
    disc_cost = _disc_term(disc_fake, disc_real)
    if s_f is not None:
        gen_cost += s_f
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1, beta2=beta2)
    return (gen_cost, disc_cost) + ops


def _local_sums(disc_fake_list, disc_real_list, ratio_list=None):
    gen_cost, disc_cost, gen_terms, disc_terms = 0, 0, [], []
    for i, (disc_fake, disc_real) in enumerate(zip(disc_fake_list, disc_real_list)):
        g, d = _gen_term(disc_fake, disc_real), _disc_term(disc_fake, disc_real)
        if ratio_list is not None:
            g, d = float(ratio_list[i]) * g, float(ratio_list[i]) * d
        gen_terms.append(g)
        disc_terms.append(d)
        gen_cost, disc_cost = gen_cost + g, disc_cost + d
    return gen_cost, disc_cost, gen_terms, disc_terms


def local_ep(disc_fake_list, disc_real_list, gen_params, disc_params, lr=2e-4, beta1=0.5, beta2=.999, s_f=None):
    """the north-star objective: one (fake, real) logit pair per local discriminator, costs averaged over the list"""
    disc_fake_list, disc_real_list = _siblings(disc_fake_list, disc_real_list)
    gen_cost, disc_cost, _, _ = _local_sums(disc_fake_list, disc_real_list)
    if s_f is not None:
        gen_cost += s_f
    gen_cost /= len(disc_fake_list)
    disc_cost /= len(disc_fake_list)
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1, beta2=beta2)
    return (gen_cost, disc_cost) + ops


def local_epce(disc_fake_list, disc_real_list, rec_penalty, gen_params, disc_params, lr=2e-4, beta1=0.5, s_f=None):
    disc_fake_list, disc_real_list = _siblings(disc_fake_list, disc_real_list)
    gen_cost, disc_cost, _, _ = _local_sums(disc_fake_list, disc_real_list)
    if s_f is not None:
        gen_cost += s_f
    gen_cost /= len(disc_fake_list)
    disc_cost /= len(disc_fake_list)
    gen_cost += rec_penalty
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1)
    return (gen_cost, disc_cost) + ops


def alice(disc_fake, disc_real, rec_penalty, gen_params, disc_params, lr=2e-4, beta1=0.5, s_f=None):
    disc_fake, disc_real = _siblings(disc_fake, disc_real)
    gen_cost = _gen_term(disc_fake, disc_real)
    if s_f is not None:
        gen_cost += s_f
    gen_cost += rec_penalty
    disc_cost = _disc_term(disc_fake, disc_real)
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1)
    return (gen_cost, disc_cost) + ops


def vegan(disc_fake, disc_real, rec_penalty, gen_params, disc_params, lamb, lr=2e-4, beta1=.5, s_f=None):
    disc_fake, disc_real = _siblings(disc_fake, disc_real)
    gen_cost = _bce(disc_fake, 1)
    if s_f is not None:
        gen_cost += s_f
    gen_cost *= lamb
    gen_cost += rec_penalty
    disc_cost = _disc_term(disc_fake, disc_real)
    disc_cost *= (lamb / 2)
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1)
    return (gen_cost, disc_cost) + ops


def vegan_wgan_gp(disc_fake, disc_real, rec_penalty, gradient_penalty, gen_params, disc_params, lamb, lr=2e-4, beta1=.5):
    disc_fake, disc_real = _siblings(disc_fake, disc_real)
    gen_cost = -tf.reduce_mean(disc_fake) + tf.reduce_mean(disc_real)
    gen_cost *= lamb
    gen_cost += rec_penalty
    disc_cost = tf.reduce_mean(disc_fake) - tf.reduce_mean(disc_real)
    disc_cost *= lamb
    disc_cost += gradient_penalty
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1)
    return (gen_cost, disc_cost) + ops


def local_ep_dynamic(disc_fake_zz, disc_real_zz, disc_fake_xz, disc_real_xz, gen_params, disc_params, lr=2e-4, beta1=0.5,
                     beta2=.999, rec_penalty=None):
    disc_fake_zz, disc_real_zz = _siblings(disc_fake_zz, disc_real_zz)
    disc_fake_xz, disc_real_xz = _siblings(disc_fake_xz, disc_real_xz)
    gen_cost, disc_cost, _, _ = _local_sums(disc_fake_zz, disc_real_zz)
    if len(disc_fake_zz) > 0:
        gen_cost /= (len(disc_fake_zz) + 1)
        disc_cost /= (len(disc_fake_zz) + 1)
    gen_cost += _gen_term(disc_fake_xz, disc_real_xz)
    disc_cost += _disc_term(disc_fake_xz, disc_real_xz)
    if rec_penalty is not None:
        gen_cost += rec_penalty
    ops = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1, beta2=beta2)
    return (gen_cost, disc_cost) + ops


def weighted_local_epce(disc_fake_list, disc_real_list, ratio_list, gen_params, disc_params, lr=2e-4, beta1=0.5,
                        rec_penalty=None):
    """SSGAN objective (ssgan_inference_moving_mnist.py:547): entries weighted by ratio_list, no division"""
    disc_fake_list, disc_real_list = _siblings(disc_fake_list, disc_real_list)
    assert len(disc_fake_list) == ratio_list.shape[0]
    gen_cost, disc_cost, gen_debug_list, disc_debug_list = _local_sums(disc_fake_list, disc_real_list, ratio_list)
    if rec_penalty is not None:
        gen_cost += rec_penalty
    gen_train_op, disc_train_op = _adam_ops(gen_cost, disc_cost, gen_params, disc_params, learning_rate=lr, beta1=beta1)
    return gen_cost, disc_cost, gen_debug_list, disc_debug_list, gen_train_op, disc_train_op
