"""tflib.objs.kl_aggregated — drop-in for tflib/objs/kl_aggregated.py:6-103: Monte-Carlo estimates of KL / inverse KL / JSD
between the AGGREGATED posterior q(z) = mean_i N(mu_i, std_i) (a mixture with one component per data point of the batch)
and a diagonal-Gaussian prior, plus the three VEGAN objectives built on them (MODE 'vegan-kl' / 'vegan-ikl' / 'vegan-jsd' of
the gan_inference_* scripts; no discriminator — SURVEY.md §8(f) N2).  Same names, arguments and return values; every
function is script-level tf.* glue, i.e. element-wise / reduction / small dense launches of libgg_b200."""
import math

import numpy as np
import tensorflow as tf

LOG_2PI = math.log(2 * math.pi)


def mixture_gaussian(n_samples, n_coms, dim_z, mu, std):
    """n_samples draws from the uniform mixture of the n_coms Gaussians (mu[k], std[k])"""
    pi = tf.constant(np.ones(n_coms).astype('float32') / n_coms)
    k = tf.cast(tf.one_hot(indices=tf.distributions.Categorical(probs=pi).sample(n_samples), depth=n_coms), tf.float32)
    eps = tf.random_normal([n_samples, dim_z])
    return tf.add(tf.matmul(k, mu), tf.multiply(tf.matmul(k, std), eps))


def log_likelihood_diagnoal_gaussian(x, mu, std):           # (sic)
    return tf.reduce_sum(-.5 * (tf.pow((x - mu) / std, 2) + LOG_2PI + 2 * tf.log(std)), axis=-1)


def _log_mean_exp(mat):
    """log(mean(exp(mat), axis=1)) with the maximum factored out"""
    top = tf.reduce_max(mat, axis=1)
    return tf.log(tf.reduce_mean(tf.exp(mat - tf.expand_dims(top, axis=1)), axis=1)) + top


def log_likelihood_mixture_gaussian(x, mu, std):
    """log q(x) under the uniform mixture: x [nz, dz] against mu / std [nx, dz] -> [nz]"""
    return _log_mean_exp(log_likelihood_diagnoal_gaussian(tf.expand_dims(x, axis=1), tf.expand_dims(mu, axis=0),
                                                          tf.expand_dims(std, axis=0)))


def log_likelihood_mixture_mixture_gaussian(x, mu_q, std_q, mu_p, std_p, n_coms):
    """log m(x) for m = (q + p) / 2: the n_coms mixture components and n_coms copies of the prior term under one mean"""
    against_q = log_likelihood_diagnoal_gaussian(tf.expand_dims(x, axis=1), tf.expand_dims(mu_q, axis=0),
                                                 tf.expand_dims(std_q, axis=0))                      # [nz, nx]
    against_p = tf.tile(tf.expand_dims(log_likelihood_diagnoal_gaussian(x, mu_p, std_p), axis=1), [1, n_coms])
    return _log_mean_exp(tf.concat([against_q, against_p], axis=1))


def kl_q_aggregated_p_diagonal_gaussian(q_z_mean, q_z_std, p_z_mean, p_z_std, n_samples, n_coms, dim_z):
    z = mixture_gaussian(n_samples, n_coms, dim_z, q_z_mean, q_z_std)                                 # z ~ q
    return tf.reduce_mean(log_likelihood_mixture_gaussian(z, q_z_mean, q_z_std) -
                          log_likelihood_diagnoal_gaussian(z, p_z_mean, p_z_std), axis=0)


def ikl_q_aggregated_p_diagonal_gaussian(q_z_mean, q_z_std, p_z_mean, p_z_std, n_samples, dim_z):
    z = tf.random_normal([n_samples, dim_z])                                                          # z ~ p
    return tf.reduce_mean(log_likelihood_diagnoal_gaussian(z, p_z_mean, p_z_std) -
                          log_likelihood_mixture_gaussian(z, q_z_mean, q_z_std), axis=0)


def jsd_q_aggregated_p_diagonal_gaussian(q_z_mean, q_z_std, p_z_mean, p_z_std, n_samples, n_coms, dim_z):
    z_q = mixture_gaussian(n_samples, n_coms, dim_z, q_z_mean, q_z_std)
    log_q = log_likelihood_mixture_gaussian(z_q, q_z_mean, q_z_std)
    log_m_q = log_likelihood_mixture_mixture_gaussian(z_q, q_z_mean, q_z_std, p_z_mean, p_z_std, n_coms)
    z_p = tf.random_normal([n_samples, dim_z])
    log_p = log_likelihood_diagnoal_gaussian(z_p, p_z_mean, p_z_std)
    log_m_p = log_likelihood_mixture_mixture_gaussian(z_p, q_z_mean, q_z_std, p_z_mean, p_z_std, n_coms)
    return tf.reduce_mean(.5 * (log_q - log_m_q + log_p - log_m_p), axis=0)


def _minimise(gen_cost, gen_params, lr, beta1):
    return tf.train.AdamOptimizer(learning_rate=lr, beta1=beta1).minimize(gen_cost, var_list=gen_params)


def vegan_jsd(q_z_mean, q_z_std, p_z_mean, p_z_std, rec_penalty, gen_params, z_samples, batchsize, dim_z, lamb, lr=2e-4, beta1=.5):
    gen_cost = lamb * jsd_q_aggregated_p_diagonal_gaussian(q_z_mean, q_z_std, p_z_mean, p_z_std, z_samples, batchsize, dim_z)
    gen_cost += rec_penalty
    return gen_cost, _minimise(gen_cost, gen_params, lr, beta1)


def vegan_kl(q_z_mean, q_z_std, p_z_mean, p_z_std, rec_penalty, gen_params, z_samples, batchsize, dim_z, lamb, lr=2e-4, beta1=.5):
    gen_cost = lamb * kl_q_aggregated_p_diagonal_gaussian(q_z_mean, q_z_std, p_z_mean, p_z_std, z_samples, batchsize, dim_z)
    gen_cost += rec_penalty
    return gen_cost, _minimise(gen_cost, gen_params, lr, beta1)


def vegan_ikl(q_z_mean, q_z_std, p_z_mean, p_z_std, rec_penalty, gen_params, z_samples, dim_z, lamb, lr=2e-4, beta1=.5):
    gen_cost = lamb * ikl_q_aggregated_p_diagonal_gaussian(q_z_mean, q_z_std, p_z_mean, p_z_std, z_samples, dim_z)
    gen_cost += rec_penalty
    return gen_cost, _minimise(gen_cost, gen_params, lr, beta1)
