"""tflib.objs.mmd — drop-in for tflib/objs/mmd.py:4-79: the (mixed-RBF) maximum mean discrepancy between posterior and prior
codes and the VEGAN-MMD objective built on it (MODE 'vegan-mmd' of the gan_inference_* scripts; no discriminator —
SURVEY.md §8(f) N2).  Same names, arguments, defaults and return values; Gram matrices are [B, B] dense launches, the rest
is element-wise / reduction glue."""
import tensorflow as tf

SIGMAS = [2., 5., 10., 20., 40., 80.]


def maximum_mean_discripancy(sample, data, batch_size, sigma=SIGMAS):          # (sic) unused by the scripts (:70)
    x = tf.concat([sample, data], axis=0)
    x2 = tf.reduce_sum(tf.multiply(x, x), axis=1, keep_dims=True)
    exponent = tf.add(tf.add(tf.matmul(x, x, transpose_b=True), tf.scalar_mul(-.5, x2)), tf.scalar_mul(-.5, tf.transpose(x2)))
    s_all = tf.concat([tf.scalar_mul(1. / batch_size, tf.ones([tf.shape(sample)[0], 1])),
                       tf.scalar_mul(-1. / batch_size, tf.ones([tf.shape(data)[0], 1]))], axis=0)
    s_mat = tf.matmul(s_all, s_all, transpose_b=True)
    mmd_loss = 0.
    for s in sigma:
        mmd_loss += tf.reduce_sum(tf.multiply(s_mat, tf.exp(tf.scalar_mul(1. / s, exponent))))
    return tf.sqrt(mmd_loss)


def _mix_rbf_kernel(X, Y, sigmas, wts=None):
    if wts is None:
        wts = [1] * len(sigmas)
    XX, XY, YY = (tf.matmul(a, b, transpose_b=True) for a, b in ((X, X), (X, Y), (Y, Y)))
    x_sq, y_sq = tf.diag_part(XX), tf.diag_part(YY)
    row = lambda v: tf.expand_dims(v, 0)
    col = lambda v: tf.expand_dims(v, 1)
    K_XX, K_XY, K_YY = 0, 0, 0
    for sigma, wt in zip(sigmas, wts):
        gamma = 1 / (2 * sigma ** 2)
        K_XX += wt * tf.exp(-gamma * (-2 * XX + col(x_sq) + row(x_sq)))
        K_XY += wt * tf.exp(-gamma * (-2 * XY + col(x_sq) + row(y_sq)))
        K_YY += wt * tf.exp(-gamma * (-2 * YY + col(y_sq) + row(y_sq)))
    return K_XX, K_XY, K_YY, float(sum(wts))


def _mmd2(K_XX, K_XY, K_YY, const_diagonal=False, biased=False):
    m, n = float(K_XX.get_shape()[0]), float(K_YY.get_shape()[0])
    if biased:
        return tf.reduce_sum(K_XX) / (m * m) + tf.reduce_sum(K_YY) / (n * n) - 2 * tf.reduce_sum(K_XY) / (m * n)
    if const_diagonal is not False:
        trace_X, trace_Y = m * const_diagonal, n * const_diagonal
    else:
        trace_X, trace_Y = tf.trace(K_XX), tf.trace(K_YY)
    return ((tf.reduce_sum(K_XX) - trace_X) / (m * (m - 1)) + (tf.reduce_sum(K_YY) - trace_Y) / (n * (n - 1))
            - 2 * tf.reduce_sum(K_XY) / (m * n))


def mix_rbf_mmd2(X, Y, sigmas=SIGMAS, wts=None, biased=True):
    K_XX, K_XY, K_YY, d = _mix_rbf_kernel(X, Y, sigmas, wts)
    return _mmd2(K_XX, K_XY, K_YY, const_diagonal=d, biased=biased)


def vegan_mmd(q_z, p_z, rec_penalty, gen_params, batch_size, lamb, lr=2e-4, beta1=.5):
    gen_cost = lamb * mix_rbf_mmd2(q_z, p_z)
    gen_cost += rec_penalty
    gen_train_op = tf.train.AdamOptimizer(learning_rate=lr, beta1=beta1).minimize(gen_cost, var_list=gen_params)
    return gen_cost, gen_train_op
