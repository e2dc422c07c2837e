"""SSGAN on the 3D-chairs sequences — Python-3 port of the reference's ssgan_inference_chairs.py (MODE 'local_ep' /
'local_epce-z', POS_MODE 'naive_mean_field', OP_DYN_MODE 'res_w' — the defaults, :28-32) on the B200 kernels.

The reference script is ssgan_inference_moving_mnist.py without the class labels, on 3x64x64 frames of LEN = 31, with the
transition operator's residual path learned (`output + Linear(name+'.ZW')(z_l)`, :108-109) instead of the identity, and
input decode 2*((x/256)-.5) (:503).  State-space latent, B*LEN frame folding, LEN-1 pair discriminators + z_g
discriminator + frame discriminator into weighted_local_epce exactly as in the moving-MNIST port.  Line numbers refer to
/root/reference/ssgan_inference_chairs.py.  The ALI / 3dcnn discriminators (:340-498) are non-default branches and are not
ported (Conv3D is off the hot path).
"""
import os
import sys
import time
import types

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import tensorflow as tf
import tflib as lib
import tflib.ops.linear
import tflib.ops.conv2d
import tflib.ops.batchnorm
import tflib.ops.deconv2d
import tflib.objs.gan_inference
import tflib.utils.distance
import tflib.plot


def build_graph(MODE='local_ep', BATCH_SIZE=50, LEN=31, DIM=32, DIM_OP=256, LR=1e-4, BN_FLAG=False, OP_DYN_MODE='res_w'):
    if MODE not in ('local_ep', 'local_epce-z'):
        raise NotImplementedError("MODE %r: only the graphical (local) modes are on the hot path" % MODE)
    DIM_LATENT_G, DIM_LATENT_L = 128, 8
    DIM_LATENT_T = DIM_LATENT_L
    OUTPUT_SHAPE = [3, 64, 64]
    OUTPUT_DIM = int(np.prod(OUTPUT_SHAPE))
    LAMBDA, BETA1 = 0.1, .5
    BN_FLAG_G = BN_FLAG_E = BN_FLAG_D = BN_FLAG
    ratio = [1.0, ] * (LEN - 1) + [1, LEN]
    ratio = np.asarray(ratio) * 1.0 / (len(ratio) + LEN - 1)                                 # :78-79
    ns = types.SimpleNamespace(MODE=MODE, BATCH_SIZE=BATCH_SIZE, LEN=LEN, ratio=ratio, epsilons=[], OUTPUT_DIM=OUTPUT_DIM)

    def LeakyReLU(x, alpha=0.2):
        return tf.maximum(alpha * x, x)

    def ImplicitOperator(z_l, epsilon, name):                                                # :94-114
        output = tf.concat([z_l, epsilon], axis=1)
        output = lib.ops.linear.Linear(name + '.Input', DIM_LATENT_L + DIM_LATENT_T, DIM_OP, output)
        output = LeakyReLU(output)
        output = lib.ops.linear.Linear(name + '.1', DIM_OP, DIM_OP, output)
        output = LeakyReLU(output)
        output = lib.ops.linear.Linear(name + '.Output', DIM_OP, DIM_LATENT_L, output)
        if OP_DYN_MODE == 'res':
            return output + z_l
        if OP_DYN_MODE == 'res_w':
            return output + lib.ops.linear.Linear(name + '.ZW', DIM_LATENT_L, DIM_LATENT_L, z_l)
        raise NotImplementedError(OP_DYN_MODE)

    def DynamicGenerator(z_l_0):                                                             # :134-141
        z_list = [z_l_0, ]
        epsilon = tf.random_normal([BATCH_SIZE, DIM_LATENT_T])
        ns.epsilons.append(epsilon)
        for i in range(LEN - 1):
            z_list.append(ImplicitOperator(z_list[-1], epsilon, 'Generator.Dynamic'))
        return tf.reshape(tf.concat(z_list, axis=1), [BATCH_SIZE, LEN, DIM_LATENT_L])

    def DynamicExtractor(z_l_pre):                                                           # :143-168, naive_mean_field
        return z_l_pre

    def Generator(z_g, z_l):                                                                 # :172-203
        z_g = tf.reshape(z_g, [BATCH_SIZE, DIM_LATENT_G])
        z_g = tf.tile(tf.expand_dims(z_g, axis=1), [1, LEN, 1])
        z_l = tf.reshape(z_l, [BATCH_SIZE, LEN, DIM_LATENT_L])
        z = tf.concat([z_g, z_l], axis=-1)
        z = tf.reshape(z, [BATCH_SIZE * LEN, DIM_LATENT_G + DIM_LATENT_L])
        output = lib.ops.linear.Linear('Generator.Input', DIM_LATENT_G + DIM_LATENT_L, 4 * 4 * 8 * DIM, z)
        if BN_FLAG_G:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN1', [0], output)
        output = tf.nn.relu(output)
        output = tf.reshape(output, [BATCH_SIZE * LEN, 8 * DIM, 4, 4])
        for i, (cin, cout) in enumerate(((8, 4), (4, 2), (2, 1))):
            output = lib.ops.deconv2d.Deconv2D('Generator.%d' % (i + 2), cin * DIM, cout * DIM, 5, output)
            if BN_FLAG_G:
                output = lib.ops.batchnorm.Batchnorm('Generator.BN%d' % (i + 2), [0, 2, 3], output)
            output = tf.nn.relu(output)
        output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM, 3, 5, output)
        output = tf.tanh(output)
        return tf.reshape(output, [BATCH_SIZE, LEN, OUTPUT_DIM])

    def _conv_trunk(prefix, output, first_in, bn):
        output = lib.ops.conv2d.Conv2D(prefix + '1', first_in, DIM, 5, output, stride=2)
        output = LeakyReLU(output)
        for i, (cin, cout) in enumerate(((1, 2), (2, 4), (4, 8))):
            output = lib.ops.conv2d.Conv2D('%s%d' % (prefix, i + 2), cin * DIM, cout * DIM, 5, output, stride=2)
            if bn:
                output = lib.ops.batchnorm.Batchnorm('%sBN%d' % (prefix, i + 2), [0, 2, 3], output)
            output = LeakyReLU(output)
        return output

    def Extractor(inputs):                                                                   # :205-230
        output = tf.reshape(inputs, [BATCH_SIZE * LEN, ] + OUTPUT_SHAPE)
        output = _conv_trunk('Extractor.', output, 3, BN_FLAG_E)
        output = tf.reshape(output, [BATCH_SIZE * LEN, 4 * 4 * 8 * DIM])
        output = lib.ops.linear.Linear('Extractor.Output', 4 * 4 * 8 * DIM, DIM_LATENT_L, output)
        return tf.reshape(output, [BATCH_SIZE, LEN, DIM_LATENT_L])

    def G_Extractor(inputs):                                                                 # :232-256 (Cin = 3*LEN)
        output = tf.reshape(inputs, [BATCH_SIZE, 3 * LEN, 64, 64])
        output = _conv_trunk('Extractor.G.', output, 3 * LEN, BN_FLAG_E)
        output = tf.reshape(output, [BATCH_SIZE, 4 * 4 * 8 * DIM])
        output = lib.ops.linear.Linear('Extractor.G.Output', 4 * 4 * 8 * DIM, DIM_LATENT_G, output)
        return tf.reshape(output, [BATCH_SIZE, DIM_LATENT_G])

    def Discriminator(x, z_g, z_l):                                                          # :259-301
        output = tf.reshape(x, [BATCH_SIZE * LEN, ] + OUTPUT_SHAPE)
        z_g = tf.tile(tf.expand_dims(tf.reshape(z_g, [BATCH_SIZE, DIM_LATENT_G]), axis=1), [1, LEN, 1])
        z_l = tf.reshape(z_l, [BATCH_SIZE, LEN, DIM_LATENT_L])
        z = tf.reshape(tf.concat([z_g, z_l], axis=-1), [BATCH_SIZE * LEN, DIM_LATENT_G + DIM_LATENT_L])
        output = _conv_trunk('Discriminator.', output, 3, BN_FLAG_D)            # dropout(training=False) is the identity
        output = tf.reshape(output, [BATCH_SIZE * LEN, 4 * 4 * 8 * DIM])
        z_output = LeakyReLU(lib.ops.linear.Linear('Discriminator.z1', DIM_LATENT_G + DIM_LATENT_L, 512, z))
        output = tf.concat([output, z_output], 1)
        output = LeakyReLU(lib.ops.linear.Linear('Discriminator.zx1', 4 * 4 * 8 * DIM + 512, 512, output))
        output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
        return tf.reshape(output, [BATCH_SIZE * LEN, ])

    def _mlp_disc(prefix, x, n_in):
        output = LeakyReLU(lib.ops.linear.Linear(prefix + '.Input', n_in, 512, x))
        output = LeakyReLU(lib.ops.linear.Linear(prefix + '.2', 512, 512, output))
        output = LeakyReLU(lib.ops.linear.Linear(prefix + '.3', 512, 512, output))
        return tf.reshape(lib.ops.linear.Linear(prefix + '.Output', 512, 1, output), [BATCH_SIZE, ])

    def DynamicDiscrminator(z1, z2):                                                         # :313-331
        z1 = tf.reshape(z1, [BATCH_SIZE, DIM_LATENT_L])
        z2 = tf.reshape(z2, [BATCH_SIZE, DIM_LATENT_L])
        return _mlp_disc('Discriminator.Dynamic', tf.concat([z1, z2], axis=1), DIM_LATENT_L * 2)

    def ZGDiscrminator(z_g):                                                                 # :333-349
        return _mlp_disc('Discriminator.ZG', tf.reshape(z_g, [BATCH_SIZE, DIM_LATENT_G]), DIM_LATENT_G)

    # ---- graph (:500-532) ----
    real_x_unit = tf.placeholder(tf.float32, shape=[BATCH_SIZE, LEN, OUTPUT_DIM])
    real_x = 2 * ((tf.cast(real_x_unit, tf.float32) / 256.) - .5)                            # :503 (frames fed as 0..255)
    q_z_l_pre = Extractor(real_x)
    q_z_g = G_Extractor(real_x)
    q_z_l = DynamicExtractor(q_z_l_pre)
    rec_x = Generator(q_z_g, q_z_l)
    p_z_l_0 = tf.random_normal([BATCH_SIZE, DIM_LATENT_L])
    p_z_l = DynamicGenerator(p_z_l_0)
    p_z_g = tf.random_normal([BATCH_SIZE, DIM_LATENT_G])
    fake_x = Generator(p_z_g, p_z_l)

    disc_fake, disc_real = [], []
    for i in range(LEN - 1):
        disc_fake.append(DynamicDiscrminator(p_z_l[:, i, :], p_z_l[:, i + 1, :]))
        disc_real.append(DynamicDiscrminator(q_z_l[:, i, :], q_z_l[:, i + 1, :]))
    disc_fake.append(ZGDiscrminator(p_z_g))
    disc_real.append(ZGDiscrminator(q_z_g))
    disc_fake.append(Discriminator(fake_x, p_z_g, p_z_l))
    disc_real.append(Discriminator(real_x, q_z_g, q_z_l))

    gen_params = lib.params_with_name('Generator')
    ext_params = lib.params_with_name('Extractor')
    disc_params = lib.params_with_name('Discriminator')
    rec_penalty = None
    if MODE == 'local_epce-z':
        rec_penalty = LAMBDA * lib.utils.distance.distance(real_x, rec_x, 'l2')
    gen_cost, disc_cost, _, _, gen_train_op, disc_train_op = lib.objs.gan_inference.weighted_local_epce(
        disc_fake, disc_real, ratio, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, rec_penalty=rec_penalty)
    ns.__dict__.update(real_x_unit=real_x_unit, real_x=real_x, q_z_l=q_z_l, q_z_g=q_z_g, p_z_l_0=p_z_l_0, p_z_l=p_z_l,
                       p_z_g=p_z_g, fake_x=fake_x, rec_x=rec_x, disc_fake=disc_fake, disc_real=disc_real, gen_params=gen_params,
                       ext_params=ext_params, disc_params=disc_params, gen_cost=gen_cost, disc_cost=disc_cost,
                       gen_train_op=gen_train_op, disc_train_op=disc_train_op)
    return ns


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='local_ep')
    ap.add_argument('--iters', type=int, default=40000)
    ap.add_argument('--batch-size', type=int, default=50)
    ap.add_argument('--len', type=int, default=31)
    args = ap.parse_args(argv)
    g = build_graph(MODE=args.mode, BATCH_SIZE=args.batch_size, LEN=args.len)
    rs = np.random.RandomState(0)
    with tf.Session() as session:
        for iteration in range(args.iters):                                                  # :600-625
            start_time = time.time()
            x = rs.randint(0, 256, size=(args.batch_size, args.len, g.OUTPUT_DIM)).astype('float32')   # synthetic sequences
            if iteration > 0:
                session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_unit: x})
            dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_unit: x})
            lib.plot.plot('train disc cost', dc)
            lib.plot.plot('time', time.time() - start_time)
            if (iteration < 5) or (iteration % 100 == 99):
                lib.plot.flush()
            lib.plot.tick()


if __name__ == '__main__':
    main()
