"""SSGAN on moving MNIST — Python-3 port of the reference's ssgan_inference_moving_mnist.py (MODE 'local_ep' /
'local_epce-z', POS_MODE 'naive_mean_field' — the defaults, :27-29) on the B200 kernels.

State-space latent: z_l^1 -> ImplicitOperator (3-layer MLP with a residual connection, ONE shared epsilon, unrolled LEN-1
times, :134-141) -> z_l^{1:T}; the frame generator / extractor / discriminator run on B*LEN images folded into the batch
dimension (:170-224, :267-311); LEN-1 pairwise latent discriminators + a z_g discriminator + the frame discriminator feed
weighted_local_epce with ratio = [1]*(LEN-1)+[1,LEN] normalised by 2*LEN (:78-79, :547).  Line numbers refer to
/root/reference/ssgan_inference_moving_mnist.py.  MODE 'ali' / 'alice-z' (:351-498, :541-547) replace the LEN+1 local
discriminators by ONE critic on (clip, all latents), in the three ALI_MODE variants of the reference: 'concat_x' (frames as
channels, Cin = LEN), 'concat_z' (per-frame trunk + 4x4 VALID conv, codes concatenated) and '3dcnn' (four Conv3D layers,
tflib/ops/conv3d.py; clips of LEN 4 or 16 as in the reference).
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


def build_graph(MODE='local_ep', BATCH_SIZE=50, LEN=16, DIM=32, DIM_OP=256, LR=1e-4, BN_FLAG=False, ALI_MODE='concat_x'):
    if MODE not in ('local_ep', 'local_epce-z', 'ali', 'alice-z'):
        raise NotImplementedError("unknown MODE %r" % (MODE,))                              # :497-498 raise('NotImplementedError')
    if MODE in ('ali', 'alice-z') and ALI_MODE not in ('concat_x', 'concat_z', '3dcnn'):
        raise NotImplementedError("unknown ALI_MODE %r" % (ALI_MODE,))                      # :494-495
    if MODE in ('ali', 'alice-z') and ALI_MODE == '3dcnn' and LEN not in (4, 16):
        raise NotImplementedError("the 3dcnn critic is defined for LEN 4 and 16 only (:367-370, :382-385)")
    DIM_LATENT_G, DIM_LATENT_L, N_C = 128, 8, 10
    DIM_LATENT_T = DIM_LATENT_L
    OUTPUT_SHAPE = [1, 64, 64]
    OUTPUT_DIM = int(np.prod(OUTPUT_SHAPE))
    LAMBDA, BETA1, BETA2 = 0.1, .5, .999
    BN_FLAG_G = BN_FLAG_E = BN_FLAG_D = BN_FLAG
    ratio = [1.0, ] * (LEN - 1) + [1, LEN]
    ratio = np.asarray(ratio) * 1.0 / (len(ratio) + LEN - 1)                                 # :78-79
    ns = types.SimpleNamespace(MODE=MODE, BATCH_SIZE=BATCH_SIZE, LEN=LEN, ratio=ratio, epsilons=[], OUTPUT_DIM=OUTPUT_DIM, N_C=N_C)

    def expand_labels(y):
        new_y = tf.tile(tf.expand_dims(y, axis=1), [1, LEN, 1])
        return tf.reshape(new_y, [-1, N_C])

    def LeakyReLU(x, alpha=0.2):
        return tf.maximum(alpha * x, x)

    def ImplicitOperator(z_l, epsilon, name):                                                # :98-114 (OP_DYN_MODE 'res')
        output = tf.concat([z_l, epsilon], axis=1)
        output = lib.ops.linear.Linear(name + '.Input', DIM_LATENT_L + DIM_LATENT_T, DIM_OP, output)
        output = LeakyReLU(output)
        output = lib.ops.linear.Linear(name + '.1', DIM_OP, DIM_OP, output)
        output = LeakyReLU(output)
        output = lib.ops.linear.Linear(name + '.Output', DIM_OP, DIM_LATENT_L, output)
        return output + z_l

    def DynamicGenerator(z_l_0):                                                             # :134-141
        z_list = [z_l_0, ]
        epsilon = tf.random_normal([BATCH_SIZE, DIM_LATENT_T])
        ns.epsilons.append(epsilon)
        for i in range(LEN - 1):
            z_list.append(ImplicitOperator(z_list[-1], epsilon, 'Generator.Dynamic'))
        return tf.reshape(tf.concat(z_list, axis=1), [BATCH_SIZE, LEN, DIM_LATENT_L])

    def DynamicExtractor(z_l_pre):                                                           # :143-168, naive_mean_field
        return z_l_pre

    def Generator(z_g, z_l, labels):                                                         # :170-205
        z_g = tf.reshape(z_g, [BATCH_SIZE, DIM_LATENT_G])
        z_g = tf.tile(tf.expand_dims(z_g, axis=1), [1, LEN, 1])
        z_l = tf.reshape(z_l, [BATCH_SIZE, LEN, DIM_LATENT_L])
        labels = tf.reshape(expand_labels(labels), [BATCH_SIZE, LEN, N_C])
        z = tf.concat([z_g, z_l, labels], axis=-1)
        z = tf.reshape(z, [BATCH_SIZE * LEN, DIM_LATENT_G + DIM_LATENT_L + N_C])
        output = lib.ops.linear.Linear('Generator.Input', DIM_LATENT_G + DIM_LATENT_L + N_C, 4 * 4 * 8 * DIM, z)
        if BN_FLAG_G:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN1', [0], output)
        output = tf.nn.relu(output)
        output = tf.reshape(output, [BATCH_SIZE * LEN, 8 * DIM, 4, 4])
        for i, (cin, cout) in enumerate(((8, 4), (4, 2), (2, 1))):
            output = lib.ops.deconv2d.Deconv2D('Generator.%d' % (i + 2), cin * DIM, cout * DIM, 5, output)
            if BN_FLAG_G:
                output = lib.ops.batchnorm.Batchnorm('Generator.BN%d' % (i + 2), [0, 2, 3], output)
            output = tf.nn.relu(output)
        output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM, 1, 5, output)
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

    def Extractor(inputs, labels):                                                           # :207-236
        output = tf.reshape(inputs, [BATCH_SIZE * LEN, ] + OUTPUT_SHAPE)
        labels = expand_labels(labels)
        output = _conv_trunk('Extractor.', output, 1, BN_FLAG_E)
        output = tf.reshape(output, [BATCH_SIZE * LEN, 4 * 4 * 8 * DIM])
        output = tf.concat([output, labels], axis=1)
        output = lib.ops.linear.Linear('Extractor.Output', 4 * 4 * 8 * DIM + N_C, DIM_LATENT_L, output)
        return tf.reshape(output, [BATCH_SIZE, LEN, DIM_LATENT_L])

    def G_Extractor(inputs, labels):                                                         # :238-265 (Cin = LEN)
        output = tf.reshape(inputs, [BATCH_SIZE, LEN, 64, 64])
        output = _conv_trunk('Extractor.G.', output, LEN, BN_FLAG_E)
        output = tf.reshape(output, [BATCH_SIZE, 4 * 4 * 8 * DIM])
        output = tf.concat([output, labels], axis=1)
        output = lib.ops.linear.Linear('Extractor.G.Output', 4 * 4 * 8 * DIM + N_C, DIM_LATENT_G, output)
        return tf.reshape(output, [BATCH_SIZE, DIM_LATENT_G])

    def Discriminator(x, z_g, z_l, labels):                                                  # :267-311
        output = tf.reshape(x, [BATCH_SIZE * LEN, ] + OUTPUT_SHAPE)
        labels = tf.reshape(expand_labels(labels), [BATCH_SIZE, LEN, N_C])
        z_g = tf.tile(tf.expand_dims(tf.reshape(z_g, [BATCH_SIZE, DIM_LATENT_G]), axis=1), [1, LEN, 1])
        z_l = tf.reshape(z_l, [BATCH_SIZE, LEN, DIM_LATENT_L])
        z = tf.reshape(tf.concat([z_g, z_l, labels], axis=-1), [BATCH_SIZE * LEN, DIM_LATENT_G + DIM_LATENT_L + N_C])
        output = _conv_trunk('Discriminator.', output, 1, BN_FLAG_D)            # dropout(training=False) is the identity
        output = tf.reshape(output, [BATCH_SIZE * LEN, 4 * 4 * 8 * DIM])
        z_output = LeakyReLU(lib.ops.linear.Linear('Discriminator.z1', DIM_LATENT_G + DIM_LATENT_L + N_C, 512, z))
        labels = tf.reshape(labels, [BATCH_SIZE * LEN, N_C])
        output = tf.concat([output, z_output, labels], 1)
        output = LeakyReLU(lib.ops.linear.Linear('Discriminator.zx1', 4 * 4 * 8 * DIM + 512 + N_C, 512, output))
        output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
        return tf.reshape(output, [BATCH_SIZE * LEN, ])

    def _ali_inputs(x, z_g, z_l, labels):
        z_l = tf.reshape(z_l, [BATCH_SIZE, LEN * DIM_LATENT_L])
        z_g = tf.reshape(z_g, [BATCH_SIZE, DIM_LATENT_G])
        labels = tf.reshape(labels, [BATCH_SIZE, N_C])
        return tf.concat([z_g, z_l, labels], axis=-1), labels

    def _ali_head(output, z, n_feat, extra=()):
        z_output = LeakyReLU(lib.ops.linear.Linear('Discriminator.z1', DIM_LATENT_G + DIM_LATENT_L * LEN + N_C, 512, z))
        output = tf.concat([output, z_output] + list(extra), 1)
        n_in = n_feat + 512 + sum(int(e.shape[1]) for e in extra)
        output = LeakyReLU(lib.ops.linear.Linear('Discriminator.zx1', n_in, 512, output))
        output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
        return tf.reshape(output, [BATCH_SIZE, ])

    def Discriminator3D(x, z_g, z_l, labels):                                                # :354-404 ALI_MODE '3dcnn'
        import tflib.ops.conv3d
        output = tf.reshape(x, [-1, LEN] + OUTPUT_SHAPE)
        output = tf.transpose(output, [0, 1, 3, 4, 2])                                       # NLHWC
        z, _ = _ali_inputs(x, z_g, z_l, labels)
        output = LeakyReLU(lib.ops.conv3d.Conv3D('Discriminator.1', 4, 1, DIM, 4, output, stride=2, stride_len=2))
        output = lib.ops.conv3d.Conv3D('Discriminator.2', 4, DIM, 2 * DIM, 4, output, stride=2, stride_len=1 if LEN == 4 else 2)
        if BN_FLAG_D:
            output = lib.ops.batchnorm.Batchnorm('Discriminator.BN2', [0, 1, 2, 3], output)
        output = LeakyReLU(output)
        output = lib.ops.conv3d.Conv3D('Discriminator.3', 4, 2 * DIM, 4 * DIM, 4, output, stride=2, stride_len=2)
        if BN_FLAG_D:
            output = lib.ops.batchnorm.Batchnorm('Discriminator.BN3', [0, 1, 2, 3], output)
        output = LeakyReLU(output)
        output = lib.ops.conv3d.Conv3D('Discriminator.4', 4, 4 * DIM, 8 * DIM, 4, output, stride=2, stride_len=1 if LEN == 4 else 2)
        if BN_FLAG_D:
            output = lib.ops.batchnorm.Batchnorm('Discriminator.BN4', [0, 1, 2, 3], output)
        output = LeakyReLU(output)
        output = tf.reshape(output, [BATCH_SIZE, 4 * 4 * 8 * DIM])
        return _ali_head(output, z, 4 * 4 * 8 * DIM)

    def DiscriminatorConcatX(x, z_g, z_l, labels):                                           # :407-446 frames as channels
        output = tf.reshape(x, [BATCH_SIZE, LEN, 64, 64])
        z, _ = _ali_inputs(x, z_g, z_l, labels)
        output = _conv_trunk('Discriminator.', output, LEN, BN_FLAG_D)
        output = tf.reshape(output, [BATCH_SIZE, 4 * 4 * 8 * DIM])
        return _ali_head(output, z, 4 * 4 * 8 * DIM)

    def DiscriminatorConcatZ(x, z_g, z_l, labels):                                           # :448-492 per-frame codes
        output = tf.reshape(x, [BATCH_SIZE * LEN, -1, 64, 64])
        z, labels = _ali_inputs(x, z_g, z_l, labels)
        output = _conv_trunk('Discriminator.', output, 1, BN_FLAG_D)
        output = lib.ops.conv2d.Conv2D('Discriminator.5', 8 * DIM, DIM_LATENT_G, 4, output, stride=1, padding='VALID')
        output = tf.reshape(output, [BATCH_SIZE, LEN * DIM_LATENT_G])
        return _ali_head(output, z, LEN * DIM_LATENT_G, extra=(labels,))

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

    # ---- graph (:505-549) ----
    real_x_unit = tf.placeholder(tf.float32, shape=[BATCH_SIZE, LEN, OUTPUT_DIM])
    real_x = 2 * (real_x_unit - .5)
    real_y = tf.placeholder(tf.float32, shape=[BATCH_SIZE, N_C])
    q_z_l_pre = Extractor(real_x, real_y)
    q_z_g = G_Extractor(real_x, real_y)
    q_z_l = DynamicExtractor(q_z_l_pre)
    rec_x = Generator(q_z_g, q_z_l, real_y)
    p_z_l_0 = tf.random_normal([BATCH_SIZE, DIM_LATENT_L])
    p_z_l = DynamicGenerator(p_z_l_0)
    p_z_g = tf.random_normal([BATCH_SIZE, DIM_LATENT_G])
    PI = tf.constant(np.asarray([1. / N_C, ] * N_C, dtype=np.float32))
    prior_y = tf.distributions.Categorical(probs=PI)
    p_y_idx = prior_y.sample(BATCH_SIZE)
    p_y = tf.one_hot(indices=p_y_idx, depth=N_C)
    fake_x = Generator(p_z_g, p_z_l, p_y)

    gen_params_of = lambda: (lib.params_with_name('Generator'), lib.params_with_name('Extractor'), lib.params_with_name('Discriminator'))
    rec_penalty = None
    if MODE in ('local_ep', 'local_epce-z'):
        disc_fake, disc_real = [], []
        for i in range(LEN - 1):
            disc_fake.append(DynamicDiscrminator(p_z_l[:, i, :], p_z_l[:, i + 1, :]))
            disc_real.append(DynamicDiscrminator(q_z_l[:, i, :], q_z_l[:, i + 1, :]))
        disc_fake.append(ZGDiscrminator(p_z_g))
        disc_real.append(ZGDiscrminator(q_z_g))
        disc_fake.append(Discriminator(fake_x, p_z_g, p_z_l, p_y))
        disc_real.append(Discriminator(real_x, q_z_g, q_z_l, real_y))
        gen_params, ext_params, disc_params = gen_params_of()
        if MODE == 'local_epce-z':
            rec_penalty = LAMBDA * lib.utils.distance.distance(real_x, rec_x, 'l2')
        gen_cost, disc_cost, _, _, gen_train_op, disc_train_op = lib.objs.gan_inference.weighted_local_epce(
            disc_fake, disc_real, ratio, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, rec_penalty=rec_penalty)
    else:                                                                                    # :537-547 one critic on everything
        critic = {'3dcnn': Discriminator3D, 'concat_x': DiscriminatorConcatX, 'concat_z': DiscriminatorConcatZ}[ALI_MODE]
        disc_real = critic(real_x, q_z_g, q_z_l, real_y)
        disc_fake = critic(fake_x, p_z_g, p_z_l, p_y)
        gen_params, ext_params, disc_params = gen_params_of()
        if MODE == 'ali':
            gen_cost, disc_cost, gen_train_op, disc_train_op = lib.objs.gan_inference.ali(
                disc_fake, disc_real, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, beta2=BETA2)
        else:
            rec_penalty = LAMBDA * lib.utils.distance.distance(real_x, rec_x, 'l2')
            gen_cost, disc_cost, gen_train_op, disc_train_op = lib.objs.gan_inference.alice(
                disc_fake, disc_real, rec_penalty, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1)
    ns.__dict__.update(real_x_unit=real_x_unit, real_y=real_y, real_x=real_x, q_z_l=q_z_l, q_z_g=q_z_g, p_z_l_0=p_z_l_0, p_z_l=p_z_l,
                       p_z_g=p_z_g, p_y_idx=p_y_idx, fake_x=fake_x, rec_x=rec_x, disc_fake=disc_fake, disc_real=disc_real, gen_params=gen_params,
                       ext_params=ext_params, disc_params=disc_params, gen_cost=gen_cost, disc_cost=disc_cost,
                       gen_train_op=gen_train_op, disc_train_op=disc_train_op)
    return ns


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='local_ep')
    ap.add_argument('--iters', type=int, default=100000)
    ap.add_argument('--batch-size', type=int, default=50)
    ap.add_argument('--len', type=int, default=16)
    args = ap.parse_args(argv)
    g = build_graph(MODE=args.mode, BATCH_SIZE=args.batch_size, LEN=args.len)
    rs = np.random.RandomState(0)
    with tf.Session() as session:
        for iteration in range(args.iters):
            start_time = time.time()
            x = rs.uniform(0, 1, size=(args.batch_size, args.len, g.OUTPUT_DIM)).astype('float32')    # synthetic frames
            y = np.eye(g.N_C, dtype='float32')[rs.randint(0, g.N_C, size=args.batch_size)]
            if iteration > 0:
                session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_unit: x, g.real_y: y})
            dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_unit: x, g.real_y: y})
            lib.plot.plot('train disc cost', dc)
            lib.plot.plot('time', time.time() - start_time)
            if (iteration < 5) or (iteration % 100 == 99):
                lib.plot.flush()
            lib.plot.tick()


if __name__ == '__main__':
    main()
