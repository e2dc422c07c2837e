"""ALI / ALICE / VEGAN / WALI(-GP) on MNIST — Python-3 port of the reference's gan_inference_mnist.py on the B200 kernels
(this is the script BASELINE.json configs[0] names; its LOCAL_EP variant is gmgan_inference_mnist.py).

1x28x28 images fed as float [0,1] (:248), sigmoid output (:142), the 8x8 -> 7x7 crop between the first two deconvolutions
(:134), 28 -> 14 -> 7 -> 4 strided convolutions, BN_FLAG = True for the non-vegan modes (:66-71) — here also INSIDE the
(x, z) critic (Discriminator.BN2 / BN3, :229-236), which also has two extra dense layers (Discriminator.2 on the z branch,
zx2; :241-254).  Consequences: sibling batching of D(fake) / D(real) stops at the critic's batch norms (rows are coupled
there: each tower keeps its own statistics, as in the reference), and MODE='wali-gp' needs the second-order gradient
of batch norm: gg/ops.py::_grad_bn_grad re-expresses the batch-norm gradient with primitive ops and differentiates those
(pinned against torch's double backward in tests/test_cpu_oracle_models.py; the reference's default MODE is 'ali').  Line numbers refer to /root/reference/gan_inference_mnist.py.
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
import tflib.objs.kl_aggregated
import tflib.objs.mmd
import tflib.utils.distance
import tflib.plot

NO_DISC = ['vegan-mmd', 'vegan-kl', 'vegan-ikl', 'vegan-jsd']                                # :46-47 CRITIC_ITERS = 0
SUPPORTED = ['ali', 'alice', 'alice-z', 'alice-x', 'vegan', 'vegan-wgan-gp', 'wali', 'wali-gp'] + NO_DISC


def build_graph(MODE='ali', BATCH_SIZE=50, DIM=64, LR=2e-4, BN_FLAG=None, Z_SAMPLES=100):
    if MODE == 'vae':
        # the reference builds this mode with rec_x_mean = rec_x_std = None (its Generator returns (x, None, None), :147) and
        # lib.objs.kl.vae(real_x, None, None, ...) fails at graph construction (:340): there is no behaviour to reproduce
        raise NotImplementedError("MODE 'vae' does not build in the reference either (Generator returns no mean / std, :147,340)")
    if MODE not in SUPPORTED:
        raise NotImplementedError("unknown MODE %r" % (MODE,))
    TYPE_Q = 'learn_std' if MODE in ['vegan-kl', 'vegan-ikl', 'vegan-jsd'] else 'no_std'     # :32-41 (TYPE_P is never read)
    DISTANCE_X = 'l2'
    CRITIC_ITERS = 0 if MODE in NO_DISC else (5 if MODE in ['vegan', 'vegan-wgan-gp', 'wali', 'wali-gp'] else 1)   # :46-51
    LAMBDA, BETA1, OUTPUT_DIM = 1., .5, 784
    small_code = MODE in ['vegan', 'vegan-wgan-gp', 'vegan-kl', 'vegan-jsd', 'vegan-ikl']    # :64-69
    if BN_FLAG is None:
        BN_FLAG = not small_code
    DIM_LATENT = 8 if small_code else 128
    N_VIS = BATCH_SIZE * 2
    DR_RATE = .2
    STD = .1                                                                                 # :42, for fix_std
    ns = types.SimpleNamespace(MODE=MODE, BATCH_SIZE=BATCH_SIZE, DIM_LATENT=DIM_LATENT, CRITIC_ITERS=CRITIC_ITERS, noise_layers=[])
    unit_std_z = tf.constant((STD * np.ones(shape=(BATCH_SIZE, DIM_LATENT))).astype('float32'))   # :86

    def LeakyReLU(x, alpha=0.2):
        return tf.maximum(alpha * x, x)

    def GaussianNoiseLayer(input_layer, std):                                               # :125-127
        noise = tf.random_normal(shape=tf.shape(input_layer), mean=0.0, stddev=std, dtype=tf.float32)
        ns.noise_layers.append(noise)
        return input_layer + noise

    def Generator(noise):                                                                   # :129-149
        output = lib.ops.linear.Linear('Generator.Input', DIM_LATENT, 4 * 4 * 4 * DIM, noise)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN1', [0], output)
        output = tf.nn.relu(output)
        output = tf.reshape(output, [-1, 4 * DIM, 4, 4])
        output = lib.ops.deconv2d.Deconv2D('Generator.2', 4 * DIM, 2 * DIM, 5, output)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN2', [0, 2, 3], output)
        output = tf.nn.relu(output)
        output = output[:, :, :7, :7]                                                        # :134
        output = lib.ops.deconv2d.Deconv2D('Generator.3', 2 * DIM, DIM, 5, output)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN3', [0, 2, 3], output)
        output = tf.nn.relu(output)
        output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM, 1, 5, output)
        output = tf.nn.sigmoid(output)
        return tf.reshape(output, [-1, OUTPUT_DIM]), None, None

    def Extractor(inputs):                                                                  # :151-180
        output = tf.reshape(inputs, [-1, 1, 28, 28])
        output = lib.ops.conv2d.Conv2D('Extractor.1', 1, DIM, 5, output, stride=2)
        output = LeakyReLU(output)
        output = lib.ops.conv2d.Conv2D('Extractor.2', DIM, 2 * DIM, 5, output, stride=2)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Extractor.BN2', [0, 2, 3], output)
        output = LeakyReLU(output)
        output = lib.ops.conv2d.Conv2D('Extractor.3', 2 * DIM, 4 * DIM, 5, output, stride=2)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Extractor.BN3', [0, 2, 3], output)
        output = LeakyReLU(output)
        output = tf.reshape(output, [-1, 4 * 4 * 4 * DIM])
        std = mean = None
        if TYPE_Q == 'learn_std':                                                           # :164-166
            log_std = lib.ops.linear.Linear('Extractor.Std', 4 * 4 * 4 * DIM, DIM_LATENT, output)
            std = tf.exp(log_std)
        output = lib.ops.linear.Linear('Extractor.Output', 4 * 4 * 4 * DIM, DIM_LATENT, output)
        if TYPE_Q == 'learn_std':                                                           # :175-178 reparameterised code
            epsilon = tf.random_normal(unit_std_z.shape)
            ns.noise_layers.append(epsilon)
            mean = output
            output = tf.add(mean, tf.multiply(epsilon, std))
        return tf.reshape(output, [-1, DIM_LATENT]), mean, std

    if MODE in NO_DISC:
        Discriminator = None                                                                # :213-214 no discriminator
    elif MODE in ['vegan', 'vegan-wgan-gp']:
        def Discriminator(z):                                                               # :184-209 critic on the code
            output = GaussianNoiseLayer(z, std=.3)
            output = lib.ops.linear.Linear('Discriminator.Input', DIM_LATENT, 1024, output)
            if BN_FLAG:
                output = lib.ops.batchnorm.Batchnorm('Discriminator.BN1', [0], output)
            output = LeakyReLU(output)
            output = GaussianNoiseLayer(output, std=.5)
            output = lib.ops.linear.Linear('Discriminator.2', 1024, 512, output)
            if BN_FLAG:
                output = lib.ops.batchnorm.Batchnorm('Discriminator.BN2', [0], output)
            output = LeakyReLU(output)
            output = GaussianNoiseLayer(output, std=.5)
            output = lib.ops.linear.Linear('Discriminator.3', 512, 256, output)
            if BN_FLAG:
                output = lib.ops.batchnorm.Batchnorm('Discriminator.BN3', [0], output)
            output = LeakyReLU(output)
            output = GaussianNoiseLayer(output, std=.5)
            output = lib.ops.linear.Linear('Discriminator.4', 256, 256, output)
            if BN_FLAG:
                output = lib.ops.batchnorm.Batchnorm('Discriminator.BN4', [0], output)
            output = LeakyReLU(output)
            output = lib.ops.linear.Linear('Discriminator.Output', 256, 1, output)
            return tf.reshape(output, [-1])
    else:
        def Discriminator(x, z):                                                            # :221-256 critic on (x, z)
            output = tf.reshape(x, [-1, 1, 28, 28])
            output = lib.ops.conv2d.Conv2D('Discriminator.1', 1, DIM, 5, output, stride=2)
            output = LeakyReLU(output)
            output = lib.ops.conv2d.Conv2D('Discriminator.2', DIM, 2 * DIM, 5, output, stride=2)
            if BN_FLAG:
                output = lib.ops.batchnorm.Batchnorm('Discriminator.BN2', [0, 2, 3], output)
            output = LeakyReLU(output)
            output = lib.ops.conv2d.Conv2D('Discriminator.3', 2 * DIM, 4 * DIM, 5, output, stride=2)
            if BN_FLAG:
                output = lib.ops.batchnorm.Batchnorm('Discriminator.BN3', [0, 2, 3], output)
            output = LeakyReLU(output)
            output = tf.reshape(output, [-1, 4 * 4 * 4 * DIM])
            z_output = lib.ops.linear.Linear('Discriminator.z1', DIM_LATENT, 512, z)
            z_output = LeakyReLU(z_output)
            z_output = tf.layers.dropout(z_output, rate=DR_RATE)
            z_output = lib.ops.linear.Linear('Discriminator.2', 512, 512, z_output)       # (sic) shares the prefix of the conv
            z_output = LeakyReLU(z_output)
            z_output = tf.layers.dropout(z_output, rate=DR_RATE)
            output = tf.concat([output, z_output], 1)
            output = lib.ops.linear.Linear('Discriminator.zx1', 4 * 4 * 4 * DIM + 512, 512, output)
            output = LeakyReLU(output)
            output = tf.layers.dropout(output, rate=DR_RATE)
            output = lib.ops.linear.Linear('Discriminator.zx2', 512, 512, output)
            output = LeakyReLU(output)
            output = tf.layers.dropout(output, rate=DR_RATE)
            output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
            return tf.reshape(output, [-1])

    # ---- losses (:246-360) ----
    real_x = tf.placeholder(tf.float32, shape=[BATCH_SIZE, OUTPUT_DIM])                      # :248: float [0,1] images
    q_z, q_z_mean, q_z_std = Extractor(real_x)
    rec_x, _, _ = Generator(q_z)
    p_z = tf.random_normal([BATCH_SIZE, DIM_LATENT])
    fake_x, _, _ = Generator(p_z)
    rec_z, _, _ = Extractor(fake_x)
    if MODE in ['vegan-kl', 'vegan-ikl', 'vegan-jsd']:                                       # :263-265 prior of the MC estimate
        p_z_mean = tf.constant((np.zeros(shape=(Z_SAMPLES, DIM_LATENT))).astype('float32'))
        p_z_std = tf.constant((np.ones(shape=(Z_SAMPLES, DIM_LATENT))).astype('float32'))

    disc_real = disc_fake = None
    if MODE in NO_DISC:
        pass
    elif MODE in ['vegan', 'vegan-wgan-gp']:
        disc_real = Discriminator(p_z)
        disc_fake = Discriminator(q_z)
    else:
        disc_real = Discriminator(real_x, q_z)
        disc_fake = Discriminator(fake_x, p_z)

    gen_params = lib.params_with_name('Generator')
    ext_params = lib.params_with_name('Extractor')
    disc_params = lib.params_with_name('Discriminator')
    gi = lib.objs.gan_inference
    ge = gen_params + ext_params
    alpha = gradient_penalty = rec_penalty = clip_disc_weights = None
    if MODE == 'ali':
        costs = gi.ali(disc_fake, disc_real, ge, disc_params, lr=LR, beta1=BETA1)
    elif MODE in ('alice', 'alice-z', 'alice-x'):
        rec_penalty = 0
        if MODE in ('alice', 'alice-z'):
            rec_penalty = rec_penalty + 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        if MODE in ('alice', 'alice-x'):
            rec_penalty = rec_penalty + 1. * lib.utils.distance.distance(p_z, rec_z, DISTANCE_X)
        costs = gi.alice(disc_fake, disc_real, rec_penalty, ge, disc_params, lr=LR, beta1=BETA1)
    elif MODE == 'vegan':
        rec_penalty = 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        costs = gi.vegan(disc_fake, disc_real, rec_penalty, ge, disc_params, LAMBDA, lr=LR, beta1=BETA1)
    elif MODE == 'vegan-wgan-gp':                                                           # :302-316
        alpha = tf.random_uniform(shape=[BATCH_SIZE, 1], minval=0., maxval=1.)
        differences = q_z - p_z
        interpolates = p_z + (alpha * differences)
        gradients = tf.gradients(Discriminator(interpolates), interpolates)[0]
        slopes = tf.sqrt(tf.reduce_sum(tf.square(gradients), reduction_indices=[1]))
        gradient_penalty = 10. * (tf.reduce_mean((slopes - 1.) ** 2))
        rec_penalty = 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        costs = gi.vegan_wgan_gp(disc_fake, disc_real, rec_penalty, gradient_penalty, ge, disc_params, LAMBDA, lr=LR, beta1=BETA1)
    elif MODE == 'wali':
        gen_cost, disc_cost, clip_disc_weights, gen_train_op, disc_train_op, clip_ops = gi.wali(disc_fake, disc_real, ge, disc_params)
        costs = (gen_cost, disc_cost, gen_train_op, disc_train_op)
    elif MODE == 'wali-gp':                                                                 # :342-357
        alpha = tf.random_uniform(shape=[BATCH_SIZE, 1], minval=0., maxval=1.)
        differences = fake_x - real_x
        interpolates = real_x + (alpha * differences)
        differences_z = p_z - q_z
        interpolates_z = q_z + (alpha * differences_z)
        gradients = tf.gradients(Discriminator(interpolates, interpolates_z), [interpolates, interpolates_z])[0]   # x part only
        slopes = tf.sqrt(tf.reduce_sum(tf.square(gradients), reduction_indices=[1]))
        gradient_penalty = 10. * (tf.reduce_mean((slopes - 1.) ** 2))
        costs = gi.wali_gp(disc_fake, disc_real, gradient_penalty, ge, disc_params)
    elif MODE in NO_DISC:                                                                   # :322-336 generator objective only
        rec_penalty = 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        ka = lib.objs.kl_aggregated
        if MODE == 'vegan-mmd':
            gc, gop = lib.objs.mmd.vegan_mmd(q_z, p_z, rec_penalty, ge, BATCH_SIZE, LAMBDA, lr=LR, beta1=BETA1)
        elif MODE == 'vegan-kl':
            gc, gop = ka.vegan_kl(q_z_mean, q_z_std, p_z_mean, p_z_std, rec_penalty, ge, Z_SAMPLES, BATCH_SIZE, DIM_LATENT, LAMBDA,
                                  lr=LR, beta1=BETA1)
        elif MODE == 'vegan-ikl':
            gc, gop = ka.vegan_ikl(q_z_mean, q_z_std, p_z_mean, p_z_std, rec_penalty, ge, Z_SAMPLES, DIM_LATENT, LAMBDA,
                                   lr=LR, beta1=BETA1)
        else:
            gc, gop = ka.vegan_jsd(q_z_mean, q_z_std, p_z_mean, p_z_std, rec_penalty, ge, Z_SAMPLES, BATCH_SIZE, DIM_LATENT, LAMBDA,
                                   lr=LR, beta1=BETA1)
        costs = (gc, None, gop, None)
    gen_cost, disc_cost, gen_train_op, disc_train_op = costs

    np_fixed = np.random.normal(size=(N_VIS, DIM_LATENT)).astype('float32')
    fixed_noise_samples, _, _ = Generator(tf.constant(np_fixed))
    ns.__dict__.update(real_x=real_x, q_z=q_z, p_z=p_z, fake_x=fake_x, rec_x=rec_x, rec_z=rec_z, alpha=alpha,
                       disc_fake=disc_fake, disc_real=disc_real, gradient_penalty=gradient_penalty, rec_penalty=rec_penalty,
                       gen_params=gen_params, ext_params=ext_params, disc_params=disc_params, gen_cost=gen_cost, disc_cost=disc_cost,
                       gen_train_op=gen_train_op, disc_train_op=disc_train_op, clip_disc_weights=clip_disc_weights,
                       fixed_noise_samples=fixed_noise_samples)
    return ns


def main(argv=None):
    import argparse
    from gmgan_inference_mnist import synthetic_batches
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='ali', choices=SUPPORTED)
    ap.add_argument('--iters', type=int, default=200000)
    ap.add_argument('--batch-size', type=int, default=50)
    ap.add_argument('--synthetic', action='store_true')
    args = ap.parse_args(argv)
    g = build_graph(MODE=args.mode, BATCH_SIZE=args.batch_size)
    gen = None
    if not args.synthetic and os.path.isfile('/tmp/mnist.pkl.gz'):
        import tflib.mnist
        train_gen, _, _ = lib.mnist.load(args.batch_size, args.batch_size)

        def inf_train_gen():
            while True:
                for images, _ in train_gen():
                    yield images
        gen = inf_train_gen()
    if gen is None:
        gen = synthetic_batches(args.batch_size)
    with tf.Session() as session:
        session.run(tf.global_variables_initializer())
        for iteration in range(args.iters):                                                  # :400-430
            start_time = time.time()
            if iteration > 0:
                gc, _ = session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x: next(gen)})
            dc = None
            for i in range(g.CRITIC_ITERS):
                dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x: next(gen)})
                if g.clip_disc_weights is not None:
                    session.run(g.clip_disc_weights)
            if g.CRITIC_ITERS == 0:                                                          # :429-431 no-discriminator modes
                if iteration > 0:
                    lib.plot.plot('train gen cost ', gc)
            else:
                lib.plot.plot('train disc cost', dc)
            lib.plot.plot('time', time.time() - start_time)
            if (iteration < 5) or (iteration % 100 == 99):
                lib.plot.flush()
            lib.plot.tick()


if __name__ == '__main__':
    main()
