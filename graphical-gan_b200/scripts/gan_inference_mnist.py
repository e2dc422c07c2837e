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
import tflib.utils.distance
import tflib.plot

SUPPORTED = ['ali', 'alice', 'alice-z', 'alice-x', 'vegan', 'vegan-wgan-gp', 'wali', 'wali-gp']


def build_graph(MODE='ali', BATCH_SIZE=50, DIM=64, LR=2e-4, BN_FLAG=None):
    if MODE not in SUPPORTED:
        raise NotImplementedError("MODE %r has no discriminator (VAE / MMD / KL baselines are off the adversarial hot path)" % MODE)
    DISTANCE_X = 'l2'
    CRITIC_ITERS = 5 if MODE in ['vegan', 'vegan-wgan-gp', 'wali', 'wali-gp'] else 1        # :46-51
    LAMBDA, BETA1, OUTPUT_DIM = 1., .5, 784
    if BN_FLAG is None:                                                                      # :66-71
        BN_FLAG = MODE not in ('vegan', 'vegan-wgan-gp')
    DIM_LATENT = 8 if MODE in ['vegan', 'vegan-wgan-gp'] else 128
    N_VIS = BATCH_SIZE * 2
    DR_RATE = .2
    ns = types.SimpleNamespace(MODE=MODE, BATCH_SIZE=BATCH_SIZE, DIM_LATENT=DIM_LATENT, CRITIC_ITERS=CRITIC_ITERS, noise_layers=[])

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
        output = lib.ops.linear.Linear('Extractor.Output', 4 * 4 * 4 * DIM, DIM_LATENT, output)
        return tf.reshape(output, [-1, DIM_LATENT]), None, None

    if MODE in ['vegan', 'vegan-wgan-gp']:
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
    q_z, _, _ = Extractor(real_x)
    rec_x, _, _ = Generator(q_z)
    p_z = tf.random_normal([BATCH_SIZE, DIM_LATENT])
    fake_x, _, _ = Generator(p_z)
    rec_z, _, _ = Extractor(fake_x)

    if MODE in ['vegan', 'vegan-wgan-gp']:
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
                session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x: next(gen)})
            for i in range(g.CRITIC_ITERS):
                dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x: next(gen)})
                if g.clip_disc_weights is not None:
                    session.run(g.clip_disc_weights)
            lib.plot.plot('train disc cost', dc)
            lib.plot.plot('time', time.time() - start_time)
            if (iteration < 5) or (iteration % 100 == 99):
                lib.plot.flush()
            lib.plot.tick()


if __name__ == '__main__':
    main()
