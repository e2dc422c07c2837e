"""ALI on CelebA 64x64 — Python-3 port of the reference's gan_inference_face.py (MODE='ali' only, :33,171-176; bs=128,
DIM_G=DIM_D=32, :36-42) on the B200 kernels.  This is BASELINE.json configs[3]: with torchrun the batch is sharded over
the ranks and the gradients are all-reduced once per optimiser step (graphical-gan_b200/gg/dist.py); the models have no
batch norm, so no statistic exchange is needed.  Input decode (:155-157): 2*((int/256)-.5) + U[0,1/128) dequantisation.
Line numbers refer to /root/reference/gan_inference_face.py."""
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
import tflib.ops.deconv2d
import tflib.objs.gan_inference
import tflib.plot


def build_graph(BATCH_SIZE=128, DIM_G=32, DIM_D=32, DIM_LATENT=128, LR=2e-4):
    OUTPUT_DIM = 64 * 64 * 3
    BETA1 = .5
    ns = types.SimpleNamespace(BATCH_SIZE=BATCH_SIZE, OUTPUT_DIM=OUTPUT_DIM, CRITIC_ITERS=1)

    def LeakyReLU(x, alpha=0.2):
        return tf.maximum(alpha * x, x)

    def Generator(noise):                                                                    # :64-82
        output = lib.ops.linear.Linear('Generator.Input', DIM_LATENT, 4 * 4 * 8 * DIM_G, noise)
        output = tf.nn.relu(output)
        output = tf.reshape(output, [-1, 8 * DIM_G, 4, 4])
        for i, (cin, cout) in enumerate(((8, 4), (4, 2), (2, 1))):
            output = lib.ops.deconv2d.Deconv2D('Generator.%d' % (i + 2), cin * DIM_G, cout * DIM_G, 5, output)
            output = tf.nn.relu(output)
        output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM_G, 3, 5, output)
        output = tf.tanh(output)
        return tf.reshape(output, [-1, OUTPUT_DIM])

    def _trunk(prefix, inputs, dim):
        output = tf.reshape(inputs, [-1, 3, 64, 64])
        output = LeakyReLU(lib.ops.conv2d.Conv2D(prefix + '.1', 3, dim, 5, output, stride=2))
        for i, (cin, cout) in enumerate(((1, 2), (2, 4), (4, 8))):
            output = LeakyReLU(lib.ops.conv2d.Conv2D('%s.%d' % (prefix, i + 2), cin * dim, cout * dim, 5, output, stride=2))
        return tf.reshape(output, [-1, 4 * 4 * 8 * dim])

    def Extractor(inputs):                                                                   # :84-99
        output = _trunk('Extractor', inputs, DIM_G)
        output = lib.ops.linear.Linear('Extractor.Output', 4 * 4 * 8 * DIM_G, DIM_LATENT, output)
        return tf.reshape(output, [-1, DIM_LATENT])

    def Discriminator(x, z):                                                                 # :101-131
        output = _trunk('Discriminator', x, DIM_D)
        z_output = LeakyReLU(lib.ops.linear.Linear('Discriminator.z1', DIM_LATENT, 512, z))
        output = tf.concat([output, z_output], 1)
        output = LeakyReLU(lib.ops.linear.Linear('Discriminator.zx1', 4 * 4 * 8 * DIM_D + 512, 512, output))
        output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
        return tf.reshape(output, [-1])

    real_x_int = tf.placeholder(tf.int32, shape=[BATCH_SIZE, OUTPUT_DIM])
    real_x = tf.reshape(2 * ((tf.cast(real_x_int, tf.float32) / 256.) - .5), [BATCH_SIZE, OUTPUT_DIM])
    dequant = tf.random_uniform(shape=[BATCH_SIZE, OUTPUT_DIM], minval=0., maxval=1. / 128)    # :157
    real_x = real_x + dequant
    q_z = Extractor(real_x)
    rec_x = Generator(q_z)
    p_z = tf.random_normal([BATCH_SIZE, DIM_LATENT])
    fake_x = Generator(p_z)
    disc_real = Discriminator(real_x, q_z)
    disc_fake = Discriminator(fake_x, p_z)
    gen_params = lib.params_with_name('Generator')
    ext_params = lib.params_with_name('Extractor')
    disc_params = lib.params_with_name('Discriminator')
    gen_cost, disc_cost, gen_train_op, disc_train_op = lib.objs.gan_inference.ali(
        disc_fake, disc_real, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1)
    ns.__dict__.update(real_x_int=real_x_int, real_x=real_x, dequant=dequant, q_z=q_z, p_z=p_z, fake_x=fake_x, rec_x=rec_x,
                       disc_real=disc_real, disc_fake=disc_fake, gen_params=gen_params, ext_params=ext_params,
                       disc_params=disc_params, gen_cost=gen_cost, disc_cost=disc_cost, gen_train_op=gen_train_op,
                       disc_train_op=disc_train_op)
    return ns


def main(argv=None):
    import argparse
    from gg import dist as ggdist
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=100000)
    ap.add_argument('--batch-size', type=int, default=128, help='GLOBAL batch; sharded over the ranks under torchrun')
    args = ap.parse_args(argv)
    rank, world = ggdist.init_from_env()
    np.random.seed(1234)
    g = build_graph(BATCH_SIZE=args.batch_size // world)
    rs = np.random.RandomState(100 + rank)
    with tf.Session() as session:
        for iteration in range(args.iters):
            start_time = time.time()
            batch = rs.randint(0, 256, size=(g.BATCH_SIZE, g.OUTPUT_DIM)).astype('int32')       # synthetic CelebA-shaped shard
            if iteration > 0:
                session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: batch})
            dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: batch})
            if rank == 0:
                lib.plot.plot('train disc cost', dc)
                lib.plot.plot('time', time.time() - start_time)
                if (iteration < 5) or (iteration % 100 == 99):
                    lib.plot.flush()
                lib.plot.tick()


if __name__ == '__main__':
    main()
