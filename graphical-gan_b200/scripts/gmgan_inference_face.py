"""GMGAN on CelebA 64x64 — Python-3 port of the reference's gmgan_inference_face.py (MODE 'local_ep' / 'ali', :33; bs=128,
DIM_G = DIM_D = 32, N_COMS = 100, :35-47) on the B200 kernels: the four-layer 5x5 stride-2 conv / deconv stacks of
gan_inference_face.py plus the mixture-of-Gaussians prior and the Gumbel-softmax soft assignment of the GMGAN scripts
(:94-106), no batch norm.  Input decode (:241-243): 2*((int/256)-.5) + U[0,1/128) dequantisation.  Under torchrun the batch is
sharded over the ranks (gg/dist.py).  Line numbers refer to /root/reference/gmgan_inference_face.py.
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
import tflib.ops.deconv2d
import tflib.objs.gan_inference
import tflib.plot


def build_graph(MODE='local_ep', BATCH_SIZE=128, DIM_G=32, DIM_D=32, DIM_LATENT=128, N_COMS=100, LR=2e-4, N_VIS=None):
    OUTPUT_DIM = 64 * 64 * 3
    BETA1, BETA2 = .5, .999
    TEMP = .1
    if N_VIS is None:
        N_VIS = N_COMS * 10
    assert N_VIS % N_COMS == 0
    ns = types.SimpleNamespace(MODE=MODE, BATCH_SIZE=BATCH_SIZE, OUTPUT_DIM=OUTPUT_DIM, CRITIC_ITERS=1, N_COMS=N_COMS,
                               DIM_LATENT=DIM_LATENT, gumbel_uniforms=[])

    PI = tf.constant(np.asarray([1. / N_COMS, ] * N_COMS, dtype=np.float32))                 # :79-80
    prior_k = tf.distributions.Categorical(probs=PI)

    def sample_gumbel(shape, eps=1e-20):                                                     # :82-85
        U = tf.random_uniform(shape, minval=0, maxval=1)
        ns.gumbel_uniforms.append(U)
        return -tf.log(-tf.log(U + eps) + eps)

    def LeakyReLU(x, alpha=0.2):
        return tf.maximum(alpha * x, x)

    def HyperGenerator(hyper_k, hyper_noise):                                                # :94-97
        com_mu = lib.param('Generator.Hyper.Mu', np.random.normal(size=(N_COMS, DIM_LATENT)).astype('float32'))
        return tf.add(tf.matmul(tf.cast(hyper_k, tf.float32), com_mu), hyper_noise)

    def HyperExtractor(latent_z):                                                            # :100-106
        com_mu = lib.param('Generator.Hyper.Mu', np.random.normal(size=(N_COMS, DIM_LATENT)).astype('float32'))
        com_logits = -.5 * tf.reduce_sum(tf.pow((tf.expand_dims(latent_z, axis=1) - tf.expand_dims(com_mu, axis=0)), 2),
                                         axis=-1) + tf.expand_dims(tf.log(PI), axis=0)
        k = tf.nn.softmax((com_logits + sample_gumbel(tf.shape(com_logits))) / TEMP)
        return com_logits, k

    def Generator(noise):                                                                    # :108-125
        output = lib.ops.linear.Linear('Generator.Input', DIM_LATENT, 4 * 4 * 8 * DIM_G, noise)
        output = tf.nn.relu(output)
        output = tf.reshape(output, [-1, 8 * DIM_G, 4, 4])
        for i, (cin, cout) in enumerate(((8, 4), (4, 2), (2, 1))):
            output = lib.ops.deconv2d.Deconv2D('Generator.%d' % (i + 2), cin * DIM_G, cout * DIM_G, 5, output)
            output = tf.nn.relu(output)
        output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM_G, 3, 5, output)
        output = tf.tanh(output)
        return tf.reshape(output, [-1, OUTPUT_DIM])

    def _trunk(prefix, inputs, dim, dropout):
        output = tf.reshape(inputs, [-1, 3, 64, 64])
        for i, (cin, cout) in enumerate(((None, 1), (1, 2), (2, 4), (4, 8))):
            output = lib.ops.conv2d.Conv2D('%s.%d' % (prefix, i + 1), 3 if cin is None else cin * dim, cout * dim, 5, output, stride=2)
            output = LeakyReLU(output)
            if dropout:
                output = tf.layers.dropout(output, rate=.2)
        return tf.reshape(output, [-1, 4 * 4 * 8 * dim])

    def Extractor(inputs):                                                                   # :127-146
        output = _trunk('Extractor', inputs, DIM_G, False)
        output = lib.ops.linear.Linear('Extractor.Output', 4 * 4 * 8 * DIM_G, DIM_LATENT, output)
        return tf.reshape(output, [-1, DIM_LATENT])

    def _mlp_head(name_in, n_in, x):
        output = LeakyReLU(lib.ops.linear.Linear(name_in, n_in, 512, x))
        return tf.layers.dropout(output, rate=.2)

    if MODE in ['local_ep', 'local_epce']:
        def HyperDiscriminator(z, k):                                                        # :150-166
            output = _mlp_head('Discriminator.HyperInput', DIM_LATENT + N_COMS, tf.concat([z, k], 1))
            output = _mlp_head('Discriminator.Hyper2', 512, output)
            output = _mlp_head('Discriminator.Hyper3', 512, output)
            output = lib.ops.linear.Linear('Discriminator.HyperOutput', 512, 1, output)
            return tf.reshape(output, [-1])

        def Discriminator(x, z):                                                             # :168-200
            output = _trunk('Discriminator', x, DIM_D, True)
            z_output = _mlp_head('Discriminator.z1', DIM_LATENT, z)
            output = tf.concat([output, z_output], 1)
            output = _mlp_head('Discriminator.zx1', 4 * 4 * 8 * DIM_D + 512, output)
            output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
            return tf.reshape(output, [-1])
    else:
        def Discriminator(x, z, k):                                                          # :204-237
            output = _trunk('Discriminator', x, DIM_D, True)
            zk_output = _mlp_head('Discriminator.zk1', DIM_LATENT + N_COMS, tf.concat([z, k], 1))
            output = tf.concat([output, zk_output], 1)
            output = _mlp_head('Discriminator.zxk1', 4 * 4 * 8 * DIM_D + 512, output)
            output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
            return tf.reshape(output, [-1])

    # ---- losses (:239-275) ----
    real_x_int = tf.placeholder(tf.int32, shape=[BATCH_SIZE, OUTPUT_DIM])
    real_x = tf.reshape(2 * ((tf.cast(real_x_int, tf.float32) / 256.) - .5), [BATCH_SIZE, OUTPUT_DIM])
    dequant = tf.random_uniform(shape=[BATCH_SIZE, OUTPUT_DIM], minval=0., maxval=1. / 128)    # :243
    real_x = real_x + dequant
    q_z = Extractor(real_x)
    q_k_logits, q_k = HyperExtractor(q_z)
    rec_x = Generator(q_z)
    hyper_p_z = tf.random_normal([BATCH_SIZE, DIM_LATENT])
    hyper_p_k_idx = prior_k.sample(BATCH_SIZE)
    hyper_p_k = tf.one_hot(indices=hyper_p_k_idx, depth=N_COMS)
    p_z = HyperGenerator(hyper_p_k, hyper_p_z)
    fake_x = Generator(p_z)
    if MODE in ['local_ep', 'local_epce']:
        disc_fake = [HyperDiscriminator(p_z, hyper_p_k), Discriminator(fake_x, p_z)]
        disc_real = [HyperDiscriminator(q_z, q_k), Discriminator(real_x, q_z)]
    else:
        disc_real = Discriminator(real_x, q_z, q_k)
        disc_fake = Discriminator(fake_x, p_z, hyper_p_k)
    gen_params = lib.params_with_name('Generator')
    ext_params = lib.params_with_name('Extractor')
    disc_params = lib.params_with_name('Discriminator')
    gi = lib.objs.gan_inference
    if MODE == 'ali':
        costs = gi.ali(disc_fake, disc_real, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, beta2=BETA2)
    elif MODE == 'local_ep':
        costs = gi.local_ep(disc_fake, disc_real, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, beta2=BETA2)
    else:
        raise NotImplementedError(MODE)                                                      # (:277, raise('NotImplementedError'))
    gen_cost, disc_cost, gen_train_op, disc_train_op = costs

    np_fixed_noise = np.random.normal(size=(N_VIS, DIM_LATENT)).astype('float32')            # :281-286
    np_fixed_k = np.tile(np.eye(N_COMS, dtype=int), (N_VIS // N_COMS, 1))
    fixed_noise_samples = Generator(HyperGenerator(tf.constant(np_fixed_k), tf.constant(np_fixed_noise)))

    ns.__dict__.update(real_x_int=real_x_int, real_x=real_x, dequant=dequant, q_z=q_z, q_k=q_k, q_k_logits=q_k_logits,
                       rec_x=rec_x, hyper_p_z=hyper_p_z, hyper_p_k_idx=hyper_p_k_idx, hyper_p_k=hyper_p_k, p_z=p_z,
                       fake_x=fake_x, disc_fake=disc_fake, disc_real=disc_real, gen_params=gen_params, ext_params=ext_params,
                       disc_params=disc_params, gen_cost=gen_cost, disc_cost=disc_cost, gen_train_op=gen_train_op,
                       disc_train_op=disc_train_op, fixed_noise_samples=fixed_noise_samples)
    return ns


def main(argv=None):
    import argparse
    from gg import dist as ggdist
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='local_ep')
    ap.add_argument('--iters', type=int, default=100000)
    ap.add_argument('--batch-size', type=int, default=128, help='GLOBAL batch; sharded over the ranks under torchrun')
    ap.add_argument('--data', default='./dataset/celebA/celebA_64x64.npy', help='tflib/celebA.py:21-35: [N,3,64,64] uint8')
    ap.add_argument('--synthetic', action='store_true')
    args = ap.parse_args(argv)
    rank, world = ggdist.init_from_env()
    np.random.seed(1234)
    g = build_graph(MODE=args.mode, BATCH_SIZE=args.batch_size // world)
    rs = np.random.RandomState(100 + rank)
    data = None
    if not args.synthetic and os.path.exists(args.data):
        data = np.load(args.data, mmap_mode='r')

    def next_batch():
        if data is None:
            return rs.randint(0, 256, size=(g.BATCH_SIZE, g.OUTPUT_DIM)).astype('int32')
        idx = np.sort(rs.choice(len(data), g.BATCH_SIZE, replace=False))
        return np.asarray(data[idx]).reshape(g.BATCH_SIZE, -1).astype('int32')
    with tf.Session() as session:
        for iteration in range(args.iters):                                                  # :318-335
            start_time = time.time()
            if iteration > 0:
                session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: next_batch()})
            dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: next_batch()})
            if rank == 0:
                lib.plot.plot('train disc cost', dc)
                lib.plot.plot('time', time.time() - start_time)
                if (iteration < 5) or (iteration % 100 == 99):
                    lib.plot.flush()
                lib.plot.tick()


if __name__ == '__main__':
    main()
