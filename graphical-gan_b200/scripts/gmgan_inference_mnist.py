"""GMGAN on MNIST — Python-3 port of the reference's gmgan_inference_mnist.py driving the B200 kernels
(BASELINE.json configs[0]: MODE='local_ep', BATCH_SIZE=50, 28x28x1 — SURVEY.md D2).

The constants block, the model functions (Generator / Extractor / HyperGenerator / HyperExtractor /
HyperDiscriminator / Discriminator) and the `session.run` training loop keep the reference's structure and
names (line references below are to /root/reference/gmgan_inference_mnist.py).  Differences, all mechanical:
  * the graph is built inside build_graph() so that tests and bench.py can import it (the reference builds it at
    module import); random tensors are also returned so parity tests can FEED them (TF lets you feed any tensor);
  * Python 3 syntax; matplotlib / inception-score / image-grid side outputs are optional and skipped when their
    dependencies or datasets are absent; `--synthetic` trains on uniform random [0,1] images (no dataset needed).
"""
import os
import sys
import time
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import tensorflow as tf          # the gg shim (graphical-gan_b200/tensorflow)
import tflib as lib
import tflib.ops.linear
import tflib.ops.conv2d
import tflib.ops.batchnorm
import tflib.ops.deconv2d
import tflib.objs.gan_inference
import tflib.utils.distance
import tflib.plot


def build_graph(MODE='local_ep', BATCH_SIZE=50, DIM=64, N_COMS=30, LR=2e-4, MODE_K='CONCRETE', N_VIS=None):
    """Lines 31-394 of the reference script.  Returns a namespace with every tensor / op the train loop uses."""
    # ---- hyperparameters (:39-87) ----
    if MODE in ['vegan-kl', 'vegan-ikl', 'vegan-jsd', 'vae', 'vegan-mmd']:
        raise NotImplementedError("MODE %s has no discriminator; not on the adversarial hot path" % MODE)
    d_list = ['alice', 'alice-z', 'alice-x', 'vegan', 'vegan-wgan-gp', 'local_epce']
    DISTANCE_X = 'l2' if MODE in d_list else None
    CRITIC_ITERS = 5 if MODE in ['vegan', 'vegan-wgan-gp', 'wali', 'wali-gp'] else 1
    LAMBDA = 1.
    BETA1 = .5
    OUTPUT_DIM = 784
    if MODE in ['vegan', 'vegan-wgan-gp']:
        BN_FLAG, DIM_LATENT = False, 8
    else:
        BN_FLAG, DIM_LATENT = True, 128
    if N_VIS is None:
        N_VIS = N_COMS * 10
    assert N_VIS % N_COMS == 0
    TEMP = .1
    CONTROL_VARIATE = .0
    DR_RATE = .2
    TYPE_Q = 'no_std'

    ns = types.SimpleNamespace(MODE=MODE, BATCH_SIZE=BATCH_SIZE, DIM=DIM, N_COMS=N_COMS, DIM_LATENT=DIM_LATENT,
                               OUTPUT_DIM=OUTPUT_DIM, CRITIC_ITERS=CRITIC_ITERS, BN_FLAG=BN_FLAG, N_VIS=N_VIS,
                               gumbel_uniforms=[])

    # ---- prior (:114-115) ----
    PI = tf.constant(np.asarray([1. / N_COMS, ] * N_COMS, dtype=np.float32))
    prior_k = tf.distributions.Categorical(probs=PI)

    def sample_gumbel(shape, eps=1e-20):
        U = tf.random_uniform(shape, minval=0, maxval=1)
        ns.gumbel_uniforms.append(U)
        return -tf.log(-tf.log(U + eps) + eps)

    def LeakyReLU(x, alpha=0.2):
        return tf.maximum(alpha * x, x)

    # ---- mixture prior / soft assignment (:150-173) ----
    def HyperGenerator(hyper_k, hyper_noise):
        com_mu = lib.param('Generator.Hyper.Mu', np.random.normal(size=(N_COMS, DIM_LATENT)).astype('float32'))
        return tf.add(tf.matmul(tf.cast(hyper_k, tf.float32), com_mu), hyper_noise)

    def HyperExtractor(latent_z):
        com_mu = lib.param('Generator.Hyper.Mu', np.random.normal(size=(N_COMS, DIM_LATENT)).astype('float32'))
        com_logits = -.5 * tf.reduce_sum(tf.pow((tf.expand_dims(latent_z, axis=1) - tf.expand_dims(com_mu, axis=0)), 2),
                                         axis=-1) + tf.expand_dims(tf.log(PI), axis=0)
        if MODE_K == 'REINFORCE':
            k = tf.one_hot(indices=tf.argmax(com_logits, axis=-1), depth=N_COMS)
        elif MODE_K == 'CONCRETE':
            k = tf.nn.softmax((com_logits + sample_gumbel(tf.shape(com_logits))) / TEMP)
        elif MODE_K == 'STRAIGHT_THROUGHT_CONCRETE':
            k = tf.nn.softmax((com_logits + sample_gumbel(tf.shape(com_logits))) / TEMP)
            k_hard = tf.one_hot(indices=tf.argmax(k, axis=-1), depth=N_COMS)
            k = tf.stop_gradient(k_hard - k) + k
        elif MODE_K == 'STRAIGHT_THROUGHT':
            k_hard = tf.one_hot(indices=tf.argmax(com_logits, axis=-1), depth=N_COMS)
            k = tf.stop_gradient(k_hard - com_logits) + com_logits
        return com_logits, k

    # ---- networks (:175-336) ----
    def Generator(noise):
        output = lib.ops.linear.Linear('Generator.Input', DIM_LATENT, 4 * 4 * 4 * DIM, noise)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN1', [0], output)
        output = tf.nn.relu(output)
        output = tf.reshape(output, [-1, 4 * DIM, 4, 4])

        output = lib.ops.deconv2d.Deconv2D('Generator.2', 4 * DIM, 2 * DIM, 5, output)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN2', [0, 2, 3], output)
        output = tf.nn.relu(output)

        output = output[:, :, :7, :7]                                            # (:179) 8x8 -> 7x7 so that 7 -> 14 -> 28

        output = lib.ops.deconv2d.Deconv2D('Generator.3', 2 * DIM, DIM, 5, output)
        if BN_FLAG:
            output = lib.ops.batchnorm.Batchnorm('Generator.BN3', [0, 2, 3], output)
        output = tf.nn.relu(output)

        output = lib.ops.deconv2d.Deconv2D('Generator.5', DIM, 1, 5, output)
        output = tf.nn.sigmoid(output)                                            # (:187)
        return tf.reshape(output, [-1, OUTPUT_DIM]), None, None

    def Extractor(inputs):
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

    def _mlp_critic(z, k):
        output = tf.concat([z, k], 1)
        output = lib.ops.linear.Linear('Discriminator.HyperInput', DIM_LATENT + N_COMS, 512, output)
        output = LeakyReLU(output)
        output = tf.layers.dropout(output, rate=DR_RATE)
        output = lib.ops.linear.Linear('Discriminator.Hyper2', 512, 512, output)
        output = LeakyReLU(output)
        output = tf.layers.dropout(output, rate=DR_RATE)
        output = lib.ops.linear.Linear('Discriminator.Hyper3', 512, 512, output)
        output = LeakyReLU(output)
        output = tf.layers.dropout(output, rate=DR_RATE)
        output = lib.ops.linear.Linear('Discriminator.HyperOutput', 512, 1, output)
        return tf.reshape(output, [-1])

    def _conv_trunk(prefix, x):
        output = tf.reshape(x, [-1, 1, 28, 28])
        output = lib.ops.conv2d.Conv2D('Discriminator.%s1' % prefix, 1, DIM, 5, output, stride=2)
        output = LeakyReLU(output)
        output = tf.layers.dropout(output, rate=DR_RATE)
        output = lib.ops.conv2d.Conv2D('Discriminator.%s2' % prefix, DIM, 2 * DIM, 5, output, stride=2)
        output = LeakyReLU(output)
        output = tf.layers.dropout(output, rate=DR_RATE)
        output = lib.ops.conv2d.Conv2D('Discriminator.%s3' % prefix, 2 * DIM, 4 * DIM, 5, output, stride=2)
        output = LeakyReLU(output)
        output = tf.layers.dropout(output, rate=DR_RATE)
        return tf.reshape(output, [-1, 4 * 4 * 4 * DIM])

    if MODE in ['vegan', 'vegan-wgan-gp']:
        Discriminator = _mlp_critic                                              # (:242-258)
    elif MODE in ['local_ep', 'local_epce']:
        HyperDiscriminator = _mlp_critic                                         # (:262-278)

        def Discriminator(x, z):                                                 # (:280-303)
            output = _conv_trunk('', x)
            z_output = lib.ops.linear.Linear('Discriminator.z1', DIM_LATENT, 512, z)
            z_output = LeakyReLU(z_output)
            z_output = tf.layers.dropout(z_output, rate=DR_RATE)
            output = tf.concat([output, z_output], 1)
            output = lib.ops.linear.Linear('Discriminator.zx1', 4 * 4 * 4 * DIM + 512, 512, output)
            output = LeakyReLU(output)
            output = tf.layers.dropout(output, rate=DR_RATE)
            output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
            return tf.reshape(output, [-1])
    else:
        def Discriminator(x, z, k):                                              # (:309-336)
            output = _conv_trunk('x', x)
            zk_output = tf.concat([z, k], 1)
            zk_output = lib.ops.linear.Linear('Discriminator.zk1', DIM_LATENT + N_COMS, 512, zk_output)
            zk_output = LeakyReLU(zk_output)
            zk_output = tf.layers.dropout(zk_output, rate=DR_RATE)
            output = tf.concat([output, zk_output], 1)
            output = lib.ops.linear.Linear('Discriminator.zkx1', 4 * 4 * 4 * DIM + 512, 512, output)
            output = LeakyReLU(output)
            output = tf.layers.dropout(output, rate=DR_RATE)
            output = lib.ops.linear.Linear('Discriminator.Output', 512, 1, output)
            return tf.reshape(output, [-1])

    # ---- losses (:341-410) ----
    real_x = tf.placeholder(tf.float32, shape=[BATCH_SIZE, OUTPUT_DIM])      # MNIST feeds float [0,1] directly (:335)
    q_z, _, _ = Extractor(real_x)
    q_k_logits, q_k = HyperExtractor(q_z)
    q_k_probs = tf.nn.softmax(q_k_logits)
    if MODE_K == 'REINFORCE':
        q_k_prob_max = tf.reduce_max(q_k_probs, axis=1)
    rec_x, _, _ = Generator(q_z)
    hyper_p_z = tf.random_normal([BATCH_SIZE, DIM_LATENT])
    hyper_p_k_idx = prior_k.sample(BATCH_SIZE)
    hyper_p_k = tf.one_hot(indices=hyper_p_k_idx, depth=N_COMS)
    p_z = HyperGenerator(hyper_p_k, hyper_p_z)
    fake_x, _, _ = Generator(p_z)
    rec_z, _, _ = Extractor(fake_x)
    rec_q_k_logits, rec_q_k = HyperExtractor(rec_z)

    score_function = None
    if MODE == 'vegan':
        disc_fake = Discriminator(p_z, hyper_p_k)
        disc_real = Discriminator(q_z, q_k)
    elif MODE in ['local_ep', 'local_epce']:
        disc_fake, disc_real = [], []
        disc_fake.append(HyperDiscriminator(p_z, hyper_p_k))
        disc_real.append(HyperDiscriminator(q_z, q_k))
        disc_fake.append(Discriminator(fake_x, p_z))
        disc_real.append(Discriminator(real_x, q_z))
    else:
        disc_real = Discriminator(real_x, q_z, q_k)
        disc_fake = Discriminator(fake_x, p_z, hyper_p_k)
    if MODE_K == 'REINFORCE':                                                    # score-function estimator for the hard k
        import tflib.objs.discrete_variables
        critic_on_real = disc_real[0] if MODE in ['local_ep', 'local_epce'] else disc_real
        score_function = lib.objs.discrete_variables.score_function(critic_on_real, q_k_prob_max, CONTROL_VARIATE)

    gen_params = lib.params_with_name('Generator')
    ext_params = lib.params_with_name('Extractor')
    disc_params = lib.params_with_name('Discriminator')

    gi = lib.objs.gan_inference
    if MODE == 'ali':
        rec_penalty = None
        costs = gi.ali(disc_fake, disc_real, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, s_f=score_function)
    elif MODE == 'alice':
        rec_penalty = 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        costs = gi.alice(disc_fake, disc_real, rec_penalty, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1,
                         s_f=score_function)
    elif MODE == 'local_ep':
        rec_penalty = None
        costs = gi.local_ep(disc_fake, disc_real, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1, s_f=score_function)
    elif MODE == 'local_epce':
        rec_penalty = 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        costs = gi.local_epce(disc_fake, disc_real, rec_penalty, gen_params + ext_params, disc_params, lr=LR, beta1=BETA1,
                              s_f=score_function)
    elif MODE == 'vegan':
        rec_penalty = 1. * lib.utils.distance.distance(real_x, rec_x, DISTANCE_X)
        costs = gi.vegan(disc_fake, disc_real, rec_penalty, gen_params + ext_params, disc_params, LAMBDA, lr=LR, beta1=BETA1,
                         s_f=score_function)
    else:
        raise NotImplementedError(MODE)
    gen_cost, disc_cost, gen_train_op, disc_train_op = costs

    # ---- fixed-noise samples for visualisation (:413-419) ----
    np_fixed_noise = np.random.normal(size=(N_VIS, DIM_LATENT)).astype('float32')
    np_fixed_k = np.tile(np.eye(N_COMS, dtype=int), (N_VIS // N_COMS, 1))
    fixed_noise = HyperGenerator(tf.constant(np_fixed_k), tf.constant(np_fixed_noise))
    fixed_noise_samples, _, _ = Generator(fixed_noise)

    ns.__dict__.update(real_x=real_x, q_z=q_z, q_k=q_k, q_k_logits=q_k_logits, rec_x=rec_x,
                       hyper_p_z=hyper_p_z, hyper_p_k_idx=hyper_p_k_idx, hyper_p_k=hyper_p_k, p_z=p_z, fake_x=fake_x,
                       rec_z=rec_z, disc_fake=disc_fake, disc_real=disc_real, gen_params=gen_params, ext_params=ext_params,
                       disc_params=disc_params, rec_penalty=rec_penalty, gen_cost=gen_cost, disc_cost=disc_cost,
                       gen_train_op=gen_train_op, disc_train_op=disc_train_op, fixed_noise_samples=fixed_noise_samples,
                       np_fixed_noise=np_fixed_noise, np_fixed_k=np_fixed_k)
    return ns


def synthetic_batches(batch_size, output_dim=784, n=8, seed=0):
    """a ring of pre-generated uniform [0,1] float32 batches"""
    rs = np.random.RandomState(seed)
    ring = [rs.uniform(0, 1, size=(batch_size, output_dim)).astype('float32') for _ in range(n)]
    while True:
        for b in ring:
            yield b


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='local_ep')
    ap.add_argument('--iters', type=int, default=200000)
    ap.add_argument('--batch-size', type=int, default=50)
    ap.add_argument('--data-dir', default='/tmp/mnist.pkl.gz')
    ap.add_argument('--synthetic', action='store_true', help='train on uniform random images (no dataset needed)')
    ap.add_argument('--out', default=None)
    args = ap.parse_args(argv)

    MODE, BATCH_SIZE, ITERS = args.mode, args.batch_size, args.iters
    outf = args.out or os.path.join("result", "gmgan_inference_mnist.MODE-%s.N_COMS-30.%d" % (MODE, int(time.time())))
    os.makedirs(outf, exist_ok=True)
    logfile = os.path.join(outf, 'logfile.txt')
    lib.print_model_settings_to_file(dict(MODE=MODE, BATCH_SIZE=BATCH_SIZE, ITERS=ITERS), logfile)

    g = build_graph(MODE=MODE, BATCH_SIZE=BATCH_SIZE)
    if args.synthetic or not os.path.exists(args.data_dir):
        gen = synthetic_batches(BATCH_SIZE)
    else:
        import gzip
        import pickle
        with gzip.open(args.data_dir, 'rb') as f:                               # tflib/mnist.py:53-57 (mnist.pkl.gz)
            train_data, _, _ = pickle.load(f, encoding='latin1')
        images = train_data[0].astype('float32')

        def inf_train_gen():
            rs = np.random.RandomState(1)
            while True:
                rs.shuffle(images)
                for i in range(len(images) // BATCH_SIZE):
                    yield images[i * BATCH_SIZE:(i + 1) * BATCH_SIZE]
        gen = inf_train_gen()

    saver = tf.train.Saver()
    with tf.Session() as session:
        session.run(tf.global_variables_initializer())
        total_num = np.sum([np.prod(v.shape) for v in tf.trainable_variables()])
        print('\nTotol number of parameters', total_num)
        for iteration in range(ITERS):                                           # (:480-503)
            start_time = time.time()
            if iteration > 0:
                _data = next(gen)
                _gen_cost, _ = session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x: _data})
            for i in range(g.CRITIC_ITERS):
                _data = next(gen)
                _disc_cost, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x: _data})
            lib.plot.plot('train disc cost', _disc_cost)
            lib.plot.plot('time', time.time() - start_time)
            if (iteration < 5) or (iteration % 100 == 99):
                lib.plot.flush(outf, logfile)
            lib.plot.tick()
            if iteration % 5000 == 4999:
                samples = session.run(g.fixed_noise_samples)
                np.save(os.path.join(outf, '%d_samples_%s.npy' % (iteration, MODE)),
                        samples.reshape((-1, 28, 28)))
            if iteration == ITERS - 1:
                saver.save(session, os.path.join(outf, '{}_model_{}.ckpt'.format(iteration, MODE)))


if __name__ == '__main__':
    main()
