"""GMGAN on SVHN — Python-3 port of the reference's gmgan_inference_svhn.py driving the B200 kernels.

The reference script is gmgan_inference_cifar10.py with three constants changed (diff of the two files): the data
loader (tflib/svhn.py instead of tflib/cifar10.py, :18,424), BN_FLAG = False (:70) and N_COMS = 50 (:72); the inception
score block is dropped (:422).  The networks, graph and training loop are therefore shared with the CIFAR-10 port and
only the constants live here.  MODEs: ali / alice / local_ep / local_epce / vegan (:33); the WGAN-GP variants of
BASELINE.json configs[2] live in gan_inference_svhn.py.
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import gmgan_inference_cifar10 as _base
from gmgan_inference_cifar10 import synthetic_batches, tf, lib

N_COMS = 50
BN_FLAG = False


def build_graph(MODE='local_ep', BATCH_SIZE=64, DIM=64, N_COMS=N_COMS, LR=2e-4, MODE_K='CONCRETE', N_VIS=None):
    return _base.build_graph(MODE=MODE, BATCH_SIZE=BATCH_SIZE, DIM=DIM, N_COMS=N_COMS, LR=LR, MODE_K=MODE_K, N_VIS=N_VIS,
                             BN_FLAG=BN_FLAG)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='local_ep')
    ap.add_argument('--iters', type=int, default=200000)
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--data-dir', default='./dataset/svhn')
    ap.add_argument('--synthetic', action='store_true', help='train on uniform random images (no dataset needed)')
    ap.add_argument('--out', default=None)
    args = ap.parse_args(argv)
    MODE, BATCH_SIZE, ITERS = args.mode, args.batch_size, args.iters
    outf = args.out or os.path.join("result", "gmgan_inference_svhn.MODE-%s.N_COMS-%d.%d" % (MODE, N_COMS, int(time.time())))
    os.makedirs(outf, exist_ok=True)
    logfile = os.path.join(outf, 'logfile.txt')
    lib.print_model_settings_to_file(dict(MODE=MODE, BATCH_SIZE=BATCH_SIZE, ITERS=ITERS, N_COMS=N_COMS, BN_FLAG=BN_FLAG), logfile)
    g = build_graph(MODE=MODE, BATCH_SIZE=BATCH_SIZE)
    train_mat = os.path.join(args.data_dir, 'train_32x32.mat')
    if args.synthetic or not os.path.exists(train_mat):
        gen = synthetic_batches(BATCH_SIZE)
    else:
        import scipy.io                                                          # tflib/svhn.py:32-49: X[32,32,3,N] -> [N,3072]
        X = scipy.io.loadmat(train_mat)['X'].transpose(3, 2, 0, 1).reshape(-1, 3072).astype('int32')

        def inf_train_gen():
            rs = np.random.RandomState(1)
            while True:
                rs.shuffle(X)
                for i in range(len(X) // BATCH_SIZE):
                    yield X[i * BATCH_SIZE:(i + 1) * BATCH_SIZE]
        gen = inf_train_gen()
    saver = tf.train.Saver()
    with tf.Session() as session:
        session.run(tf.global_variables_initializer())
        for iteration in range(ITERS):                                           # (:458-481)
            start_time = time.time()
            if iteration > 0:
                _gen_cost, _ = session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: next(gen)})
            for i in range(g.CRITIC_ITERS):
                _disc_cost, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: next(gen)})
            lib.plot.plot('train disc cost', _disc_cost)
            lib.plot.plot('time', time.time() - start_time)
            if (iteration < 5) or (iteration % 100 == 99):
                lib.plot.flush(outf, logfile)
            lib.plot.tick()
            if iteration == ITERS - 1:
                saver.save(session, os.path.join(outf, '{}_model_{}.ckpt'.format(iteration, MODE)))


if __name__ == '__main__':
    main()
