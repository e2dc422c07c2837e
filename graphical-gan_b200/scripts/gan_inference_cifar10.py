"""ALI / ALICE / VEGAN / WALI(-GP) on CIFAR-10 — Python-3 port of the reference's gan_inference_cifar10.py on the B200 kernels.

The reference script is gan_inference_svhn.py with the CIFAR-10 loader (:18,394), an inception-score block (:381-392,
needs a frozen TF graph: not ported) and ONE model constant changed (diff of the two files): BN_FLAG = True for the
non-vegan modes (:72-77) — batch norm in the Generator and the Extractor, none in the (x, z) critic (:226-250), so the
WGAN-GP double backward of MODE='wali-gp' never meets a batch norm.  Networks, graph and objectives are shared with the
SVHN port; only the constants live here.
"""
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import gan_inference_svhn as _base
from gan_inference_svhn import SUPPORTED, tf, lib


def build_graph(MODE='ali', BATCH_SIZE=64, DIM=64, LR=2e-4, Z_SAMPLES=100):
    bn = MODE not in ('vegan', 'vegan-wgan-gp', 'vegan-kl', 'vegan-jsd', 'vegan-ikl')        # :72-77
    return _base.build_graph(MODE=MODE, BATCH_SIZE=BATCH_SIZE, DIM=DIM, LR=LR, BN_FLAG=bn, Z_SAMPLES=Z_SAMPLES)


def main(argv=None):
    import argparse
    from gmgan_inference_cifar10 import synthetic_batches
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='ali', choices=SUPPORTED)
    ap.add_argument('--iters', type=int, default=200000)
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--data-dir', default='./dataset/cifar10/cifar-10-batches-py')
    ap.add_argument('--synthetic', action='store_true')
    args = ap.parse_args(argv)
    g = build_graph(MODE=args.mode, BATCH_SIZE=args.batch_size)
    if args.synthetic or not os.path.isdir(args.data_dir):
        gen = synthetic_batches(args.batch_size)
    else:
        import tflib.cifar10
        train_gen, _ = lib.cifar10.load(args.batch_size, data_dir=args.data_dir)

        def inf_train_gen():
            while True:
                for images, _ in train_gen():
                    yield images.astype('int32')
        gen = inf_train_gen()
    with tf.Session() as session:
        session.run(tf.global_variables_initializer())
        for iteration in range(args.iters):                                                  # :440-470
            start_time = time.time()
            if iteration > 0:
                gc, _ = session.run([g.gen_cost, g.gen_train_op], feed_dict={g.real_x_int: next(gen)})
            for i in range(g.CRITIC_ITERS):
                dc, _ = session.run([g.disc_cost, g.disc_train_op], feed_dict={g.real_x_int: next(gen)})
                if g.clip_disc_weights is not None:
                    session.run(g.clip_disc_weights)
            if g.CRITIC_ITERS == 0:                                                          # no-discriminator modes
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
