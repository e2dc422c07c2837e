"""tflib.svhn — SVHN cropped digits ({train,test}_32x32.mat: X uint8 [32,32,3,N], y [N,1] with label 10 meaning 0);
Python-3 counterpart of tflib/svhn.py:9-49.  Images are returned flattened in CHW order ([N, 3072] uint8: X transposed
[3,2,0,1], :42-45), the layout every script's `tf.reshape(x, [-1, 3, 32, 32])` expects.  Nothing is downloaded."""
import os

import numpy as np

from ._batches import epoch_factory


def _read(path):
    from scipy.io import loadmat
    d = loadmat(path)
    y = d['y'].flatten().astype(np.int32)
    y[y == 10] = 0
    x = np.transpose(d['X'], [3, 2, 0, 1]).reshape([-1, 32 * 32 * 3])
    return x, y


def svhn_generator(data, batch_size):
    return epoch_factory(data, batch_size)


def load(batch_size, data_dir):
    train, test = os.path.join(data_dir, 'train_32x32.mat'), os.path.join(data_dir, 'test_32x32.mat')
    if not (os.path.isfile(train) and os.path.isfile(test)):
        raise IOError("SVHN .mat files not found under %r (no download here; use --synthetic)" % data_dir)
    return svhn_generator(_read(train), batch_size), svhn_generator(_read(test), batch_size)
