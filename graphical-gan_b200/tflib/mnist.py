"""tflib.mnist — MNIST batches (mnist.pkl.gz: three (images float32 [N,784] in [0,1], targets) tuples); Python-3 counterpart
of tflib/mnist.py:8-64.  `load(batch_size, test_batch_size, n_labelled=None)` returns (train, dev, test) epoch callables that
yield (images [B,784] float32, targets [B]) — plus a `labelled` 0/1 vector when n_labelled is given, as the reference does
(:14-16,31-41).  Nothing is downloaded: a missing file raises."""
import gzip
import os
import pickle

import numpy as np

FILEPATH = '/tmp/mnist.pkl.gz'


def mnist_generator(data, batch_size, n_labelled, limit=None):
    images, targets = np.asarray(data[0], dtype=np.float32), np.asarray(data[1], dtype=np.int32)
    if limit is not None:
        images, targets = images[:limit], targets[:limit]
    labelled = None
    if n_labelled is not None:
        labelled = np.zeros(len(images), dtype='int32')
        labelled[:n_labelled] = 1

    def get_epoch():
        perm = np.random.permutation(len(images))
        lab = labelled[perm] if labelled is not None else None
        for i in range(len(images) // batch_size):
            idx = perm[i * batch_size:(i + 1) * batch_size]
            if lab is not None:
                yield images[idx], targets[idx], lab.copy()
            else:
                yield images[idx], targets[idx]
    return get_epoch


def load(batch_size, test_batch_size, n_labelled=None, filepath=FILEPATH):
    if not os.path.isfile(filepath):
        raise IOError("%s not found (no download here; use --synthetic)" % filepath)
    with gzip.open(filepath, 'rb') as f:
        train_data, dev_data, test_data = pickle.load(f, encoding='latin1')
    return (mnist_generator(train_data, batch_size, n_labelled), mnist_generator(dev_data, test_batch_size, n_labelled),
            mnist_generator(test_data, test_batch_size, n_labelled))
