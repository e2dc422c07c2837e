"""tflib.cifar10 — CIFAR-10 (python version) batches for the gmgan / gan_inference_cifar10 scripts; Python-3 counterpart of
tflib/cifar10.py:8-48.  On-disk format: the `cifar-10-batches-py` pickles with keys data (uint8 [10000, 3072], CHW order)
and labels.  `load(batch_size, data_dir)` returns (train_epoch, dev_epoch): callables that yield (images [B,3072] uint8,
labels [B]) for one shuffled epoch.  Nothing is downloaded (no network): a missing directory raises."""
import os
import pickle

import numpy as np

from ._batches import epoch_factory

TRAIN_FILES = ['data_batch_1', 'data_batch_2', 'data_batch_3', 'data_batch_4', 'data_batch_5']


def unpickle(path):
    with open(path, 'rb') as fo:
        d = pickle.load(fo, encoding='latin1')
    return np.asarray(d['data'], dtype=np.uint8), np.asarray(d['labels'], dtype=np.int32)


def _read(filenames, data_dir):
    parts = [unpickle(os.path.join(data_dir, f)) for f in filenames]
    return np.concatenate([p[0] for p in parts], axis=0), np.concatenate([p[1] for p in parts], axis=0)


def get_reconstruction_data(n_samples, data_dir):
    """fixed reconstruction samples for comparison (tflib/cifar10.py:14-19: seed 1234, shuffled test batch)"""
    np.random.seed(1234)
    data, _ = unpickle(os.path.join(data_dir, 'test_batch'))
    np.random.shuffle(data)
    return data[:n_samples]


def cifar_generator(filenames, batch_size, data_dir):
    images, labels = _read(filenames, data_dir)
    return epoch_factory((images, labels), batch_size)


def load(batch_size, data_dir):
    if not os.path.isdir(data_dir):
        raise IOError("CIFAR-10 directory %r not found (no download here; use --synthetic)" % data_dir)
    return cifar_generator(TRAIN_FILES, batch_size, data_dir), cifar_generator(['test_batch'], batch_size, data_dir)
