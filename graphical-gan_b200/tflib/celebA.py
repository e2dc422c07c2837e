"""tflib.celebA — CelebA 64x64 (celebA_64x64.npy: uint8 [N,3,64,64], produced by the reference's convert_to_numpy,
tflib/celebA.py:61-73); Python-3 counterpart of tflib/celebA.py:11-35.  `load(batch_size, data_dir, num_dev=5000)` shuffles
once, holds out the first num_dev images and returns (train_epoch, dev_epoch) callables yielding [B, 12288] uint8 batches."""
import os

import numpy as np

from ._batches import epoch_factory


def celeba_generator(batch_size, images):
    return epoch_factory((images,), batch_size, single=True)


def load(batch_size, data_dir, num_dev=5000):
    path = os.path.join(data_dir, 'celebA_64x64.npy')
    if not os.path.isfile(path):
        raise IOError("%s not found (use --synthetic)" % path)
    data = np.load(path)
    data = data.reshape(data.shape[0], -1)
    data = data[np.random.permutation(len(data))]
    return celeba_generator(batch_size, data[num_dev:]), celeba_generator(batch_size, data[:num_dev])
