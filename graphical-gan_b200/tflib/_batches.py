"""Shared epoch iterator of the dataset modules: the reference shuffles images and labels in place with the same NumPy RNG
state at the start of every epoch and yields `len // batch_size` full batches (tflib/cifar10.py:32-41, svhn.py:19-30,
mnist.py:24-47, celebA.py:11-19).  Here one permutation is drawn from the same global NumPy RNG and applied to every
array, which keeps images and labels aligned without the save / restore of the RNG state."""
import numpy as np


def epoch_factory(arrays, batch_size, single=False):
    arrays = [np.asarray(a) for a in arrays]
    n = len(arrays[0])

    def get_epoch():
        perm = np.random.permutation(n)
        for i in range(n // batch_size):
            idx = perm[i * batch_size:(i + 1) * batch_size]
            out = tuple(a[idx] for a in arrays)
            yield out[0] if single else out
    return get_epoch
