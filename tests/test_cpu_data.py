"""CPU: the dataset modules behind the reference's tflib.{cifar10,mnist,svhn,celebA,save_images} surface (SURVEY.md §8(f)
N3/N4), on small synthetic files written in the datasets' own on-disk formats."""
import gzip
import os
import pickle

import numpy as np
import pytest


def test_cifar10_pickles_round_trip(tmp_path):
    import tflib.cifar10 as C
    rs = np.random.RandomState(0)
    ref = {}
    for name in C.TRAIN_FILES + ['test_batch']:
        d = {'data': rs.randint(0, 256, size=(20, 3072)).astype(np.uint8), 'labels': list(rs.randint(0, 10, size=20))}
        ref[name] = d
        with open(tmp_path / name, 'wb') as f:
            pickle.dump(d, f, protocol=2)
    train, dev = C.load(8, str(tmp_path))
    batches = list(train())
    assert len(batches) == 100 // 8 and batches[0][0].shape == (8, 3072) and batches[0][0].dtype == np.uint8
    allx = np.concatenate([ref[n]['data'] for n in C.TRAIN_FILES]); ally = np.concatenate([ref[n]['labels'] for n in C.TRAIN_FILES])
    for x, y in batches[:3]:                                   # images and labels stay aligned through the shuffle
        for xi, yi in zip(x, y):
            j = np.where((allx == xi).all(1))[0][0]
            assert ally[j] == yi
    assert sum(len(b[0]) for b in dev()) == 16
    assert C.get_reconstruction_data(5, str(tmp_path)).shape == (5, 3072)
    with pytest.raises(IOError):
        C.load(8, str(tmp_path / "missing"))


def test_mnist_svhn_celeba_formats(tmp_path):
    import scipy.io
    import tflib.celebA as A
    import tflib.mnist as M
    import tflib.svhn as S
    rs = np.random.RandomState(1)
    sets = tuple((rs.uniform(0, 1, size=(n, 784)).astype(np.float32), rs.randint(0, 10, size=n)) for n in (30, 10, 10))
    path = str(tmp_path / "mnist.pkl.gz")
    with gzip.open(path, 'wb') as f:
        pickle.dump(sets, f, protocol=2)
    train, dev, test = M.load(10, 5, filepath=path)
    b = list(train())
    assert len(b) == 3 and b[0][0].shape == (10, 784) and b[0][0].dtype == np.float32 and len(list(dev())) == 2
    tl, _, _ = M.load(10, 5, n_labelled=7, filepath=path)
    assert next(tl())[2].sum() == 7
    X = rs.randint(0, 256, size=(32, 32, 3, 12)).astype(np.uint8)
    y = rs.randint(1, 11, size=(12, 1))
    for name in ('train_32x32.mat', 'test_32x32.mat'):
        scipy.io.savemat(str(tmp_path / name), {'X': X, 'y': y})
    tr, te = S.load(4, str(tmp_path))
    xb, yb = next(tr())
    assert xb.shape == (4, 3072) and yb.max() <= 9
    j = int(np.where((np.transpose(X, [3, 2, 0, 1]).reshape(12, -1) == xb[0]).all(1))[0][0])   # CHW flattening
    assert (y.flatten()[j] % 10) == yb[0]
    np.save(str(tmp_path / 'celebA_64x64.npy'), rs.randint(0, 256, size=(20, 3, 64, 64)).astype(np.uint8))
    tr, dv = A.load(4, str(tmp_path), num_dev=8)
    assert next(tr()).shape == (4, 12288) and sum(len(b) for b in dv()) == 8


def test_save_images_grid(tmp_path):
    import tflib.save_images as V
    from PIL import Image
    x = np.zeros((6, 3, 4, 5), np.float32)
    x[4, 1] = 1.0                                              # image 4 -> grid row 1, column 1 for size (2, 3)
    g = V.large_image(x, size=(2, 3))
    assert g.shape == (8, 15, 3) and g.dtype == np.uint8
    assert (g[4:8, 5:10, 1] == 255).all() and g[:, :, 0].max() == 0 and g[0:4].max() == 0
    V.save_images(x, str(tmp_path / "grid.png"), size=(2, 3))
    assert Image.open(str(tmp_path / "grid.png")).size == (15, 8)
    assert V.large_image(np.zeros((9, 784), np.float32)).shape == (84, 84)


def test_moving_mnist_sequences(tmp_path):
    import tflib.simple_moving_mnist as MM
    np.random.seed(3)
    sy, sx = MM.GetRandomTrajectory(0.1, 40, 50, 64, 28)
    assert sy.shape == (40, 50) and sy.min() >= 0 and sy.max() <= 36 and sx.min() >= 0 and sx.max() <= 36
    step = np.hypot(np.diff(sy, axis=0), np.diff(sx, axis=0))
    assert step.max() <= 0.1 * 36 * np.sqrt(2) + 2                 # unit speed, step 0.1 of the 36-pixel canvas
    assert (np.abs(np.diff(np.sign(np.diff(sx, axis=0)), axis=0)) > 0).any()   # somebody bounced
    digit = np.zeros((1, 28, 28), np.float32); digit[0, 3:7, 10:12] = 0.8
    v = MM.render(digit, np.array([[5], [30]]), np.array([[2], [36]]))
    assert v.shape == (1, 2, 64, 64) and v[0, 0, 8:12, 12:14].min() == np.float32(0.8) and v[0, 0].sum() == np.float32(0.8) * 8
    assert v[0, 1, 33:37, 46:48].min() == np.float32(0.8)
    rs = np.random.RandomState(1)
    sets = tuple((rs.uniform(0, 1, size=(n, 784)).astype(np.float32), rs.randint(0, 10, size=n)) for n in (12, 4, 6))
    path = str(tmp_path / "mnist.pkl.gz")
    with gzip.open(path, 'wb') as f:
        pickle.dump(sets, f, protocol=2)
    train, test = MM.load_video(5, 4, filepath=path)
    vids, labels = next(train())
    assert vids.shape == (4, 5, 4096) and vids.dtype == np.float32 and labels.shape == (4,) and len(list(train())) == 4
    assert len(list(test())) == 1
