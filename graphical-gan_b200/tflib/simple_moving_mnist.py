"""tflib.simple_moving_mnist — bouncing-digit sequences for the SSGAN script; Python-3, vectorised counterpart of
tflib/simple_moving_mnist.py:9-153.  One MNIST digit (28x28) moves inside a 64x64 frame: position uniform in the unit box,
unit-speed direction uniform, step 0.1 per frame, velocity component flipped (and position clamped) at the walls (:9-48);
frames are the pixel-wise maximum of canvas and digit (:50-52).  `load_video(seq_length, batch_size, cla=None)` returns
(train_epoch, test_epoch) callables yielding (videos [B, LEN, 4096] float32, labels [B]); train = MNIST train + dev (:99-100).
Nothing is downloaded: a missing /tmp/mnist.pkl.gz raises."""
import gzip
import os
import pickle

import numpy as np

IMAGE_SIZE, DIGIT_SIZE, STEP_LENGTH = 64, 28, 0.1


def GetRandomTrajectory(step_length, seq_length, batch_size, image_size, digit_size):
    """top-left corners [seq_length, batch_size] (int32) of a digit bouncing inside the frame"""
    canvas = image_size - digit_size
    y, x = np.random.rand(batch_size), np.random.rand(batch_size)
    theta = np.random.rand(batch_size) * 2 * np.pi
    v_y, v_x = np.sin(theta), np.cos(theta)
    start_y, start_x = np.zeros((seq_length, batch_size)), np.zeros((seq_length, batch_size))
    for i in range(seq_length):
        y, x = y + v_y * step_length, x + v_x * step_length
        for pos, vel in ((x, v_x), (y, v_y)):                      # bounce: clamp to the wall and reverse that component
            low, high = pos <= 0, pos >= 1.0
            pos[low], pos[high] = 0.0, 1.0
            vel[low | high] *= -1
        start_y[i], start_x[i] = y, x
    return (canvas * start_y).astype(np.int32), (canvas * start_x).astype(np.int32)


def render(images, start_y, start_x, image_size=IMAGE_SIZE):
    """images [N, 28, 28] + corners [LEN, N] -> videos [N, LEN, image_size, image_size] (maximum overlap)"""
    n, d = images.shape[0], images.shape[1]
    seq = start_y.shape[0]
    data = np.zeros((n, seq, image_size, image_size), dtype=np.float32)
    for j in range(n):
        for i in range(seq):
            top, left = start_y[i, j], start_x[i, j]
            view = data[j, i, top:top + d, left:left + d]
            np.maximum(view, images[j], out=view)
    return data


def moving_mnist_generator_video(data_all, seq_length, batch_size):
    images = np.asarray(data_all[0], dtype=np.float32).reshape([-1, DIGIT_SIZE, DIGIT_SIZE])
    labels = np.asarray(data_all[1])

    def get_epoch():
        perm = np.random.permutation(len(images))
        start_y, start_x = GetRandomTrajectory(STEP_LENGTH, seq_length, len(images), IMAGE_SIZE, DIGIT_SIZE)
        for ind in range(len(images) // batch_size):
            idx = perm[ind * batch_size:(ind + 1) * batch_size]
            video = render(images[idx], start_y[:, idx], start_x[:, idx])
            yield video.reshape(batch_size, seq_length, IMAGE_SIZE * IMAGE_SIZE), labels[idx]
    return get_epoch


def _mnist(filepath):
    if not os.path.isfile(filepath):
        raise IOError("%s not found (no download here; the SSGAN script trains on synthetic sequences by default)" % filepath)
    with gzip.open(filepath, 'rb') as f:
        return pickle.load(f, encoding='latin1')


def load_video(seq_length, batch_size, cla=None, filepath='/tmp/mnist.pkl.gz'):
    train_data, dev_data, test_data = _mnist(filepath)
    x = np.concatenate([train_data[0], dev_data[0]], axis=0)
    y = np.concatenate([train_data[1], dev_data[1]], axis=0)
    tx, ty = np.asarray(test_data[0]), np.asarray(test_data[1])
    if cla is not None:
        x, y, tx, ty = x[y == cla], y[y == cla], tx[ty == cla], ty[ty == cla]
    return moving_mnist_generator_video((x, y), seq_length, batch_size), moving_mnist_generator_video((tx, ty), seq_length, batch_size)
