"""tflib.save_images — sample grids; counterpart of tflib/save_images.py:11-86 without scipy.misc / imageio (both absent from
modern environments): the grid is assembled with NumPy and written with Pillow.  Accepts [N, H*W] (square grey images),
[N, H, W] or BCHW [N, 3, H, W]; floats in [0,1] are scaled by 255.99 like the reference (:14-15)."""
import numpy as np


def large_image(X, size=None):
    X = np.asarray(X)
    if np.issubdtype(X.dtype, np.floating):
        X = (255.99 * X).astype('uint8')
    n = X.shape[0]
    if size is None:
        rows = int(np.sqrt(n))
        while n % rows != 0:
            rows -= 1
        nh, nw = rows, n // rows
    else:
        nh, nw = int(size[0]), int(size[1])
        assert nh * nw == n
    if X.ndim == 2:
        side = int(np.sqrt(X.shape[1]))
        X = X.reshape(n, side, side)
    if X.ndim == 4:
        X = X.transpose(0, 2, 3, 1)                      # BCHW -> BHWC
    h, w = X.shape[1:3]
    grid = X.reshape((nh, nw, h, w) + X.shape[3:])
    grid = grid.swapaxes(1, 2).reshape((nh * h, nw * w) + X.shape[3:])
    return grid.astype('uint8')


def save_images(X, save_path, size=None):
    from PIL import Image
    Image.fromarray(large_image(X, size)).save(save_path)


def save_gifs(x, save_path, size=None):
    """x [N, T, C, H, W] -> animated GIF of T grids"""
    from PIL import Image
    frames = [Image.fromarray(large_image(x[:, i], size=size)) for i in range(x.shape[1])]
    frames[0].save(save_path, save_all=True, append_images=frames[1:], duration=100, loop=0)
