"""NumPy weight initialisers shared by tflib.ops.{linear,conv2d,deconv2d}.

Draw-for-draw compatible with the reference so that replaying a numpy seed reproduces its initial weights:
one `np.random.uniform(-s*sqrt(3), s*sqrt(3), size)` float64 draw cast to float32 per layer CALL (the reference
draws before it looks the name up in the registry, so repeated calls of a shared layer consume random numbers
too — tflib/ops/conv2d.py:74-88, deconv2d.py:60-74, linear.py:39-111).
"""
import numpy as np

SQRT3 = np.sqrt(3)


def uniform(stdev, size):
    return np.random.uniform(low=-stdev * SQRT3, high=stdev * SQRT3, size=size).astype('float32')


def fan_stdev(fan_in, fan_out, he_init):
    """He-style sqrt(4/(fan_in+fan_out)) or Glorot sqrt(2/(fan_in+fan_out))  (conv2d.py:62-72, deconv2d.py:50-57)"""
    return np.sqrt((4. if he_init else 2.) / (fan_in + fan_out))


def linear_weights(input_dim, output_dim, initialization, override_stdev):
    """linear.py:39-106.  `None` takes the Glorot branch (the later `orthogonal` test for None is unreachable)."""
    def u(stdev):
        return uniform(override_stdev if override_stdev is not None else stdev, (input_dim, output_dim))
    if initialization == 'lecun':
        return u(np.sqrt(1. / input_dim))
    if initialization == 'glorot' or initialization is None:
        return u(np.sqrt(2. / (input_dim + output_dim)))
    if initialization == 'he':
        return u(np.sqrt(2. / input_dim))
    if initialization == 'glorot_he':
        return u(np.sqrt(4. / (input_dim + output_dim)))
    if initialization == 'orthogonal':
        a = np.random.normal(0.0, 1.0, (input_dim, output_dim))
        uu, _, vv = np.linalg.svd(a, full_matrices=False)
        q = uu if uu.shape == (input_dim, output_dim) else vv
        return q.reshape((input_dim, output_dim)).astype('float32')
    if isinstance(initialization, (tuple, list)) and initialization[0] == 'uniform':
        r = initialization[1]
        return np.random.uniform(low=-r, high=r, size=(input_dim, output_dim)).astype('float32')
    raise Exception('Invalid initialization!')


def pixelcnn_mask(mask_type, filter_size, input_dim, output_dim):
    """PixelCNN 'a'/'b' filter masks (conv2d.py:29-52); unused by every training script, kept for the API."""
    kind, n_ch = mask_type
    mask = np.ones((filter_size, filter_size, input_dim, output_dim), dtype='float32')
    c = filter_size // 2
    mask[c + 1:] = 0.
    mask[c, c + 1:] = 0.
    for i in range(n_ch):
        for j in range(n_ch):
            if (kind == 'a' and i >= j) or (kind == 'b' and i > j):
                mask[c, c, i::n_ch, j::n_ch] = 0.
    return mask
