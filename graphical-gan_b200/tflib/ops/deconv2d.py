"""tflib.ops.deconv2d — drop-in for the reference's Deconv2D (tflib/ops/deconv2d.py:20-119): same name, arguments
and parameter names ('.Filters' (k,k,Cout,Cin), '.Biases' [Cout]); NCHW in, NCHW out at twice the resolution.

tf.nn.conv2d_transpose(stride 2, SAME) is the input-gradient of the mirrored 5x5 stride-2 conv, so the forward of
this layer runs on the same dgrad kernel (gg_conv2d_dgrad) as Conv2D's backward, with bias and the following
activation fused; the reference's two explicit layout transposes (deconv2d.py:91,116) are views here.
"""
import numpy as np
import tensorflow as tf

import tflib as lib
from gg import initializers as _init
from gg import layers as _L
from gg import ops as _O

_default_weightnorm = False
_weights_stdev = None


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Deconv2D(name, input_dim, output_dim, filter_size, inputs, he_init=True, weightnorm=None, biases=True, gain=1.,
             mask_type=None, stride=2, padding='SAME'):
    if mask_type is not None:
        raise Exception('Unsupported configuration')
    kshape = (filter_size, filter_size, output_dim, input_dim)
    fan_in = input_dim * filter_size ** 2 // (stride ** 2)      # Python-2 integer division in the reference (deconv2d.py:52)
    fan_out = output_dim * filter_size ** 2
    stdev = _weights_stdev if _weights_stdev is not None else _init.fan_stdev(fan_in, fan_out, he_init)
    filter_values = _init.uniform(stdev, kshape) * gain
    filters = lib.param(name + '.Filters', filter_values)

    if weightnorm is None:
        weightnorm = _default_weightnorm
    if weightnorm:
        target_norms = lib.param(name + '.g', np.sqrt(np.sum(np.square(filter_values), axis=(0, 1, 3))))
        norms = tf.sqrt(tf.reduce_sum(tf.square(filters), reduction_indices=[0, 1, 3]))
        filters = filters * tf.expand_dims(target_norms / norms, 1)

    x = _O.to_nhwc(inputs)
    B, H, W, _ = x.shape
    if padding == 'VALID':   # the reference uses the height for both extents (deconv2d.py:99)
        out_shape = [B, stride * (H - 1) + filter_size, stride * (H - 1) + filter_size, output_dim]
    else:
        out_shape = [B, stride * H, stride * W, output_dim]
    bias = lib.param(name + '.Biases', np.zeros(output_dim, dtype='float32')) if biases else None
    return _O.to_nchw(_L.conv2d_transpose_nhwc(x, filters, out_shape, stride, padding, bias=bias))
