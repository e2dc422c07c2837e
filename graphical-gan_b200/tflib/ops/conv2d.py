"""tflib.ops.conv2d — drop-in for the reference's Conv2D (tflib/ops/conv2d.py:20-123): same name, arguments,
parameter names ('.Filters' (k,k,Cin,Cout), '.Biases' [Cout], optional '.g') and NCHW in/out convention.

tf.nn.conv2d(NCHW) + tf.nn.bias_add become ONE sm_100a kernel (gg_conv2d_fwd), and the activation / batch norm
the script applies next is folded into it by the graph builder; between image layers the data stays NHWC.
"""
import numpy as np
import tensorflow as tf

import tflib as lib
from gg import initializers as _init
from gg import layers as _L

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


def Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=True, mask_type=None, stride=1, weightnorm=None,
           biases=True, gain=1., padding='SAME'):
    """inputs / returns: (batch size, num channels, height, width); mask_type: None or ('a'|'b', n_channels)."""
    kshape = (filter_size, filter_size, input_dim, output_dim)
    fan_in = input_dim * filter_size ** 2
    fan_out = output_dim * filter_size ** 2 // (stride ** 2)     # Python-2 integer division in the reference (conv2d.py:63)
    if mask_type is not None:            # "only approximately correct" (conv2d.py:65-67)
        fan_in, fan_out = fan_in / 2., fan_out / 2.
    stdev = _weights_stdev if _weights_stdev is not None else _init.fan_stdev(fan_in, fan_out, he_init)
    filter_values = _init.uniform(stdev, kshape) * gain          # drawn on every call, before the registry lookup
    filters = lib.param(name + '.Filters', filter_values)

    if weightnorm is None:
        weightnorm = _default_weightnorm
    if weightnorm:
        target_norms = lib.param(name + '.g', np.sqrt(np.sum(np.square(filter_values), axis=(0, 1, 2))))
        norms = tf.sqrt(tf.reduce_sum(tf.square(filters), reduction_indices=[0, 1, 2]))
        filters = filters * (target_norms / norms)
    if mask_type is not None:
        filters = filters * tf.constant(_init.pixelcnn_mask(mask_type, filter_size, input_dim, output_dim))

    bias = lib.param(name + '.Biases', np.zeros(output_dim, dtype='float32')) if biases else None
    return _L.conv2d_nchw(inputs, filters, stride, padding, bias=bias)
