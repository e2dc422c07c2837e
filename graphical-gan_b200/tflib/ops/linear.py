"""tflib.ops.linear — drop-in for the reference's Linear (tflib/ops/linear.py:24-148): parameters '.W' [in,out],
'.b' [out], optional '.g'.  tf.matmul + tf.nn.bias_add (+ the script's next activation) are one GEMM launch with a
fused epilogue (gg_gemm); M = batch = 64 makes these layers weight-bandwidth bound, see DESIGN.md.
"""
import numpy as np
import tensorflow as tf

import tflib as lib
from gg import initializers as _init
from gg import ops as _O

_default_weightnorm = False
_weights_stdev = None


def enable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = True


def disable_default_weightnorm():
    global _default_weightnorm
    _default_weightnorm = False


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Linear(name, input_dim, output_dim, inputs, biases=True, initialization=None, weightnorm=None, gain=1.):
    """initialization: None, `lecun`, 'glorot', `he`, 'glorot_he', `orthogonal`, `("uniform", range)`"""
    weight_values = _init.linear_weights(input_dim, output_dim, initialization, _weights_stdev) * gain
    weight = lib.param(name + '.W', weight_values)

    if weightnorm is None:
        weightnorm = _default_weightnorm
    if weightnorm:
        target_norms = lib.param(name + '.g', np.sqrt(np.sum(np.square(weight_values), axis=0)))
        norms = tf.sqrt(tf.reduce_sum(tf.square(weight), reduction_indices=[0]))
        weight = weight * (target_norms / norms)

    bias = lib.param(name + '.b', np.zeros((output_dim,), dtype='float32')) if biases else None
    lead = list(inputs.get_shape())[:-1]
    x2d = inputs if len(lead) == 1 else tf.reshape(inputs, [-1, input_dim])
    result = _O.matmul(x2d, weight, bias=bias)
    return result if len(lead) == 1 else tf.reshape(result, lead + [output_dim])
