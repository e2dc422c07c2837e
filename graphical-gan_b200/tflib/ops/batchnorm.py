"""tflib.ops.batchnorm — drop-in for tflib/ops/batchnorm.py:6-87.

Every reference call site passes is_training=None, i.e. batch statistics at train AND sample time; the moving
averages are created (they are parameters returned by params_with_name, batchnorm.py:26-27) but never read.
Both branches — fused [0,2,3] / [0,2] (batchnorm.py:29-30) and the generic moments path for axes [0]
(batchnorm.py:77-84) — run on the same two sm_100a kernels (column statistics + fused normalise/affine/activation)
over a channels-last [rows, C] view; under data parallelism the statistics are all-reduced (SyncBN) so the result
equals the un-sharded reference.
"""
import numpy as np
import tensorflow as tf

import tflib as lib
from gg import ops as _O


def Batchnorm(name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True):
    if ((axes == [0, 2, 3]) or (axes == [0, 2])) and fused is True:
        if axes == [0, 2]:
            inputs = tf.expand_dims(inputs, 3)
        C = inputs.get_shape()[1]
        offset = lib.param(name + '.offset', np.zeros(C, dtype='float32'))
        scale = lib.param(name + '.scale', np.ones(C, dtype='float32'))
        lib.param(name + '.moving_mean', np.zeros(C, dtype='float32'), trainable=False)
        lib.param(name + '.moving_variance', np.ones(C, dtype='float32'), trainable=False)
        if is_training is not None:
            raise NotImplementedError("Batchnorm(is_training=<tensor>): the inference/moving-average branch "
                                      "(batchnorm.py:32-68) is never taken by the reference's scripts")
        # tf.nn.fused_batch_norm(inputs, scale, offset, epsilon=1e-5, data_format='NCHW')
        outputs = _O.to_nchw(_O.batchnorm(_O.to_nhwc(inputs), scale, offset, 1e-5))
        if axes == [0, 2]:
            return outputs[:, :, :, 0]
        return outputs
    else:
        nd = inputs.get_shape().ndims
        shape = [1 if a in axes else s for a, s in enumerate(inputs.get_shape())]
        if 0 not in axes:
            print("WARNING ({}): didn't find 0 in axes, but not using separate BN params for each item in batch".format(name))
            shape[0] = 1
        offset = lib.param(name + '.offset', np.zeros(shape, dtype='float32'))
        scale = lib.param(name + '.scale', np.ones(shape, dtype='float32'))
        if nd == 2 and list(axes) == [0]:
            # moments over the batch + batch_normalization(eps=1e-5): the fused [rows, C] kernel
            return _O.batchnorm(inputs, scale, offset, 1e-5)
        if nd > 2 and list(axes) == list(range(nd - 1)):
            # channels-last statistics over every other axis (Batchnorm(..., [0,1,2,3], ndhwc) in the SSGAN 3dcnn critic,
            # ssgan_inference_moving_mnist.py:372): the same kernel on the [rows, C] view
            C = int(inputs.get_shape()[-1])
            flat = _O.batchnorm(tf.reshape(inputs, [-1, C]), tf.reshape(scale, [C]), tf.reshape(offset, [C]), 1e-5)
            return tf.reshape(flat, [int(d) for d in inputs.get_shape()])
        mean, var = tf.nn.moments(inputs, axes, keep_dims=True)
        return tf.nn.batch_normalization(inputs, mean, var, offset, scale, 1e-5)
