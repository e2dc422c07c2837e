"""Convolution geometry (TensorFlow SAME / VALID rules) and the conv node builders used by tflib.ops.* and tf.nn.*.

tf.nn.conv2d(padding='SAME', stride s): out = ceil(in/s), pad_total = max((out-1)*s + k - in, 0),
pad_before = pad_total // 2 (the remainder goes AFTER) — so a 5x5 stride-2 conv on an even input pads (1,2), not the
symmetric (2,2) of most frameworks (call site tflib/ops/conv2d.py:106-112).
tf.nn.conv2d_transpose is defined as the input-gradient of that conv (tflib/ops/deconv2d.py:101-107).
"""
from . import ops as O


def same_padding(in_size, k, stride):
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2


def conv_geometry(B, H, W, Ci, Co, k, stride, padding):
    if padding == "SAME":
        Ho, pt = same_padding(H, k, stride)
        Wo, pl = same_padding(W, k, stride)
    elif padding == "VALID":
        Ho, Wo, pt, pl = (H - k) // stride + 1, (W - k) // stride + 1, 0, 0
    else:
        raise ValueError("padding must be 'SAME' or 'VALID', got %r" % (padding,))
    return dict(B=B, H=H, W=W, Ci=Ci, Co=Co, k=k, stride=stride, pad_t=pt, pad_l=pl, Ho=Ho, Wo=Wo)


def conv2d_nhwc(x, filters, stride, padding, bias=None):
    B, H, W, Ci = x.shape
    k, k2, Ci2, Co = filters.shape
    if k != k2 or Ci != Ci2:
        raise ValueError("filter %s does not match input %s" % (tuple(filters.shape), tuple(x.shape)))
    return O.conv("fwd", x, filters, conv_geometry(B, H, W, Ci, Co, k, stride, padding), bias)


def conv3d_ndhwc(x, filters, stride_len, stride, padding):
    """tf.nn.conv3d(x [N,L,H,W,Ci], filters [fl,k,k,Ci,Co], strides [1,stride_len,stride,stride,1], 'SAME', 'NDHWC')
    (tflib/ops/conv3d.py:33-39) as ONE 2-D convolution: the fl depth taps of every output depth index are laid side by side
    along the channel axis (x' [Lo*N, H, W, fl*Ci], zero slices where TF's SAME padding reaches outside the clip) and the
    filter's depth taps are concatenated the same way (w' [k, k, fl*Ci, Co]).  Same FLOPs as the 3-D form, and every piece is
    a graph op with a gradient rule, so first- and second-order gradients come with it."""
    N, L, H, W, Ci = x.shape
    fl, k, k2, Ci2, Co = filters.shape
    if k != k2 or Ci != Ci2:
        raise ValueError("filter %s does not match input %s" % (tuple(filters.shape), tuple(x.shape)))
    if padding != 'SAME':
        raise NotImplementedError("conv3d padding %r (the reference only uses 'SAME', conv3d.py:37)" % (padding,))
    Lo = -(-L // stride_len)
    pad_total = max((Lo - 1) * stride_len + fl - L, 0)
    before = pad_total // 2                                   # TF SAME: the smaller half in front
    xp = O.pad_axis(x, 1, before, L + pad_total) if pad_total else x
    frames = []
    for lo in range(Lo):
        taps = [O.getitem(xp, (slice(None), lo * stride_len + t)) for t in range(fl)]          # fl x [N, H, W, Ci]
        frames.append(O.concat(taps, 3))
    xcat = O.concat(frames, 0)                                                                   # [Lo*N, H, W, fl*Ci], lo-major
    wcat = O.concat([O.getitem(filters, t) for t in range(fl)], 2)                             # [k, k, fl*Ci, Co]
    y = conv2d_nhwc(xcat, wcat, stride, 'SAME')                                                  # [Lo*N, Ho, Wo, Co]
    _, Ho, Wo, _ = y.shape
    y = O.transpose(O.reshape(y, [Lo, N, Ho * Wo * Co]), (1, 0, 2))
    return O.reshape(y, [N, Lo, Ho, Wo, Co])


def conv2d_nchw(x, filters, stride, padding, bias=None):
    return O.to_nchw(conv2d_nhwc(O.to_nhwc(x), filters, stride, padding, bias))


def conv2d_transpose_nhwc(x, filters, output_shape, stride, padding, bias=None):
    """x [B,Hin,Win,Cin]; filters (k,k,Cout,Cin); output_shape [B,Hout,Wout,Cout]: the dgrad of the conv
    Cout -> Cin over an Hout x Wout image."""
    B, Hin, Win, Cin = x.shape
    k, k2, Cout, Cin2 = filters.shape
    if k != k2 or Cin != Cin2:
        raise ValueError("filter %s does not match input %s" % (tuple(filters.shape), tuple(x.shape)))
    Hout, Wout = int(output_shape[1]), int(output_shape[2])
    geom = conv_geometry(B, Hout, Wout, Cout, Cin, k, stride, padding)
    if (geom["Ho"], geom["Wo"]) != (Hin, Win):
        raise ValueError("conv2d_transpose: output_shape %s is not consistent with input %s" % (output_shape, tuple(x.shape)))
    return O.conv("dgrad", x, filters, geom, bias)
