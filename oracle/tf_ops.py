"""CPU oracle — restatement of the TensorFlow-1.x arithmetic behind the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under graphical-gan_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.

PARITY UNPINNED: the reference (zhenxuan00/graphical-gan) ships no tests, golden vectors or fixtures,
and its arithmetic lives in TensorFlow 1.x (>=1.4, <2.0; version not pinned by the reference), which is
not installed here and cannot be (no network; the scripts are Python 2).  This file therefore restates
the *documented* semantics of the TF ops at the reference's call sites in PyTorch-CPU (fp32 by default,
fp64 on request) and is itself validated only by internal identities (tests/test_oracle.py): SAME padding
arithmetic against an explicit loop implementation, conv_transpose == autograd input-gradient of conv,
batch-norm against its definition, Adam against a scalar hand computation.

Every function cites the reference call site (file:line under /root/reference) it follows.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# padding arithmetic of tf.nn.conv2d(padding='SAME')        [tflib/ops/conv2d.py:106-112]
# ----------------------------------------------------------------------------------------------
def same_padding(in_size, k, stride):
    """TF SAME: out = ceil(in/stride); pad_total = max((out-1)*stride + k - in, 0); before = total//2."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def conv_geometry(H, W, k, stride, padding):
    if padding == 'SAME':
        Ho, pt, pb = same_padding(H, k, stride)
        Wo, pl, pr = same_padding(W, k, stride)
    elif padding == 'VALID':
        Ho, Wo = (H - k) // stride + 1, (W - k) // stride + 1
        pt = pb = pl = pr = 0
    else:
        raise ValueError(padding)
    return Ho, Wo, pt, pb, pl, pr


def conv2d(x, filters, stride=1, padding='SAME', bias=None):
    """tf.nn.conv2d(NCHW) + tf.nn.bias_add  — tflib/ops/conv2d.py:106-120.

    x [B,Cin,H,W]; filters HWIO (k,k,Cin,Cout) as created at conv2d.py:75-83; cross-correlation (no flip).
    """
    k = filters.shape[0]
    _, _, H, W = x.shape
    _, _, pt, pb, pl, pr = conv_geometry(H, W, k, stride, padding)
    xp = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(xp, filters.permute(3, 2, 0, 1), stride=stride)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def conv2d_transpose(x, filters, stride=2, padding='SAME', bias=None):
    """tf.nn.conv2d_transpose + bias — tflib/ops/deconv2d.py:91-116 (NCHW in/out after the two transposes).

    x [B,Cin,H,W]; filters (k,k,Cout,Cin) as created at deconv2d.py:60-69.  TF defines the op as the input
    gradient of conv2d(out_shape -> x.shape) with the same filter, i.e. for SAME/stride s the full transposed
    correlation of size s*(H-1)+k cropped to s*H starting at pad_before of the *forward* conv.
    """
    k = filters.shape[0]
    B, Cin, H, W = x.shape
    if padding == 'SAME':
        Hout, Wout = stride * H, stride * W
        _, pt, _ = same_padding(Hout, k, stride)
        _, pl, _ = same_padding(Wout, k, stride)
    else:
        Hout, Wout = stride * (H - 1) + k, stride * (W - 1) + k
        pt = pl = 0
    # conv_transpose2d weight layout is (Cin, Cout, kH, kW); our filter is (kH,kW,Cout,Cin)
    full = F.conv_transpose2d(x, filters.permute(3, 2, 0, 1), stride=stride)
    y = full[:, :, pt:pt + Hout, pl:pl + Wout]
    if y.shape[2] < Hout or y.shape[3] < Wout:  # crop window reaches past the full output: zero pad
        y = F.pad(y, (0, Wout - y.shape[3], 0, Hout - y.shape[2]))
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def linear(x, W, b=None):
    """tf.matmul(inputs, weight) + bias_add — tflib/ops/linear.py:132-146."""
    y = x.reshape(-1, W.shape[0]) @ W
    if b is not None:
        y = y + b
    return y


def batchnorm(x, scale, offset, axes, eps=1e-5):
    """Batch-statistics batch norm — tflib/ops/batchnorm.py:29-30 (fused, axes [0,2,3]) and :77-84 (axes [0]).

    The scripts always pass is_training=None, so batch statistics are used at train and sample time.
    Biased variance; y = (x-mean) * rsqrt(var+eps) * scale + offset.
    """
    if list(axes) == [0, 2, 3]:
        mean = x.mean(dim=(0, 2, 3), keepdim=True)
        var = ((x - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
        return (x - mean) * torch.rsqrt(var + eps) * scale.view(1, -1, 1, 1) + offset.view(1, -1, 1, 1)
    if list(axes) == [0]:
        mean = x.mean(dim=0, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=0, keepdim=True)
        return (x - mean) * torch.rsqrt(var + eps) * scale.view(1, -1) + offset.view(1, -1)
    raise ValueError(axes)


def leaky_relu(x, alpha=0.2):
    """tf.maximum(alpha*x, x) — gmgan_inference_cifar10.py:122-123."""
    return torch.maximum(alpha * x, x)


def sigmoid_cross_entropy_with_logits(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1+exp(-|x|)) — gan_inference.py:85-101."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-torch.abs(logits)))


def bce_mean(logits, label):
    return sigmoid_cross_entropy_with_logits(logits, torch.full_like(logits, float(label))).mean()


def distance(x, y, d_type):
    """tflib/utils/distance.py:3-17."""
    x = x.reshape(-1, x.shape[-1])
    y = y.reshape(-1, y.shape[-1])
    if d_type == 'l1':
        return (x - y).abs().mean()
    if d_type == 'l2':
        return ((x - y) ** 2).mean()
    raise ValueError(d_type)


def sample_gumbel_from_uniform(U, eps=1e-20):
    """gmgan_inference_cifar10.py:117-120 with the uniform draw injected."""
    return -torch.log(-torch.log(U + eps) + eps)


def gradient_penalty(grad, weight=10.0):
    """slopes = sqrt(sum(g^2, axis=1)); weight * mean((slopes-1)^2) — gan_inference_svhn.py:353-354."""
    slopes = torch.sqrt((grad ** 2).sum(dim=1))
    return weight * ((slopes - 1.0) ** 2).mean()


# ----------------------------------------------------------------------------------------------
# local_ep / ali / wali_gp objectives                 [tflib/objs/gan_inference.py]
# ----------------------------------------------------------------------------------------------
def local_ep_costs(disc_fake_list, disc_real_list, s_f=None):
    """gan_inference.py:81-104: returns (gen_cost, disc_cost)."""
    gen, disc = 0.0, 0.0
    for df, dr in zip(disc_fake_list, disc_real_list):
        gen = gen + bce_mean(df, 1.0) + bce_mean(dr, 0.0)
        disc = disc + bce_mean(df, 0.0) + bce_mean(dr, 1.0)
    if s_f is not None:
        gen = gen + s_f
    n = len(disc_fake_list)
    return gen / n, disc / n


def ali_costs(disc_fake, disc_real, s_f=None):
    """gan_inference.py:47-66."""
    gen = bce_mean(disc_fake, 1.0) + bce_mean(disc_real, 0.0)
    disc = bce_mean(disc_fake, 0.0) + bce_mean(disc_real, 1.0)
    if s_f is not None:
        gen = gen + s_f
    return gen, disc


def local_epce_costs(disc_fake_list, disc_real_list, rec_penalty, s_f=None):
    """gan_inference.py:121-147: local_ep, then the reconstruction penalty is added AFTER the division by the list length."""
    gen, disc = local_ep_costs(disc_fake_list, disc_real_list, s_f)
    return gen + rec_penalty, disc


def alice_costs(disc_fake, disc_real, rec_penalty, s_f=None):
    """gan_inference.py:161-181."""
    gen, disc = ali_costs(disc_fake, disc_real, s_f)
    return gen + rec_penalty, disc


def vegan_costs(disc_fake, disc_real, rec_penalty, lamb, s_f=None):
    """gan_inference.py:194-212: the generator term sees only the fake logits; disc_cost is scaled by lamb/2."""
    gen = bce_mean(disc_fake, 1.0)
    if s_f is not None:
        gen = gen + s_f
    gen = gen * lamb + rec_penalty
    disc = (bce_mean(disc_fake, 0.0) + bce_mean(disc_real, 1.0)) * (lamb / 2)
    return gen, disc


def local_ep_dynamic_costs(disc_fake_zz, disc_real_zz, disc_fake_xz, disc_real_xz, rec_penalty=None):
    """gan_inference.py:246-293: the zz pairs are averaged over len+1, the xz pair is added un-normalised."""
    gen, disc = 0.0, 0.0
    for df, dr in zip(disc_fake_zz, disc_real_zz):
        gen = gen + bce_mean(df, 1.0) + bce_mean(dr, 0.0)
        disc = disc + bce_mean(df, 0.0) + bce_mean(dr, 1.0)
    if len(disc_fake_zz) > 0:
        gen = gen / (len(disc_fake_zz) + 1)
        disc = disc / (len(disc_fake_zz) + 1)
    gen = gen + bce_mean(disc_fake_xz, 1.0) + bce_mean(disc_real_xz, 0.0)
    disc = disc + bce_mean(disc_fake_xz, 0.0) + bce_mean(disc_real_xz, 1.0)
    if rec_penalty is not None:
        gen = gen + rec_penalty
    return gen, disc


def wali_gp_costs(disc_fake, disc_real, gp):
    """gan_inference.py:28-32."""
    gen = -disc_fake.mean() + disc_real.mean()
    disc = disc_fake.mean() - disc_real.mean() + gp
    return gen, disc


def weighted_local_epce_costs(disc_fake_list, disc_real_list, ratio_list, rec_penalty=None):
    """gan_inference.py:307-345 (no division by the list length)."""
    gen, disc = 0.0, 0.0
    for df, dr, r in zip(disc_fake_list, disc_real_list, ratio_list):
        gen = gen + float(r) * bce_mean(df, 1.0) + float(r) * bce_mean(dr, 0.0)
        disc = disc + float(r) * bce_mean(df, 0.0) + float(r) * bce_mean(dr, 1.0)
    if rec_penalty is not None:
        gen = gen + rec_penalty
    return gen, disc


# ----------------------------------------------------------------------------------------------
# tf.train.AdamOptimizer (ApplyAdam kernel)           [constructed at gan_inference.py:108-117]
# ----------------------------------------------------------------------------------------------
class TFAdam:
    """TensorFlow's Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
    p -= lr_t * m / (sqrt(v) + eps)  — epsilon is OUTSIDE the bias correction ("epsilon hat")."""

    def __init__(self, params, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def step(self, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                if g is None:
                    continue
                m.add_((g - m) * (1.0 - self.b1))
                v.add_((g * g - v) * (1.0 - self.b2))
                p.sub_(lr_t * m / (v.sqrt() + self.eps))


# ----------------------------------------------------------------------------------------------
# weight initialisers (numpy RandomState draws, float64 -> float32)  [conv2d.py:55-83, deconv2d.py:43-69, linear.py:39-106]
# ----------------------------------------------------------------------------------------------
def _uniform(rs, stdev, size):
    return rs.uniform(low=-stdev * np.sqrt(3), high=stdev * np.sqrt(3), size=size).astype('float32')


def init_conv2d(rs, input_dim, output_dim, k, stride=1, he_init=True):
    fan_in = input_dim * k ** 2
    fan_out = output_dim * k ** 2 / (stride ** 2)
    stdev = np.sqrt(4. / (fan_in + fan_out)) if he_init else np.sqrt(2. / (fan_in + fan_out))
    return _uniform(rs, stdev, (k, k, input_dim, output_dim))


def init_deconv2d(rs, input_dim, output_dim, k, stride=2, he_init=True):
    fan_in = input_dim * k ** 2 / (stride ** 2)
    fan_out = output_dim * k ** 2
    stdev = np.sqrt(4. / (fan_in + fan_out)) if he_init else np.sqrt(2. / (fan_in + fan_out))
    return _uniform(rs, stdev, (k, k, output_dim, input_dim))


def init_linear(rs, input_dim, output_dim, initialization=None):
    if initialization == 'lecun':
        return _uniform(rs, np.sqrt(1. / input_dim), (input_dim, output_dim))
    if initialization == 'glorot' or initialization is None:
        return _uniform(rs, np.sqrt(2. / (input_dim + output_dim)), (input_dim, output_dim))
    if initialization == 'he':
        return _uniform(rs, np.sqrt(2. / input_dim), (input_dim, output_dim))
    if initialization == 'glorot_he':
        return _uniform(rs, np.sqrt(4. / (input_dim + output_dim)), (input_dim, output_dim))
    raise ValueError(initialization)
