"""CPU oracle of one ALI training step on 64x64x3 faces — restates /root/reference/gan_inference_face.py (models :75-148,
graph :150-176) functionally over oracle/tf_ops.py.  TEST INFRASTRUCTURE ONLY (see tf_ops.py header; parity unpinned: the
reference has no golden vectors and TensorFlow cannot run here).

BASELINE.json configs[3]: bs=128, DIM_G=DIM_D=32, DIM_LATENT=128, MODE='ali' (:33-42), no batch norm anywhere.  Input decode
(:151-152): 2*((int/256)-.5) + U[0,1/128) dequantisation noise.  Everything random is INJECTED (weights by tflib name,
p_z, dequantisation noise); gradients come from torch autograd; the optimiser is tf_ops.TFAdam (lr 2e-4, beta1 .5).
"""
import numpy as np
import torch

from . import tf_ops as O

DIM_LATENT = 128
OUTPUT_DIM = 64 * 64 * 3


class GANFace(object):
    def __init__(self, params, dtype=torch.float32, dim_g=32, dim_d=32, lr=2e-4, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.dtype, self.dim_g, self.dim_d = dtype, dim_g, dim_d
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(True) for k, v in params.items()}
        self.gen_names = sorted(k for k in self.p if 'Generator' in k or 'Extractor' in k)
        self.disc_names = sorted(k for k in self.p if 'Discriminator' in k)
        self.gen_opt = O.TFAdam([self.p[k] for k in self.gen_names], lr=lr, beta1=0.5, beta2=0.999)
        self.disc_opt = O.TFAdam([self.p[k] for k in self.disc_names], lr=lr, beta1=0.5, beta2=0.999)

    def generator(self, noise):                                                     # :75-94
        p, D = self.p, self.dim_g
        out = torch.relu(O.linear(noise, p['Generator.Input.W'], p['Generator.Input.b'])).reshape(-1, 8 * D, 4, 4)
        for i in (2, 3, 4):
            out = torch.relu(O.conv2d_transpose(out, p['Generator.%d.Filters' % i], 2, 'SAME', p['Generator.%d.Biases' % i]))
        out = torch.tanh(O.conv2d_transpose(out, p['Generator.5.Filters'], 2, 'SAME', p['Generator.5.Biases']))
        return out.reshape(-1, OUTPUT_DIM)

    def _trunk(self, prefix, x):
        out = x.reshape(-1, 3, 64, 64)
        for i in (1, 2, 3, 4):
            out = O.leaky_relu(O.conv2d(out, self.p['%s.%d.Filters' % (prefix, i)], 2, 'SAME', self.p['%s.%d.Biases' % (prefix, i)]))
        return out.reshape(out.shape[0], -1)

    def extractor(self, x):                                                         # :96-114
        out = self._trunk('Extractor', x)
        return O.linear(out, self.p['Extractor.Output.W'], self.p['Extractor.Output.b'])

    def discriminator(self, x, z):                                                  # :116-148 (dropout = identity)
        p = self.p
        out = self._trunk('Discriminator', x)
        zo = O.leaky_relu(O.linear(z, p['Discriminator.z1.W'], p['Discriminator.z1.b']))
        out = O.leaky_relu(O.linear(torch.cat([out, zo], 1), p['Discriminator.zx1.W'], p['Discriminator.zx1.b']))
        return O.linear(out, p['Discriminator.Output.W'], p['Discriminator.Output.b']).reshape(-1)

    def costs(self, real_x_int, dequant, p_z):                                      # :150-170
        t = lambda a: torch.as_tensor(np.asarray(a)).to(self.dtype)
        real_x = 2 * ((t(real_x_int) / 256.) - .5) + t(dequant)
        q_z = self.extractor(real_x)
        p_z = t(p_z)
        fake_x = self.generator(p_z)
        disc_real = self.discriminator(real_x, q_z)
        disc_fake = self.discriminator(fake_x, p_z)
        gen_cost, disc_cost = O.ali_costs(disc_fake, disc_real)
        return gen_cost, disc_cost, dict(q_z=q_z, fake_x=fake_x, disc_fake=disc_fake, disc_real=disc_real)

    def _step(self, which, apply, **inp):
        gen_cost, disc_cost, _ = self.costs(**inp)
        cost, names, opt = (gen_cost, self.gen_names, self.gen_opt) if which == "gen" else (disc_cost, self.disc_names, self.disc_opt)
        ps = [self.p[k] for k in names]
        grads = torch.autograd.grad(cost, ps, allow_unused=True)
        if apply:
            opt.step(grads)
        return float(cost.detach()), dict(zip(names, grads))

    def gen_step(self, apply=True, **inp):
        return self._step("gen", apply, **inp)

    def disc_step(self, apply=True, **inp):
        return self._step("disc", apply, **inp)


def synthetic_inputs(batch_size, step):
    rs = np.random.RandomState(3000 + step)
    return dict(real_x_int=rs.randint(0, 256, size=(batch_size, OUTPUT_DIM)).astype(np.int32),
                dequant=rs.uniform(0, 1. / 128, size=(batch_size, OUTPUT_DIM)).astype(np.float32),
                p_z=rs.randn(batch_size, DIM_LATENT).astype(np.float32))
