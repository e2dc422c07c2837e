"""CPU oracle of the WGAN-GP rows of the hot path — restates gan_inference_svhn.py MODE='wali-gp' (:342-357, objective
tflib/objs/gan_inference.py:28-45) and MODE='vegan-wgan-gp' (:302-316, objective :225-244) over oracle/tf_ops.py.
TEST INFRASTRUCTURE ONLY; parity unpinned (see tf_ops.py).  All noise (p_z, alpha, the critic's Gaussian noise layers) is
injected.  The gradient penalty is differentiated through with torch autograd (create_graph=True) — the counterpart of
TensorFlow differentiating through tf.gradients."""
import numpy as np
import torch

from . import tf_ops as O


class GanSvhn(object):
    def __init__(self, params, mode, dtype=torch.float64, dim=64):
        self.mode, self.dtype, self.dim = mode, dtype, dim
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(True) for k, v in params.items()}
        self.disc_names = sorted(k for k in self.p if 'Discriminator' in k)
        self.gen_names = sorted(k for k in self.p if 'Generator' in k or 'Extractor' in k)

    def t(self, a):
        return torch.as_tensor(np.asarray(a)).to(self.dtype)

    def generator(self, z):                                                          # :129-149 (BN_FLAG False)
        p, D = self.p, self.dim
        out = torch.relu(O.linear(z, p['Generator.Input.W'], p['Generator.Input.b'])).reshape(-1, 4 * D, 4, 4)
        out = torch.relu(O.conv2d_transpose(out, p['Generator.2.Filters'], 2, 'SAME', p['Generator.2.Biases']))
        out = torch.relu(O.conv2d_transpose(out, p['Generator.3.Filters'], 2, 'SAME', p['Generator.3.Biases']))
        out = torch.tanh(O.conv2d_transpose(out, p['Generator.5.Filters'], 2, 'SAME', p['Generator.5.Biases']))
        return out.reshape(-1, 3072)

    def extractor(self, x):                                                          # :151-180
        p = self.p
        out = x.reshape(-1, 3, 32, 32)
        for i in (1, 2, 3):
            out = O.leaky_relu(O.conv2d(out, p['Extractor.%d.Filters' % i], 2, 'SAME', p['Extractor.%d.Biases' % i]))
        return O.linear(out.reshape(out.shape[0], -1), p['Extractor.Output.W'], p['Extractor.Output.b'])

    def critic_xz(self, x, z):                                                       # :216-240
        p = self.p
        out = x.reshape(-1, 3, 32, 32)
        for i in (1, 2, 3):
            out = O.leaky_relu(O.conv2d(out, p['Discriminator.%d.Filters' % i], 2, 'SAME', p['Discriminator.%d.Biases' % i]))
        zo = O.leaky_relu(O.linear(z, p['Discriminator.z1.W'], p['Discriminator.z1.b']))
        out = torch.cat([out.reshape(out.shape[0], -1), zo], 1)
        out = O.leaky_relu(O.linear(out, p['Discriminator.zx1.W'], p['Discriminator.zx1.b']))
        return O.linear(out, p['Discriminator.Output.W'], p['Discriminator.Output.b']).reshape(-1)

    def critic_z(self, z, noises):                                                   # :184-209, noise layers injected
        p = self.p
        out = z + noises[0]
        out = O.leaky_relu(O.linear(out, p['Discriminator.Input.W'], p['Discriminator.Input.b'])) + noises[1]
        out = O.leaky_relu(O.linear(out, p['Discriminator.2.W'], p['Discriminator.2.b'])) + noises[2]
        out = O.leaky_relu(O.linear(out, p['Discriminator.3.W'], p['Discriminator.3.b'])) + noises[3]
        out = O.leaky_relu(O.linear(out, p['Discriminator.4.W'], p['Discriminator.4.b']))
        return O.linear(out, p['Discriminator.Output.W'], p['Discriminator.Output.b']).reshape(-1)

    def wali_gp(self, real_x_int, p_z, alpha):
        real_x = 2 * ((self.t(real_x_int) / 255.) - .5)
        p_z, alpha = self.t(p_z), self.t(alpha)
        q_z = self.extractor(real_x)
        fake_x = self.generator(p_z)
        disc_real, disc_fake = self.critic_xz(real_x, q_z), self.critic_xz(fake_x, p_z)
        xi = real_x + alpha * (fake_x - real_x)
        zi = q_z + alpha * (p_z - q_z)
        gx = torch.autograd.grad(self.critic_xz(xi, zi).sum(), xi, create_graph=True)[0]      # [0]: the x part only
        gp = O.gradient_penalty(gx, 10.0)
        gen_cost, disc_cost = O.wali_gp_costs(disc_fake, disc_real, gp)
        return gen_cost, disc_cost, gp

    def vegan_wgan_gp(self, real_x_int, p_z, alpha, noises, lamb=1.0):
        """noises: three lists (real / fake / interpolate critic calls) of the four GaussianNoiseLayer draws"""
        real_x = 2 * ((self.t(real_x_int) / 255.) - .5)
        p_z, alpha = self.t(p_z), self.t(alpha)
        q_z = self.extractor(real_x)
        rec_x = self.generator(q_z)
        disc_real = self.critic_z(p_z, [self.t(n) for n in noises[0]])
        disc_fake = self.critic_z(q_z, [self.t(n) for n in noises[1]])
        zi = p_z + alpha * (q_z - p_z)
        g = torch.autograd.grad(self.critic_z(zi, [self.t(n) for n in noises[2]]).sum(), zi, create_graph=True)[0]
        gp = O.gradient_penalty(g, 10.0)
        rec = O.distance(real_x, rec_x, 'l2')
        gen_cost = (-disc_fake.mean() + disc_real.mean()) * lamb + rec
        disc_cost = (disc_fake.mean() - disc_real.mean()) * lamb + gp
        return gen_cost, disc_cost, gp

    def grads(self, cost, names):
        return dict(zip(names, torch.autograd.grad(cost, [self.p[k] for k in names], allow_unused=True, retain_graph=True)))
