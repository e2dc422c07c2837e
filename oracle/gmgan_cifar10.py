"""CPU oracle of one GMGAN-CIFAR10 LOCAL_EP training step — restates gmgan_inference_cifar10.py (models :150-303,
graph :341-397, loop :480-494) functionally over oracle/tf_ops.py.  TEST INFRASTRUCTURE ONLY (see tf_ops.py header;
parity unpinned: the reference has no golden vectors and TensorFlow cannot run here).

Everything random is INJECTED (weights by name, p_z noise, prior component indices, Gumbel uniforms), because TF's
Philox streams are not reproducible without TF.  Gradients come from torch autograd over the
restated forward; the optimiser is the TF-form Adam of tf_ops.TFAdam.
"""
import numpy as np
import torch

from . import tf_ops as O

DIM = 64
DIM_LATENT = 128
N_COMS = 30
TEMP = 0.1


class GMGANCifar10(object):
    def __init__(self, params, dtype=torch.float32, dim=DIM, n_coms=N_COMS, threads=None, mode='local_ep', bn=None, lamb=1.):
        """params: {name: ndarray} with the tflib names ('Generator.2.Filters', 'Discriminator.zx1.W', ...).
        mode: 'local_ep' (north star) | 'local_epce' | 'ali' | 'alice' | 'vegan'  (:355-410); batch norm is on except in
        'vegan' (:59-63 of the script: BN_FLAG False, DIM_LATENT 8)."""
        if threads:
            torch.set_num_threads(threads)
        self.mode, self.lamb = mode, lamb
        self.bn = (mode != 'vegan') if bn is None else bn
        self.dtype = dtype
        self.dim = dim
        self.n_coms = n_coms
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(True) for k, v in params.items()
                  if 'moving_' not in k}
        self.gen_names = sorted(k for k in self.p if 'Generator' in k or 'Extractor' in k)
        self.disc_names = sorted(k for k in self.p if 'Discriminator' in k)
        self.gen_opt = O.TFAdam([self.p[k] for k in self.gen_names], lr=2e-4, beta1=0.5, beta2=0.999)
        self.disc_opt = O.TFAdam([self.p[k] for k in self.disc_names], lr=2e-4, beta1=0.5, beta2=0.999)

    # ---- networks -----------------------------------------------------------------------------
    def generator(self, noise):                                                     # :175-195
        p, D = self.p, self.dim
        bn = lambda t, name, axes: O.batchnorm(t, p[name + '.scale'], p[name + '.offset'], axes) if self.bn else t
        out = O.linear(noise, p['Generator.Input.W'], p['Generator.Input.b'])
        out = bn(out, 'Generator.BN1', [0])
        out = torch.relu(out).reshape(-1, 4 * D, 4, 4)
        out = O.conv2d_transpose(out, p['Generator.2.Filters'], 2, 'SAME', p['Generator.2.Biases'])
        out = torch.relu(bn(out, 'Generator.BN2', [0, 2, 3]))
        out = O.conv2d_transpose(out, p['Generator.3.Filters'], 2, 'SAME', p['Generator.3.Biases'])
        out = torch.relu(bn(out, 'Generator.BN3', [0, 2, 3]))
        out = O.conv2d_transpose(out, p['Generator.5.Filters'], 2, 'SAME', p['Generator.5.Biases'])
        return torch.tanh(out).reshape(-1, 3072)

    def extractor(self, x):                                                         # :197-231
        p = self.p
        out = x.reshape(-1, 3, 32, 32)
        out = O.leaky_relu(O.conv2d(out, p['Extractor.1.Filters'], 2, 'SAME', p['Extractor.1.Biases']))
        bn = lambda t, name, axes: O.batchnorm(t, p[name + '.scale'], p[name + '.offset'], axes) if self.bn else t
        out = O.conv2d(out, p['Extractor.2.Filters'], 2, 'SAME', p['Extractor.2.Biases'])
        out = O.leaky_relu(bn(out, 'Extractor.BN2', [0, 2, 3]))
        out = O.conv2d(out, p['Extractor.3.Filters'], 2, 'SAME', p['Extractor.3.Biases'])
        out = O.leaky_relu(bn(out, 'Extractor.BN3', [0, 2, 3]))
        out = out.reshape(-1, 4 * 4 * 4 * self.dim)
        return O.linear(out, p['Extractor.Output.W'], p['Extractor.Output.b'])

    def hyper_generator(self, k_onehot, noise):                                     # :150-153
        return k_onehot @ self.p['Generator.Hyper.Mu'] + noise

    def hyper_extractor(self, z, U):                                                # :156-173 (MODE_K == 'CONCRETE')
        mu = self.p['Generator.Hyper.Mu']
        log_pi = torch.log(torch.full((self.n_coms,), 1.0 / self.n_coms, dtype=torch.float32)).to(self.dtype)
        logits = -.5 * ((z[:, None, :] - mu[None, :, :]) ** 2).sum(-1) + log_pi[None, :]
        k = torch.softmax((logits + O.sample_gumbel_from_uniform(U)) / TEMP, dim=-1)
        return logits, k

    def hyper_discriminator(self, z, k):                                            # :262-278 (dropout = identity)
        p = self.p
        out = torch.cat([z, k], 1)
        out = O.leaky_relu(O.linear(out, p['Discriminator.HyperInput.W'], p['Discriminator.HyperInput.b']))
        out = O.leaky_relu(O.linear(out, p['Discriminator.Hyper2.W'], p['Discriminator.Hyper2.b']))
        out = O.leaky_relu(O.linear(out, p['Discriminator.Hyper3.W'], p['Discriminator.Hyper3.b']))
        return O.linear(out, p['Discriminator.HyperOutput.W'], p['Discriminator.HyperOutput.b']).reshape(-1)

    def discriminator(self, x, z):                                                  # :280-303
        p = self.p
        out = x.reshape(-1, 3, 32, 32)
        out = O.leaky_relu(O.conv2d(out, p['Discriminator.1.Filters'], 2, 'SAME', p['Discriminator.1.Biases']))
        out = O.leaky_relu(O.conv2d(out, p['Discriminator.2.Filters'], 2, 'SAME', p['Discriminator.2.Biases']))
        out = O.leaky_relu(O.conv2d(out, p['Discriminator.3.Filters'], 2, 'SAME', p['Discriminator.3.Biases']))
        out = out.reshape(-1, 4 * 4 * 4 * self.dim)
        zo = O.leaky_relu(O.linear(z, p['Discriminator.z1.W'], p['Discriminator.z1.b']))
        out = torch.cat([out, zo], 1)
        out = O.leaky_relu(O.linear(out, p['Discriminator.zx1.W'], p['Discriminator.zx1.b']))
        return O.linear(out, p['Discriminator.Output.W'], p['Discriminator.Output.b']).reshape(-1)

    def discriminator_xzk(self, x, z, k):                                           # :309-336 (ali / alice: one joint critic)
        p = self.p
        out = x.reshape(-1, 3, 32, 32)
        for i in (1, 2, 3):
            out = O.leaky_relu(O.conv2d(out, p['Discriminator.x%d.Filters' % i], 2, 'SAME', p['Discriminator.x%d.Biases' % i]))
        out = out.reshape(-1, 4 * 4 * 4 * self.dim)
        zk = O.leaky_relu(O.linear(torch.cat([z, k], 1), p['Discriminator.zk1.W'], p['Discriminator.zk1.b']))
        out = O.leaky_relu(O.linear(torch.cat([out, zk], 1), p['Discriminator.zkx1.W'], p['Discriminator.zkx1.b']))
        return O.linear(out, p['Discriminator.Output.W'], p['Discriminator.Output.b']).reshape(-1)

    # ---- graph (:341-397) -----------------------------------------------------------------------
    def costs(self, real_x_int, hyper_p_z, k_idx, U):
        t = lambda a: torch.as_tensor(np.asarray(a)).to(self.dtype)
        real_x = 2 * ((t(real_x_int) / 255.) - .5)
        q_z = self.extractor(real_x)
        _, q_k = self.hyper_extractor(q_z, t(U))
        k1h = torch.nn.functional.one_hot(torch.as_tensor(np.asarray(k_idx)).long(), self.n_coms).to(self.dtype)
        p_z = self.hyper_generator(k1h, t(hyper_p_z))
        fake_x = self.generator(p_z)
        rec = lambda: 1. * O.distance(real_x, self.generator(q_z), 'l2')           # DISTANCE_X = 'l2' (:52-55)
        if self.mode in ('local_ep', 'local_epce'):
            disc_fake = [self.hyper_discriminator(p_z, k1h), self.discriminator(fake_x, p_z)]
            disc_real = [self.hyper_discriminator(q_z, q_k), self.discriminator(real_x, q_z)]
            if self.mode == 'local_ep':
                gen_cost, disc_cost = O.local_ep_costs(disc_fake, disc_real)
            else:
                gen_cost, disc_cost = O.local_epce_costs(disc_fake, disc_real, rec())
        elif self.mode == 'vegan':                                                 # :355-359, critic on (z, k) only
            disc_fake, disc_real = self.hyper_discriminator(p_z, k1h), self.hyper_discriminator(q_z, q_k)
            gen_cost, disc_cost = O.vegan_costs(disc_fake, disc_real, rec(), self.lamb)
        else:                                                                      # ali / alice :371-376
            disc_real, disc_fake = self.discriminator_xzk(real_x, q_z, q_k), self.discriminator_xzk(fake_x, p_z, k1h)
            if self.mode == 'ali':
                gen_cost, disc_cost = O.ali_costs(disc_fake, disc_real)
            else:
                gen_cost, disc_cost = O.alice_costs(disc_fake, disc_real, rec())
        return gen_cost, disc_cost, dict(q_z=q_z, p_z=p_z, fake_x=fake_x, q_k=q_k, disc_fake=disc_fake, disc_real=disc_real)

    def disc_step(self, real_x_int, hyper_p_z, k_idx, U, apply=True):
        """session.run([disc_cost, disc_train_op])  (:489-494)"""
        _, disc_cost, _ = self.costs(real_x_int, hyper_p_z, k_idx, U)
        ps = [self.p[k] for k in self.disc_names]
        grads = torch.autograd.grad(disc_cost, ps, allow_unused=True)
        if apply:
            self.disc_opt.step(grads)
        return float(disc_cost.detach()), dict(zip(self.disc_names, grads))

    def gen_step(self, real_x_int, hyper_p_z, k_idx, U, apply=True):
        """session.run([gen_cost, gen_train_op])  (:483-487)"""
        gen_cost, _, _ = self.costs(real_x_int, hyper_p_z, k_idx, U)
        ps = [self.p[k] for k in self.gen_names]
        grads = torch.autograd.grad(gen_cost, ps, allow_unused=True)
        if apply:
            self.gen_opt.step(grads)
        return float(gen_cost.detach()), dict(zip(self.gen_names, grads))

    def sample(self, k_onehot, noise):
        with torch.no_grad():
            t = lambda a: torch.as_tensor(np.asarray(a)).to(self.dtype)
            return self.generator(self.hyper_generator(t(k_onehot), t(noise)))


def synthetic_inputs(batch_size, step, n_coms=N_COMS, dim_latent=DIM_LATENT):
    """seeded per-step inputs shared by the oracle and the CUDA path"""
    rs = np.random.RandomState(1000 + step)
    return dict(real_x_int=rs.randint(0, 256, size=(batch_size, 3072)).astype(np.int32),
                hyper_p_z=rs.randn(batch_size, dim_latent).astype(np.float32),
                k_idx=rs.randint(0, n_coms, size=(batch_size,)).astype(np.int32),
                U=rs.uniform(0, 1, size=(batch_size, n_coms)).astype(np.float32))
