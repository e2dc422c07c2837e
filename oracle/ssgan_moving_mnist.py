"""CPU oracle of one SSGAN training step on moving MNIST — restates /root/reference/ssgan_inference_moving_mnist.py
(ImplicitOperator :98-114, DynamicGenerator :134-141, DynamicExtractor naive_mean_field :143-168, Generator :170-205,
Extractor :207-236, G_Extractor :238-265, Discriminator :267-311, DynamicDiscrminator :313-331, ZGDiscrminator :333-349,
graph :508-549) and tflib/objs/gan_inference.py:307-358 (weighted_local_epce) functionally over oracle/tf_ops.py.
TEST INFRASTRUCTURE ONLY (see tf_ops.py header; parity unpinned: no golden vectors in the reference, TensorFlow absent).

Defaults of the script: MODE 'local_ep' (or 'local_epce-z'), POS_MODE 'naive_mean_field', OP_DYN_MODE 'res', BN off,
DIM 32, DIM_OP 256, z_g 128, z_l 8, 10 classes, frames 1x64x64.  BASELINE.json configs[4]: LEN 8, bs 32.
Everything random is INJECTED: weights by tflib name, p_z_l_0, the ONE epsilon shared by all LEN-1 unrolled transitions
(:137-139), p_z_g and the prior class indices.
"""
import numpy as np
import torch

from . import tf_ops as O

DIM_LATENT_G, DIM_LATENT_L, N_C = 128, 8, 10
OUTPUT_DIM = 64 * 64
LAMBDA = 0.1


class SSGANMovingMNIST(object):
    def __init__(self, params, batch_size, length, mode='local_ep', dtype=torch.float32, dim=32, lr=1e-4, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.B, self.LEN, self.mode, self.dtype, self.dim = batch_size, length, mode, dtype, dim
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(True) for k, v in params.items()}
        self.gen_names = sorted(k for k in self.p if 'Generator' in k or 'Extractor' in k)
        self.disc_names = sorted(k for k in self.p if 'Discriminator' in k)
        self.gen_opt = O.TFAdam([self.p[k] for k in self.gen_names], lr=lr, beta1=0.5, beta2=0.999)
        self.disc_opt = O.TFAdam([self.p[k] for k in self.disc_names], lr=lr, beta1=0.5, beta2=0.999)
        ratio = [1.0, ] * (length - 1) + [1, length]                               # :78-79
        self.ratio = np.asarray(ratio) * 1.0 / (len(ratio) + length - 1)

    def _lin(self, name, x):
        return O.linear(x, self.p[name + '.W'], self.p[name + '.b'])

    def expand_labels(self, y):                                                    # :88-90
        return y[:, None, :].expand(-1, self.LEN, -1).reshape(-1, N_C)

    def implicit_operator(self, z_l, epsilon, name='Generator.Dynamic'):           # :98-114, OP_DYN_MODE 'res'
        out = O.leaky_relu(self._lin(name + '.Input', torch.cat([z_l, epsilon], 1)))
        out = O.leaky_relu(self._lin(name + '.1', out))
        return self._lin(name + '.Output', out) + z_l

    def dynamic_generator(self, z_l_0, epsilon):                                   # :134-141: ONE epsilon for every transition
        zs = [z_l_0]
        for _ in range(self.LEN - 1):
            zs.append(self.implicit_operator(zs[-1], epsilon))
        return torch.cat(zs, 1).reshape(self.B, self.LEN, DIM_LATENT_L)

    def _z(self, z_g, z_l, labels):
        z_g = z_g.reshape(self.B, 1, DIM_LATENT_G).expand(-1, self.LEN, -1)
        z_l = z_l.reshape(self.B, self.LEN, DIM_LATENT_L)
        lab = self.expand_labels(labels).reshape(self.B, self.LEN, N_C)
        return torch.cat([z_g, z_l, lab], -1).reshape(self.B * self.LEN, DIM_LATENT_G + DIM_LATENT_L + N_C)

    def generator(self, z_g, z_l, labels):                                         # :170-205
        p, D = self.p, self.dim
        out = torch.relu(self._lin('Generator.Input', self._z(z_g, z_l, labels))).reshape(self.B * self.LEN, 8 * D, 4, 4)
        for i in (2, 3, 4):
            out = torch.relu(O.conv2d_transpose(out, p['Generator.%d.Filters' % i], 2, 'SAME', p['Generator.%d.Biases' % i]))
        out = torch.tanh(O.conv2d_transpose(out, p['Generator.5.Filters'], 2, 'SAME', p['Generator.5.Biases']))
        return out.reshape(self.B, self.LEN, OUTPUT_DIM)

    def _trunk(self, prefix, out):
        for i in (1, 2, 3, 4):
            out = O.leaky_relu(O.conv2d(out, self.p['%s%d.Filters' % (prefix, i)], 2, 'SAME', self.p['%s%d.Biases' % (prefix, i)]))
        return out.reshape(out.shape[0], -1)

    def extractor(self, x, labels):                                                # :207-236
        out = self._trunk('Extractor.', x.reshape(self.B * self.LEN, 1, 64, 64))
        out = torch.cat([out, self.expand_labels(labels)], 1)
        return self._lin('Extractor.Output', out).reshape(self.B, self.LEN, DIM_LATENT_L)

    def g_extractor(self, x, labels):                                              # :238-265: the LEN frames are the channels
        out = self._trunk('Extractor.G.', x.reshape(self.B, self.LEN, 64, 64))
        return self._lin('Extractor.G.Output', torch.cat([out, labels], 1)).reshape(self.B, DIM_LATENT_G)

    def discriminator(self, x, z_g, z_l, labels):                                  # :267-311 (dropout = identity)
        z = self._z(z_g, z_l, labels)
        out = self._trunk('Discriminator.', x.reshape(self.B * self.LEN, 1, 64, 64))
        zo = O.leaky_relu(self._lin('Discriminator.z1', z))
        lab = self.expand_labels(labels)
        out = O.leaky_relu(self._lin('Discriminator.zx1', torch.cat([out, zo, lab], 1)))
        return self._lin('Discriminator.Output', out).reshape(self.B * self.LEN)

    def _mlp(self, prefix, x):
        out = O.leaky_relu(self._lin(prefix + '.Input', x))
        out = O.leaky_relu(self._lin(prefix + '.2', out))
        out = O.leaky_relu(self._lin(prefix + '.3', out))
        return self._lin(prefix + '.Output', out).reshape(self.B)

    def dynamic_discriminator(self, z1, z2):                                       # :313-331
        return self._mlp('Discriminator.Dynamic', torch.cat([z1.reshape(self.B, -1), z2.reshape(self.B, -1)], 1))

    def zg_discriminator(self, z_g):                                               # :333-349
        return self._mlp('Discriminator.ZG', z_g.reshape(self.B, DIM_LATENT_G))

    def costs(self, real_x_unit, real_y, p_z_l_0, epsilon, p_z_g, p_y_idx):        # :508-549
        t = lambda a: torch.as_tensor(np.asarray(a)).to(self.dtype)
        real_x = 2 * (t(real_x_unit) - .5)
        real_y = t(real_y)
        q_z_l = self.extractor(real_x, real_y)                                     # naive_mean_field: q_z_l = q_z_l_pre
        q_z_g = self.g_extractor(real_x, real_y)
        p_z_l = self.dynamic_generator(t(p_z_l_0), t(epsilon))
        p_z_g = t(p_z_g)
        p_y = torch.nn.functional.one_hot(torch.as_tensor(np.asarray(p_y_idx)).long(), N_C).to(self.dtype)
        fake_x = self.generator(p_z_g, p_z_l, p_y)
        disc_fake, disc_real = [], []
        for i in range(self.LEN - 1):
            disc_fake.append(self.dynamic_discriminator(p_z_l[:, i, :], p_z_l[:, i + 1, :]))
            disc_real.append(self.dynamic_discriminator(q_z_l[:, i, :], q_z_l[:, i + 1, :]))
        disc_fake.append(self.zg_discriminator(p_z_g))
        disc_real.append(self.zg_discriminator(q_z_g))
        disc_fake.append(self.discriminator(fake_x, p_z_g, p_z_l, p_y))
        disc_real.append(self.discriminator(real_x, q_z_g, q_z_l, real_y))
        rec_penalty = None
        if self.mode == 'local_epce-z':
            rec_x = self.generator(q_z_g, q_z_l, real_y)
            rec_penalty = LAMBDA * O.distance(real_x, rec_x, 'l2')
        gen_cost, disc_cost = O.weighted_local_epce_costs(disc_fake, disc_real, self.ratio, rec_penalty)
        return gen_cost, disc_cost, dict(p_z_l=p_z_l, q_z_l=q_z_l, q_z_g=q_z_g, fake_x=fake_x)

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


def synthetic_inputs(batch_size, length, step):
    rs = np.random.RandomState(5000 + step)
    return dict(real_x_unit=rs.uniform(0, 1, size=(batch_size, length, OUTPUT_DIM)).astype(np.float32),
                real_y=np.eye(N_C, dtype=np.float32)[rs.randint(0, N_C, size=batch_size)],
                p_z_l_0=rs.randn(batch_size, DIM_LATENT_L).astype(np.float32),
                epsilon=rs.randn(batch_size, DIM_LATENT_L).astype(np.float32),
                p_z_g=rs.randn(batch_size, DIM_LATENT_G).astype(np.float32),
                p_y_idx=rs.randint(0, N_C, size=batch_size).astype(np.int32))
