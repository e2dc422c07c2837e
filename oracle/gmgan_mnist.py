"""CPU oracle of one GMGAN-MNIST LOCAL_EP training step (BASELINE.json configs[0]) — restates gmgan_inference_mnist.py
(models :141-296, graph :335-390) over oracle/tf_ops.py by overriding the three image networks of the CIFAR-10 oracle:
1x28x28 images fed as float [0,1], a 8x8 -> 7x7 crop between the first two deconvolutions (:179), sigmoid output (:187),
28 -> 14 -> 7 -> 4 strided convolutions (the 7 -> 4 layer pads (2,2)).  TEST INFRASTRUCTURE ONLY; parity unpinned (see
tf_ops.py header).
"""
import numpy as np
import torch

from . import tf_ops as O
from .gmgan_cifar10 import GMGANCifar10


class GMGANMnist(GMGANCifar10):
    def generator(self, noise):                                                     # :167-189
        p, D = self.p, self.dim
        out = O.linear(noise, p['Generator.Input.W'], p['Generator.Input.b'])
        out = O.batchnorm(out, p['Generator.BN1.scale'], p['Generator.BN1.offset'], [0])
        out = torch.relu(out).reshape(-1, 4 * D, 4, 4)
        out = O.conv2d_transpose(out, p['Generator.2.Filters'], 2, 'SAME', p['Generator.2.Biases'])
        out = torch.relu(O.batchnorm(out, p['Generator.BN2.scale'], p['Generator.BN2.offset'], [0, 2, 3]))
        out = out[:, :, :7, :7]
        out = O.conv2d_transpose(out, p['Generator.3.Filters'], 2, 'SAME', p['Generator.3.Biases'])
        out = torch.relu(O.batchnorm(out, p['Generator.BN3.scale'], p['Generator.BN3.offset'], [0, 2, 3]))
        out = O.conv2d_transpose(out, p['Generator.5.Filters'], 2, 'SAME', p['Generator.5.Biases'])
        return torch.sigmoid(out).reshape(-1, 784)

    def extractor(self, x):                                                         # :191-228
        p = self.p
        out = x.reshape(-1, 1, 28, 28)
        out = O.leaky_relu(O.conv2d(out, p['Extractor.1.Filters'], 2, 'SAME', p['Extractor.1.Biases']))
        out = O.conv2d(out, p['Extractor.2.Filters'], 2, 'SAME', p['Extractor.2.Biases'])
        out = O.leaky_relu(O.batchnorm(out, p['Extractor.BN2.scale'], p['Extractor.BN2.offset'], [0, 2, 3]))
        out = O.conv2d(out, p['Extractor.3.Filters'], 2, 'SAME', p['Extractor.3.Biases'])
        out = O.leaky_relu(O.batchnorm(out, p['Extractor.BN3.scale'], p['Extractor.BN3.offset'], [0, 2, 3]))
        out = out.reshape(-1, 4 * 4 * 4 * self.dim)
        return O.linear(out, p['Extractor.Output.W'], p['Extractor.Output.b'])

    def discriminator(self, x, z):                                                  # :267-296
        p = self.p
        out = x.reshape(-1, 1, 28, 28)
        out = O.leaky_relu(O.conv2d(out, p['Discriminator.1.Filters'], 2, 'SAME', p['Discriminator.1.Biases']))
        out = O.leaky_relu(O.conv2d(out, p['Discriminator.2.Filters'], 2, 'SAME', p['Discriminator.2.Biases']))
        out = O.leaky_relu(O.conv2d(out, p['Discriminator.3.Filters'], 2, 'SAME', p['Discriminator.3.Biases']))
        out = out.reshape(-1, 4 * 4 * 4 * self.dim)
        zo = O.leaky_relu(O.linear(z, p['Discriminator.z1.W'], p['Discriminator.z1.b']))
        out = torch.cat([out, zo], 1)
        out = O.leaky_relu(O.linear(out, p['Discriminator.zx1.W'], p['Discriminator.zx1.b']))
        return O.linear(out, p['Discriminator.Output.W'], p['Discriminator.Output.b']).reshape(-1)

    def costs(self, real_x, hyper_p_z, k_idx, U):                                   # :335-390 (float images, no decode)
        t = lambda a: torch.as_tensor(np.asarray(a)).to(self.dtype)
        real_x = t(real_x)
        q_z = self.extractor(real_x)
        _, q_k = self.hyper_extractor(q_z, t(U))
        k1h = torch.nn.functional.one_hot(torch.as_tensor(np.asarray(k_idx)).long(), self.n_coms).to(self.dtype)
        p_z = self.hyper_generator(k1h, t(hyper_p_z))
        fake_x = self.generator(p_z)
        disc_fake = [self.hyper_discriminator(p_z, k1h), self.discriminator(fake_x, p_z)]
        disc_real = [self.hyper_discriminator(q_z, q_k), self.discriminator(real_x, q_z)]
        gen_cost, disc_cost = O.local_ep_costs(disc_fake, disc_real)
        return gen_cost, disc_cost, dict(q_z=q_z, p_z=p_z, fake_x=fake_x, q_k=q_k, disc_fake=disc_fake, disc_real=disc_real)

    def disc_step(self, real_x, hyper_p_z, k_idx, U, apply=True):
        return GMGANCifar10.disc_step(self, real_x, hyper_p_z, k_idx, U, apply)

    def gen_step(self, real_x, hyper_p_z, k_idx, U, apply=True):
        return GMGANCifar10.gen_step(self, real_x, hyper_p_z, k_idx, U, apply)


def synthetic_inputs(batch_size, step, n_coms=30, dim_latent=128):
    rs = np.random.RandomState(2000 + step)
    return dict(real_x=rs.uniform(0, 1, size=(batch_size, 784)).astype(np.float32),
                hyper_p_z=rs.randn(batch_size, dim_latent).astype(np.float32),
                k_idx=rs.randint(0, n_coms, size=(batch_size,)).astype(np.int32),
                U=rs.uniform(0, 1, size=(batch_size, n_coms)).astype(np.float32))
