"""GPU diagnostic: run the gmgan-CIFAR training plans launch by launch (eager, a synchronize after every launch) and name the
first launch that fails; then once through the captured graphs."""
import faulthandler
import os
import sys

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graphical-gan_b200"), os.path.join(ROOT, "graphical-gan_b200", "scripts")):
    sys.path.insert(0, p)
import numpy as np
import torch

from gg import cabi
from gg.executor import RT, Plan
import tensorflow as tf  # noqa
import gmgan_inference_cifar10 as S

_real = cabi.call
_last = [None]


def traced(name, *args):
    _last[0] = (name, args)
    return _real(name, *args)


def main():
    np.random.seed(1234)
    g = S.build_graph(BATCH_SIZE=64)
    x = np.random.randint(0, 256, size=(64, 3072)).astype(np.int32)
    for label, fetch in (("D", [g.disc_cost, g.disc_train_op]), ("G", [g.gen_cost, g.gen_train_op])):
        plan = Plan(RT, fetch, [g.real_x_int])
        d, _h = RT.feed_buffer(g.real_x_int)
        d.copy_(torch.from_numpy(x.reshape(-1)).cuda())
        torch.cuda.synchronize()
        cabi.call = traced
        st = cabi.stream_ptr()
        for si, f in enumerate(plan.steps):
            try:
                f(st)
                torch.cuda.synchronize()
            except Exception as e:  # noqa
                print("%s step %d FAILED after call %s%r: %r" % (label, si, _last[0][0], _last[0][1][:12], e), flush=True)
                return 1
        cabi.call = _real
        print("%s eager ok: %d steps" % (label, len(plan.steps)), flush=True)
        out = plan.run({g.real_x_int: x})
        torch.cuda.synchronize()
        print("%s graph ok: cost %r" % (label, out[0]), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
