"""tflib.plot — text-only equivalent of tflib/plot.py:12-41 (matplotlib is not a dependency of the hot path):
plot(name, value) records a scalar for the current iteration, flush() prints the mean since the last flush and
appends the same line to logfile.txt, tick() advances the iteration counter."""
import collections

import numpy as np

_since_beginning = collections.defaultdict(lambda: {})
_since_last_flush = collections.defaultdict(lambda: {})
_iter = [0]


def tick():
    _iter[0] += 1


def plot(name, value):
    _since_last_flush[name][_iter[0]] = value


def flush(outf=None, logfile=None):
    prints = []
    for name, vals in _since_last_flush.items():
        prints.append("{}\t{:.3f}".format(name, np.mean(list(vals.values()))))
        _since_beginning[name].update(vals)
    line = "iter {}\t{}".format(_iter[0], "\t".join(prints))
    print(line)
    if logfile is not None:
        with open(logfile, 'a') as f:
            f.write(line + '\n')
    _since_last_flush.clear()
