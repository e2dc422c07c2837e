"""`import tensorflow as tf` shim: re-exports graphical-gan_b200/gg/tf_api.py (the ~45 tf.* symbols the reference's
scripts use) so py3 ports of the *_inference_*.py scripts run against the B200 kernels without touching their bodies.
This is NOT TensorFlow."""
import os as _os
import sys as _sys

_pkg = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _pkg not in _sys.path:
    _sys.path.insert(0, _pkg)

from gg.tf_api import *          # noqa: F401,F403
from gg.tf_api import nn, layers, distributions, train, float32, float64, int32, int64, abs, pow   # noqa: F401,A004
from gg import tf_api as _api

__version__ = "1.4.0-gg_b200-shim"
