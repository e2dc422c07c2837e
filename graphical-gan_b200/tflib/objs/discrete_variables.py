"""tflib.objs.discrete_variables — drop-in for tflib/objs/discrete_variables.py:4-9: the REINFORCE (score-function) surrogate
for the discrete mixture assignment, used by the gmgan scripts when MODE_K = 'REINFORCE' (default is 'CONCRETE')."""
import tensorflow as tf


def score_function(f_k, p_k, c_v):
    """estimates grad E_{p(k|params)} f(k) as grad( stop_gradient(f(k) - c_v) * log p(k|params) )"""
    return tf.stop_gradient(f_k - c_v) * tf.log(p_k)
