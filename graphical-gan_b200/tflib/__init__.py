"""tflib — drop-in for the reference's helper package (tflib/__init__.py:7-121), over the B200 graph runtime.

Same module paths, names and signatures as /root/reference/tflib; the parameter registry keeps the reference's
semantics: `param(name, value)` creates the variable once and returns the SAME object on every later call with
that name (weight sharing between repeated Generator(...)/Discriminator(...) calls), and `params_with_name(s)`
selects by substring — which is how the scripts build the G/E and D optimiser var-lists
(gmgan_inference_cifar10.py:381-383).
"""
import numpy as np
import tensorflow as tf

_params = {}
_param_aliases = {}


def param(name, *args, **kwargs):
    """tflib/__init__.py:9-33."""
    if name not in _params:
        kwargs['name'] = name
        param = tf.Variable(*args, **kwargs)
        param.param = True
        _params[name] = param
    result = _params[name]
    while result in _param_aliases:
        result = _param_aliases[result]
    return result


def params_with_name(name):
    """tflib/__init__.py:35-36."""
    return [p for n, p in _params.items() if name in n]


def delete_all_params():
    _params.clear()


def alias_params(replace_dict):
    for old, new in replace_dict.items():
        _param_aliases[old] = new


def delete_param_aliases():
    _param_aliases.clear()


def _settings(locals_):
    all_vars = [(k, v) for (k, v) in locals_.items()
                if (k.isupper() and k != 'T' and k != 'SETTINGS' and k != 'ALL_SETTINGS')]
    return sorted(all_vars, key=lambda x: x[0])


def print_model_settings(locals_):
    """tflib/__init__.py:100-105."""
    print("Uppercase local vars:")
    for var_name, var_value in _settings(locals_):
        print("\t{}: {}".format(var_name, var_value))


def print_model_settings_to_file(locals_, logfile):
    """tflib/__init__.py:107-114."""
    print("Uppercase local vars:")
    for var_name, var_value in _settings(locals_):
        print("\t{}: {}".format(var_name, var_value))
        with open(logfile, 'a') as f:
            f.write("\t{}: {}".format(var_name, var_value))


def print_model_settings_dict(settings):
    """tflib/__init__.py:116-121."""
    print("Settings dict:")
    for var_name, var_value in sorted(settings.items(), key=lambda x: x[0]):
        print("\t{}: {}".format(var_name, var_value))
