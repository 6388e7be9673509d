import math as _math

from .. import random as _random


def glorot_uniform():
    """variance_scaling(1.0, 'fan_avg', 'uniform'): U(-a, a), a = sqrt(6 / (fan_in + fan_out))."""
    def init(key, shape, dtype=None):
        fan_in, fan_out = shape[-2], shape[-1]
        a = _math.sqrt(3.0 * 1.0 / ((fan_in + fan_out) / 2.0))
        return _random.uniform(key, shape, minval=-a, maxval=a)
    return init


def zeros(key, shape, dtype=None):
    from .. import numpy as jnp
    return jnp.zeros(shape)
