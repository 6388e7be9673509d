"""`jax.nn` activations and initializers (test-only shim)."""
import math as _math

import torch as _t

from .._core import Array, asarray
from .. import random as _random
from . import initializers  # noqa: F401


def relu(x):
    x = asarray(x)
    return _t.maximum(x, _t.zeros_like(x)).as_subclass(Array)       # jnp.maximum(x, 0)


def sigmoid(x):
    return _t.sigmoid(asarray(x)).as_subclass(Array)                 # expit


def softplus(x):
    x = asarray(x)
    return _t.logaddexp(x, _t.zeros_like(x)).as_subclass(Array)     # jnp.logaddexp(x, 0)
