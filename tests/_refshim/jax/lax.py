"""`jax.lax` surface used by the reference (test-only shim)."""
import torch as _t

from ._core import Array, asarray
from . import tree_util as _tu


class Precision:
    DEFAULT = 'default'
    HIGH = 'high'
    HIGHEST = 'highest'


def stop_gradient(x):
    return _tu.tree_map(lambda v: asarray(v).detach() if isinstance(v, _t.Tensor) else v, x)


def pmean(x, axis_name=None):
    """One device per process in this shim: the mean over the mapped axis is the identity."""
    return x


def psum(x, axis_name=None):
    return x


def all_gather(x, axis_name=None):
    return _tu.tree_map(lambda v: asarray(v)[None], x)


def square(x):
    x = asarray(x)
    return x * x


def rsqrt(x):
    return _t.rsqrt(asarray(x))
