"""`jax.random` stand-in.  Threefry cannot be reproduced without JAX, so a key is a path in the split tree and every draw
comes from an injectable provider and is appended to `draw_log`; the golden-vector generator hands those exact draws to
the oracle / CUDA path as explicit buffers."""
import hashlib as _hashlib

import numpy as _np
import torch as _t

from ._core import Array, asarray, to_torch_dtype

draw_log = []          # dicts: kind, path, shape, u (the raw unit-interval / standard-normal draw, float32 numpy)
_provider = None       # callable(kind, path, shape) -> numpy float32 array or None


def set_provider(fn):
    global _provider
    _provider = fn


def clear_log():
    del draw_log[:]


class PRNGKeyArray:
    def __init__(self, path):
        self.path = tuple(path)
        self.shape = (2,)

    def __repr__(self):
        return f"Key{self.path}"

    def __iter__(self):          # `a, b = random.split(k)` goes through split(); a bare key is not iterable
        raise TypeError("key is not iterable")


class _KeyBatch(list):
    pass


def PRNGKey(seed):
    return PRNGKeyArray((int(seed),))


def split(key, num=2):
    return _KeyBatch(PRNGKeyArray(key.path + (i,)) for i in range(num))


def fold_in(key, data):
    return PRNGKeyArray(key.path + ('f', int(data)))


def _raw(kind, key, shape):
    shape = tuple(int(s) for s in shape)
    out = _provider(kind, key.path, shape) if _provider is not None else None
    if out is None:
        seed = int.from_bytes(_hashlib.sha256(repr((kind, key.path)).encode()).digest()[:8], 'little')
        g = _np.random.Generator(_np.random.Philox(seed))
        if kind == 'normal':
            out = g.standard_normal(shape, dtype=_np.float32)
        else:
            out = (g.integers(0, 1 << 23, size=shape, dtype=_np.int64).astype(_np.float32) * _np.float32(2.0 ** -23))
    out = _np.asarray(out, dtype=_np.float32).reshape(shape)
    draw_log.append(dict(kind=kind, path=key.path, shape=shape, u=out.copy()))
    return asarray(out)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    """jax.random.uniform: u01 * (maxval - minval) + minval, clamped below by minval (jax/_src/random.py)."""
    u = _raw('uniform', key, shape)
    minval_a, maxval_a = asarray(minval, _t.float32), asarray(maxval, _t.float32)
    out = u * (maxval_a - minval_a) + minval_a
    return _t.maximum(minval_a.expand_as(out), out).as_subclass(Array)


def normal(key, shape=(), dtype=None):
    return _raw('normal', key, shape)


def randint(key, shape, minval, maxval, dtype=None):
    """Integer draws in [minval, maxval); float bounds are cast to the integer dtype first (mip.py:324 -> always 0)."""
    lo, hi = int(minval), int(maxval)
    shape = tuple(int(s) for s in shape)
    if hi - lo <= 1:
        return asarray(_np.full(shape, lo, dtype=_np.int32))
    u = _raw('uniform', key, shape)
    return (asarray(_np.floor(_np.asarray(u) * (hi - lo)).astype(_np.int32)) + lo)
