"""`jax.numpy` surface used by the reference, backed by torch-CPU float32.  Test-only; see ../README.md."""
import builtins as _b
import math as _math

import numpy as _np
import torch as _t

from ._core import Array, asarray, is_scalar, lift, to_torch_dtype

ndarray = Array
pi = _math.pi
inf = float('inf')
nan = float('nan')
newaxis = None
# numpy scalar types: usable as dtype arguments of numpy AND of this module, and callable (`jnp.float32(x)`)
float32 = _np.float32
float64 = _np.float32
float16 = _np.float16
int32 = _np.int32
int64 = _np.int32
uint8 = _np.uint8
bool_ = _np.bool_
float_ = _np.float32


def finfo(dt):
    if isinstance(dt, _t.dtype):
        dt = {_t.float32: 'float32', _t.float16: 'float16'}[dt]
    return _np.finfo(dt)


def iinfo(dt):
    return _np.iinfo(dt)


def _w(x):
    return x if isinstance(x, Array) else x.as_subclass(Array)


def _pair(a, b):
    """Binary-op operands: arrays converted, python scalars left weak (at least one side becomes an array)."""
    a, b = lift(a), lift(b)
    if is_scalar(a) and is_scalar(b):
        a = asarray(a)
    return a, b


def _axis(axis):
    return tuple(axis) if isinstance(axis, list) else axis


# ---------------------------------------------------------------- creation
def array(x, dtype=None, copy=True):
    out = asarray(x, dtype)
    return out


def asarray_(x, dtype=None):
    return asarray(x, dtype)


def zeros(shape, dtype=None):
    shape = (shape,) if isinstance(shape, int) else tuple(int(s) for s in shape)
    return _w(_t.zeros(shape, dtype=to_torch_dtype(dtype) or _t.float32))


def ones(shape, dtype=None):
    shape = (shape,) if isinstance(shape, int) else tuple(int(s) for s in shape)
    return _w(_t.ones(shape, dtype=to_torch_dtype(dtype) or _t.float32))


def zeros_like(x, dtype=None):
    return _w(_t.zeros_like(asarray(x), dtype=to_torch_dtype(dtype)))


def ones_like(x, dtype=None):
    return _w(_t.ones_like(asarray(x), dtype=to_torch_dtype(dtype)))


def full(shape, v, dtype=None):
    return _w(_t.full(tuple(shape), v, dtype=to_torch_dtype(dtype) or _t.float32))


def eye(n, dtype=None):
    return _w(_t.eye(int(n), dtype=to_torch_dtype(dtype) or _t.float32))


def arange(*a, dtype=None):
    floaty = _b.any(isinstance(v, float) for v in a)
    return _w(_t.arange(*a, dtype=to_torch_dtype(dtype) or (_t.float32 if floaty else _t.int32)))


def linspace(start, stop, num=50, endpoint=True, dtype=None):
    """jnp.linspace (jax 0.2.x): start + iota(div) * ((stop - start) / div), the endpoint appended verbatim."""
    num = int(num)
    start_a, stop_a = asarray(start, _t.float32), asarray(stop, _t.float32)
    if num == 1:
        return _w(start_a.reshape(1))
    div = num - 1 if endpoint else num
    delta = (stop_a - start_a) / div
    body = start_a + _t.arange(div, dtype=_t.float32) * delta
    if endpoint:
        return _w(_t.cat([body, stop_a.reshape(1)]))
    return _w(body)


# ---------------------------------------------------------------- elementwise
def _unary(fn):
    def f(x):
        return _w(fn(asarray(x)))
    return f


def _float_unary(fn):
    def f(x):
        x = asarray(x)
        if not x.dtype.is_floating_point:
            x = x.to(_t.float32)
        return _w(fn(x))
    return f


sin = _float_unary(_t.sin)
cos = _float_unary(_t.cos)
exp = _float_unary(_t.exp)
log = _float_unary(_t.log)
sqrt = _float_unary(_t.sqrt)
tanh = _float_unary(_t.tanh)
abs = _unary(_t.abs)
absolute = abs
sign = _unary(_t.sign)
isnan = _unary(_t.isnan)
isinf = _unary(_t.isinf)
isfinite = _unary(_t.isfinite)
floor = _unary(_t.floor)
round = _unary(_t.round)
arccos = _float_unary(_t.arccos)
arcsin = _float_unary(_t.arcsin)
log2 = _float_unary(_t.log2)
log10 = _float_unary(_t.log10)


def mod(a, b):
    a, b = _pair(a, b)
    return _w(_t.remainder(a, b))


def histogram(x, bins):
    """numpy semantics (right-most edge inclusive), computed in float64 numpy from the float32 values."""
    h, e = _np.histogram(_np.asarray(asarray(x)), _np.asarray(asarray(bins)))
    return asarray(h), asarray(e)

square = _unary(lambda x: x * x)
logical_not = _unary(_t.logical_not)


def reciprocal(x):
    return _w(1.0 / asarray(x))            # lax.div(1, x)


def maximum(a, b):
    a, b = _pair(a, b)
    if is_scalar(b):
        b = _t.full_like(a, b) if a.dtype.is_floating_point or isinstance(b, int) else asarray(b)
    if is_scalar(a):
        a = _t.full_like(b, a) if b.dtype.is_floating_point or isinstance(a, int) else asarray(a)
    return _w(_t.maximum(a, b))


def minimum(a, b):
    a, b = _pair(a, b)
    if is_scalar(b):
        b = _t.full_like(a, b) if a.dtype.is_floating_point or isinstance(b, int) else asarray(b)
    if is_scalar(a):
        a = _t.full_like(b, a) if b.dtype.is_floating_point or isinstance(a, int) else asarray(a)
    return _w(_t.minimum(a, b))


def clip(x, a_min=None, a_max=None):
    """jnp.clip = minimum(maximum(x, lo), hi)."""
    x = asarray(x)
    if a_min is not None:
        x = maximum(x, a_min)
    if a_max is not None:
        x = minimum(x, a_max)
    return x


def where(c, a=None, b=None):
    c = asarray(c)
    if a is None and b is None:
        return tuple(_w(i.to(_t.int32)) for i in _t.nonzero(c, as_tuple=True))
    a, b = _pair(a, b)
    if is_scalar(a) and is_scalar(b):
        a = asarray(a)
    return _w(_t.where(c.to(_t.bool), a, b))


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    """Second positional argument is `copy` (the reference passes eps / 0 / inf there: mip.py:313,320; math.py:282)."""
    x = asarray(x)
    if not x.dtype.is_floating_point:
        return x
    return _w(_t.nan_to_num(x, nan=nan, posinf=posinf, neginf=neginf))


def power(a, b):
    a, b = _pair(a, b)
    return _w(_t.pow(a, b))


def add(a, b):
    a, b = _pair(a, b)
    return _w(a + b)


def multiply(a, b):
    a, b = _pair(a, b)
    return _w(a * b)


def any(x, axis=None):
    x = asarray(x)
    return _w(_t.any(x) if axis is None else _t.any(x, _axis(axis)))


def all(x, axis=None):
    x = asarray(x)
    return _w(_t.all(x) if axis is None else _t.all(x, _axis(axis)))


# ---------------------------------------------------------------- reductions
def sum(x, axis=None, keepdims=False, dtype=None):
    return asarray(x).sum(axis=_axis(axis), keepdims=keepdims)


def mean(x, axis=None, keepdims=False):
    return asarray(x).mean(axis=_axis(axis), keepdims=keepdims)


def prod(x, axis=None, keepdims=False):
    return asarray(x).prod(axis=_axis(axis), keepdims=keepdims)


def max(x, axis=None, keepdims=False):
    return asarray(x).max(axis=_axis(axis), keepdims=keepdims)


def min(x, axis=None, keepdims=False):
    return asarray(x).min(axis=_axis(axis), keepdims=keepdims)


amax, amin = max, min


def argmax(x, axis=None):
    x = asarray(x)
    return _w(_t.argmax(x) if axis is None else _t.argmax(x, axis)).to(_t.int32)


def cumsum(x, axis=None):
    x = asarray(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    return _w(_t.cumsum(x, axis))


def median(x, axis=None):
    x = asarray(x)
    return _w(_t.quantile(x.reshape(-1), 0.5) if axis is None else _t.quantile(x, 0.5, dim=axis))


# ---------------------------------------------------------------- shape
def reshape(x, shape):
    if isinstance(shape, int):
        shape = (shape,)
    return asarray(x).reshape(*[int(s) for s in shape])


def concatenate(xs, axis=0):
    xs = [asarray(x) for x in xs]
    dt = xs[0].dtype
    for x in xs[1:]:
        dt = _t.promote_types(dt, x.dtype)
    return _w(_t.cat([x.to(dt) for x in xs], dim=axis))


def stack(xs, axis=0):
    return _w(_t.stack([asarray(x) for x in xs], dim=axis))


def broadcast_to(x, shape):
    return _w(asarray(x).expand(*[int(s) for s in shape]))


def expand_dims(x, axis):
    x = asarray(x)
    if isinstance(axis, (tuple, list)):
        for a in sorted(axis):
            x = x.unsqueeze(a)
        return _w(x)
    return _w(x.unsqueeze(axis))


def squeeze(x, axis=None):
    return asarray(x).squeeze(axis)


def repeat(x, repeats, axis=None):
    x = asarray(x)
    if axis is None:
        return _w(_t.repeat_interleave(x.reshape(-1), int(repeats)))
    return _w(_t.repeat_interleave(x, int(repeats), dim=axis))


def tile(x, reps):
    x = asarray(x)
    reps = (reps,) if isinstance(reps, int) else tuple(reps)
    return _w(x.repeat(*reps)) if len(reps) >= x.dim() else _w(x.repeat(*((1,) * (x.dim() - len(reps)) + reps)))


def moveaxis(x, src, dst):
    return _w(_t.movedim(asarray(x), src, dst))


def swapaxes(x, a, b):
    return _w(_t.Tensor.transpose(asarray(x), a, b))


def transpose(x, axes=None):
    x = asarray(x)
    return x.transpose(*axes) if axes is not None else x.transpose()


def resize(x, new_shape):
    """numpy.resize semantics: flatten, repeat cyclically to the new size, reshape (train_boxpose.py:148)."""
    x = asarray(x).reshape(-1)
    new_shape = (new_shape,) if isinstance(new_shape, int) else tuple(int(s) for s in new_shape)
    n = 1
    for s in new_shape:
        n *= s
    if n == 0 or x.numel() == 0:
        return zeros(new_shape, x.dtype)
    reps = -(-n // x.numel())
    return _w(x.repeat(reps)[:n].reshape(*new_shape))


def pad(x, pad_width, mode='constant', constant_values=0):
    x = asarray(x)
    if isinstance(pad_width, int):
        pad_width = [(pad_width, pad_width)] * x.dim()
    pad_width = [tuple(p) if isinstance(p, (tuple, list)) else (p, p) for p in pad_width]
    out = x
    for ax, (lo, hi) in enumerate(pad_width):
        if lo == 0 and hi == 0:
            continue
        parts = []
        if mode == 'edge':
            first = out.narrow(ax, 0, 1)
            last = out.narrow(ax, out.shape[ax] - 1, 1)
            if lo:
                parts.append(_t.repeat_interleave(first, lo, dim=ax))
            parts.append(out)
            if hi:
                parts.append(_t.repeat_interleave(last, hi, dim=ax))
        elif mode == 'constant':
            shp = list(out.shape)
            if lo:
                shp[ax] = lo
                parts.append(_t.full(shp, constant_values, dtype=out.dtype))
            parts.append(out)
            if hi:
                shp[ax] = hi
                parts.append(_t.full(shp, constant_values, dtype=out.dtype))
        else:
            raise NotImplementedError(mode)
        out = _t.cat(parts, dim=ax)
    return _w(out)


def diag(x):
    return _w(_t.diag(asarray(x)))


def diagonal(x, offset=0, axis1=0, axis2=1):
    return _w(_t.diagonal(asarray(x), offset, axis1, axis2))


def matmul(a, b, precision=None):
    a, b = asarray(a), asarray(b)
    dt = _t.promote_types(a.dtype, b.dtype)
    return _w(_t.matmul(a.to(dt), b.to(dt)))


def dot(a, b, precision=None):
    return matmul(a, b)


def outer(a, b):
    return _w(_t.outer(asarray(a), asarray(b)))


def sort(x, axis=-1):
    return _w(_t.sort(asarray(x), dim=axis)[0])


def isscalar(x):
    return is_scalar(x)


def shape(x):
    return tuple(asarray(x).shape)


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None, keepdims=False):
        x = asarray(x)
        if axis is None:
            return _w(_t.sqrt(_t.sum(x * x)))
        return _w(_t.sqrt(_t.sum(x * x, dim=_axis(axis), keepdim=keepdims)))     # jnp: sqrt(sum(x*conj(x)))

    @staticmethod
    def inv(x):
        return _w(_t.linalg.inv(asarray(x)))


linalg = _Linalg()
