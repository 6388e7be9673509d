"""Array type of the test-only jax shim: a torch.Tensor subclass with jnp method semantics.  See ../README.md."""
import numpy as _np
import torch

torch.set_default_dtype(torch.float32)

_DTYPE_NAMES = {
    'float32': torch.float32, 'float64': torch.float32,      # x64 is disabled in the reference's JAX
    'float16': torch.float16, 'bfloat16': torch.bfloat16,
    'int32': torch.int32, 'int64': torch.int32, 'uint8': torch.uint8, 'int8': torch.int8, 'bool': torch.bool,
}


def to_torch_dtype(dt):
    if dt is None:
        return None
    if isinstance(dt, torch.dtype):
        return {torch.float64: torch.float32, torch.int64: torch.int32}.get(dt, dt)
    if dt is float:
        return torch.float32
    if dt is int:
        return torch.int32
    if dt is bool:
        return torch.bool
    return _DTYPE_NAMES[_np.dtype(dt).name]


def _integer_pow(x, n):
    """lax.integer_pow: square-and-multiply (what `x ** <python int>` lowers to in JAX)."""
    if n == 0:
        return torch.ones_like(x)
    neg = n < 0
    n = abs(n)
    acc = None
    base = x
    while n > 0:
        if n & 1:
            acc = base if acc is None else acc * base
        n >>= 1
        if n:
            base = base * base
    return 1.0 / acc if neg else acc


class _AxisArgs:
    @staticmethod
    def norm(axis):
        if isinstance(axis, list):
            axis = tuple(axis)
        return axis


class Array(torch.Tensor):
    """float32 / int32 / bool array with the method surface the reference uses on jnp arrays."""

    __array_priority__ = 1000

    # ---- immutability: `a += b` rebinds instead of writing through (jnp arrays are immutable) ----
    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def __setitem__(self, k, v):
        raise TypeError("jax arrays are immutable")

    # ---- operators whose jnp semantics differ from torch's ----
    def __pow__(self, e):
        if isinstance(e, int) and not isinstance(e, bool):
            return _integer_pow(self, e)
        return torch.Tensor.__pow__(self, e)

    def __mod__(self, o):
        return torch.remainder(self, o)          # floored remainder built on an exact fmod, like lax.rem + sign fix

    def __matmul__(self, o):
        return torch.matmul(self, asarray(o))

    # ---- reductions / shape methods with numpy keywords ----
    def _red(self, fn, axis, keepdims):
        axis = _AxisArgs.norm(axis)
        if axis is None:
            out = fn(torch.Tensor.reshape(self, (-1,)), 0)
            out = out[0] if isinstance(out, tuple) else out
            return out.reshape([1] * self.dim()) if keepdims else out
        if isinstance(axis, tuple):
            out = self
            for a in sorted([a % self.dim() for a in axis], reverse=True):
                out = fn(out, a)
                out = out[0] if isinstance(out, tuple) else out
                if keepdims:
                    out = out.unsqueeze(a)
            return out
        out = fn(self, axis)
        out = out[0] if isinstance(out, tuple) else out
        return out.unsqueeze(axis) if keepdims else out

    def sum(self, axis=None, dtype=None, keepdims=False, out=None, **_):
        x = self
        if x.dtype == torch.bool:
            x = x.to(torch.int32)
        return x._red(lambda t, a: torch.sum(t, a), axis, keepdims)

    def mean(self, axis=None, dtype=None, keepdims=False, out=None, **_):
        x = self if self.dtype.is_floating_point else self.to(torch.float32)
        return x._red(lambda t, a: torch.mean(t, a), axis, keepdims)

    def prod(self, axis=None, dtype=None, keepdims=False, out=None, **_):
        return self._red(lambda t, a: torch.prod(t, a), axis, keepdims)

    def max(self, axis=None, keepdims=False, out=None, **_):
        return self._red(lambda t, a: torch.max(t, a), axis, keepdims)

    def min(self, axis=None, keepdims=False, out=None, **_):
        return self._red(lambda t, a: torch.min(t, a), axis, keepdims)

    def all(self, axis=None, keepdims=False, out=None, **_):
        return self.to(torch.bool)._red(lambda t, a: torch.all(t, a), axis, keepdims)

    def any(self, axis=None, keepdims=False, out=None, **_):
        return self.to(torch.bool)._red(lambda t, a: torch.any(t, a), axis, keepdims)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        """numpy ufuncs applied to (or mixed with) an Array compute in numpy and return numpy: `np.abs(a)`,
        `ndarray += a`, `ndarray / a` in the reference's tests."""
        conv = [_np.asarray(x) if isinstance(x, torch.Tensor) else x for x in inputs]
        if out is not None:
            kw['out'] = tuple(_np.asarray(o) if isinstance(o, torch.Tensor) else o for o in out)
        return getattr(ufunc, method)(*conv, **kw)

    def astype(self, dt):
        return self.to(to_torch_dtype(dt))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple, torch.Size)):
            shape = tuple(shape[0])
        return torch.Tensor.reshape(self, tuple(int(s) for s in shape))

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (list, tuple)):
            axes = tuple(axes[0])
        if len(axes) == 0:
            axes = tuple(reversed(range(self.dim())))
        if len(axes) == 2 and self.dim() > 2:      # torch-internal callers use the (dim0, dim1) swap form
            return torch.Tensor.transpose(self, *axes)
        return self.permute(*axes)

    @property
    def T(self):
        return self.permute(*reversed(range(self.dim())))

    def squeeze(self, axis=None):
        if axis is None:
            return torch.Tensor.squeeze(self)
        return torch.Tensor.squeeze(self, axis)

    def ravel(self):
        return torch.Tensor.reshape(self, (-1,))

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return "Array(" + repr(self.detach().as_subclass(torch.Tensor)) + ")"

    def __format__(self, spec):
        if self.numel() == 1:
            return format(self.item(), spec)
        return repr(self)

    def __index__(self):
        return int(self.item())

    def __hash__(self):
        return id(self)


def asarray(x, dtype=None):
    """numpy / python / torch value -> Array with JAX's default (x64-disabled) dtypes."""
    dt = to_torch_dtype(dtype)
    if isinstance(x, torch.Tensor):
        out = x if isinstance(x, Array) else x.as_subclass(Array)
        if out.dtype == torch.float64:
            out = out.to(torch.float32)
        elif out.dtype == torch.int64:
            out = out.to(torch.int32)
        return out.to(dt) if dt is not None and out.dtype != dt else out
    if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) or
                                            (isinstance(e, (list, tuple)) and any(isinstance(f, torch.Tensor) for f in e))
                                            for e in x):
        out = torch.stack([asarray(e) for e in x]).as_subclass(Array)
        return out.to(dt) if dt is not None else out
    a = _np.asarray(x)
    if a.dtype == _np.float64:
        a = a.astype(_np.float32)
    elif a.dtype == _np.int64:
        a = a.astype(_np.int32)
    elif a.dtype == object:
        raise TypeError(f"cannot convert {type(x)} to an array")
    out = torch.from_numpy(_np.ascontiguousarray(a)).clone().as_subclass(Array)
    return out.to(dt) if dt is not None else out


def is_scalar(x):
    return isinstance(x, (int, float, bool)) and not isinstance(x, torch.Tensor)


def lift(x):
    """Keep python scalars weak (torch treats them like jnp's weak types); convert everything else."""
    return x if is_scalar(x) else asarray(x)
