"""A small `flax.linen`: dataclass modules, @compact, auto-naming, parameter sharing across calls, init/apply, Dense."""
import dataclasses as _dc
import functools as _functools
from typing import Any, Callable

import torch as _t

import jax
from jax import random as _random
from jax._core import Array, asarray
from jax.nn import relu, sigmoid, softplus  # noqa: F401
from .core import FrozenDict, freeze

_frames = []      # stack of _Frame: one per executing compact method


class _Frame:
    def __init__(self, module, params, mode, rng):
        self.module, self.params, self.mode, self.rng = module, params, mode, rng
        self.counters = {}
        self.n_rng = 0

    def next_key(self):
        k = _random.fold_in(self.rng, self.n_rng)
        self.n_rng += 1
        return k


def compact(fn):
    @_functools.wraps(fn)
    def wrapper(self, *a, **k):
        if self._frame_override is not None:                     # top-level init / apply
            frame = self._frame_override
        else:
            parent = self._parent_frame
            if parent is None:
                raise RuntimeError("module called outside init/apply")
            if parent.mode == 'init':
                params = parent.params.setdefault(self.name, {})
                rng = _random.fold_in(parent.rng, hash(self.name) & 0x7fffffff)
            else:
                params = parent.params[self.name]
                rng = None
            frame = _Frame(self, params, parent.mode, rng)
        _frames.append(frame)
        try:
            return fn(self, *a, **k)
        finally:
            _frames.pop()
    wrapper._is_compact = True
    return wrapper


class Module:
    name: str = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        ann = dict(cls.__dict__.get('__annotations__', {}))
        ann.pop('name', None)
        ann['name'] = str                                        # keep `name` last, keyword with default None
        cls.__annotations__ = ann
        if 'name' not in cls.__dict__:
            cls.name = None
        _dc.dataclass(cls, eq=False, repr=False)

    def __post_init__(self):
        object.__setattr__(self, '_frame_override', None)
        parent = _frames[-1] if _frames else None
        object.__setattr__(self, '_parent_frame', parent)
        if parent is not None and self.name is None:
            cname = type(self).__name__
            i = parent.counters.get(cname, 0)
            parent.counters[cname] = i + 1
            object.__setattr__(self, 'name', f"{cname}_{i}")

    # ---- parameters
    def param(self, name, init_fn, *init_args):
        frame = _frames[-1]
        if frame.mode == 'init' and name not in frame.params:
            frame.params[name] = asarray(init_fn(frame.next_key(), *init_args))
        return asarray(frame.params[name])

    # ---- entry points
    def init(self, rngs, *a, **k):
        params = {}
        object.__setattr__(self, '_frame_override', _Frame(self, params, 'init', rngs))
        try:
            with _t.no_grad():
                self(*a, **k)
        finally:
            object.__setattr__(self, '_frame_override', None)
        return freeze({'params': params})

    def apply(self, variables, *a, rngs=None, **k):
        object.__setattr__(self, '_frame_override', _Frame(self, variables['params'], 'apply', None))
        try:
            return self(*a, **k)
        finally:
            object.__setattr__(self, '_frame_override', None)


class Dense(Module):
    """y = x @ kernel[in, features] + bias; kernel_init default lecun-normal is never used by the reference."""
    features: int
    use_bias: bool = True
    kernel_init: Callable[..., Any] = None
    bias_init: Callable[..., Any] = None

    @compact
    def __call__(self, x):
        x = asarray(x)
        kinit = self.kernel_init or jax.nn.initializers.glorot_uniform()
        kernel = self.param('kernel', kinit, (x.shape[-1], self.features))
        y = _t.matmul(x, kernel)
        if self.use_bias:
            bias = self.param('bias', self.bias_init or jax.nn.initializers.zeros, (self.features,))
            y = y + bias
        return y.as_subclass(Array)
