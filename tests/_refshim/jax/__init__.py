"""Test-only stand-in for the `jax` package (see ../README.md).  NOT JAX: torch-CPU float32 underneath."""
import functools as _functools

import torch as _t

from ._core import Array, asarray
from . import numpy, lax, random, tree_util, nn, scipy  # noqa: F401
from .tree_util import tree_map, tree_multimap  # noqa: F401

__version__ = '0.0-refshim'
IS_REFSHIM = True


class _Config:
    def parse_flags_with_absl(self):
        pass

    def update(self, *a, **k):
        pass


config = _Config()


def device_count():
    return 1


def local_device_count():
    return 1


def host_id():
    return 0


def process_index():
    return 0


def host_count():
    return 1


def devices():
    return ['cpu:0']


def device_get(x):
    return x


def device_put(x, device=None):
    return x


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


grad_tap = []


def _leaves_requiring_grad(tree):
    leaves, treedef = tree_util.tree_flatten(tree)
    new = [asarray(l).detach().clone().requires_grad_(True) if isinstance(l, _t.Tensor) and l.dtype.is_floating_point
           else l for l in leaves]
    return new, tree_util.tree_unflatten(treedef, new), treedef


def value_and_grad(fn, argnums=0, has_aux=False):
    """Reverse-mode gradient of a scalar function of a pytree, through torch autograd."""
    def wrapped(*args, **kw):
        leaves, tree, treedef = _leaves_requiring_grad(args[argnums])
        args = list(args)
        args[argnums] = tree
        out = fn(*args, **kw)
        val, aux = out if has_aux else (out, None)
        diff = [l for l in leaves if isinstance(l, _t.Tensor) and l.requires_grad]
        gs = _t.autograd.grad(val, diff, allow_unused=True)
        it = iter(gs)
        grads = []
        for l in leaves:
            if isinstance(l, _t.Tensor) and l.requires_grad:
                g = next(it)
                grads.append((_t.zeros_like(l) if g is None else g).detach().as_subclass(Array))
            else:
                grads.append(None)
        gtree = tree_util.tree_unflatten(treedef, grads)
        grad_tap.append(gtree)          # the reference's train_step does not return its gradients: tests read them here
        if has_aux:
            aux = tree_util.tree_map(lambda v: v.detach() if isinstance(v, _t.Tensor) else v, aux)
            return (val.detach(), aux), gtree
        return val.detach(), gtree
    return wrapped


def grad(fn, argnums=0, has_aux=False):
    vg = value_and_grad(fn, argnums, has_aux)

    def wrapped(*a, **k):
        v, g = vg(*a, **k)
        return (g, v[1]) if has_aux else g
    return wrapped


def linearize(fn, x):
    """(fn(x), jvp) with jvp(t) = J_fn(x) t, by the double-backward identity d/dv <J^T v, t>; differentiable in x."""
    x = asarray(x)
    with _t.enable_grad():
        xin = x if x.requires_grad else x.detach().clone().requires_grad_(True)
        y = fn(xin)

    def jvp(t):
        t = asarray(t)
        with _t.enable_grad():
            v = _t.zeros_like(y, requires_grad=True)
            (jt_v,) = _t.autograd.grad(y, xin, v, create_graph=True)
            (out,) = _t.autograd.grad(jt_v, v, t, create_graph=x.requires_grad)
        return out.as_subclass(Array) if x.requires_grad else out.detach().as_subclass(Array)

    return (y if x.requires_grad else y.detach()), jvp


def jvp(fn, primals, tangents):
    y, lin = linearize(fn, primals[0])
    return y, lin(tangents[0])


def vmap(fn, in_axes=0, out_axes=0):
    """Map over one axis of a single array argument by an explicit loop (math.py:111-112)."""
    def wrapped(x):
        x = asarray(x)
        outs = [fn(xi) for xi in _t.unbind(x, dim=in_axes)]
        return _t.stack(outs, dim=out_axes).as_subclass(Array)
    return wrapped


def pmap(fn, axis_name=None, in_axes=0, out_axes=0, donate_argnums=(), static_broadcasted_argnums=()):
    """One device: strip the leading (size-1) device axis of mapped arguments, call, put it back."""
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        call = []
        for a, ax in zip(args, axes):
            call.append(a if ax is None else tree_util.tree_map(lambda v: v[0], a))
        out = fn(*call)
        return tree_util.tree_map(lambda v: asarray(v)[None] if isinstance(v, _t.Tensor) else v, out)
    return wrapped
