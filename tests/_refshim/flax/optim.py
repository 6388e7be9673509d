"""`flax.optim.Adam` (the pre-optax optimizer API the reference uses: train_boxpose.py:288,343).

flax is a third-party dependency absent from /root/reference and from this image (requirements_jax.txt:4 pins
flax>=0.2.2).  Its published Adam update (flax/optim/adam.py, `apply_param_gradient`) is restated here:

    grad_ema    = b1 * grad_ema + (1 - b1) * g
    grad_sq_ema = b2 * grad_sq_ema + (1 - b2) * g^2
    t = step + 1
    p <- p - lr * (grad_ema / (1 - b1^t)) / (sqrt(grad_sq_ema / (1 - b2^t)) + eps)  - lr * weight_decay * p
"""
import dataclasses as _dc

import torch as _t

from jax import tree_util as _tu
from jax._core import asarray
from . import struct as _struct


@_struct.dataclass
class _AdamParamState:
    grad_ema: object
    grad_sq_ema: object


@_struct.dataclass
class OptimizerState:
    step: object
    param_states: object


@_struct.dataclass
class Optimizer:
    optimizer_def: object
    state: object
    target: object

    def apply_gradient(self, grads, **hyper):
        target, state = self.optimizer_def.apply_gradient(hyper, self.target, self.state, grads)
        return self.replace(target=target, state=state)


class Adam:
    def __init__(self, learning_rate=None, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
        self.learning_rate, self.beta1, self.beta2, self.eps, self.weight_decay = learning_rate, beta1, beta2, eps, weight_decay

    def create(self, target):
        ps = _tu.tree_map(lambda p: _AdamParamState(_t.zeros_like(asarray(p)), _t.zeros_like(asarray(p))), target)
        return Optimizer(self, OptimizerState(asarray(0), ps), target)

    def apply_gradient(self, hyper, params, state, grads):
        lr = hyper.get('learning_rate', self.learning_rate)
        b1, b2, eps, wd = self.beta1, self.beta2, self.eps, self.weight_decay
        step = state.step
        t = asarray(step).to(_t.float32) + 1.0
        p_leaves, treedef = _tu.tree_flatten(params)
        g_leaves = _tu.tree_leaves(grads)
        s_leaves = [s for s in _flatten_states(state.param_states, len(p_leaves))]
        new_p, new_s = [], []
        for p, g, s in zip(p_leaves, g_leaves, s_leaves):
            p, g = asarray(p), asarray(g)
            grad_sq = g * g
            grad_ema = b1 * s.grad_ema + (1.0 - b1) * g
            grad_sq_ema = b2 * s.grad_sq_ema + (1.0 - b2) * grad_sq
            grad_ema_corr = grad_ema / (1 - b1 ** t)
            grad_sq_ema_corr = grad_sq_ema / (1 - b2 ** t)
            denom = _t.sqrt(grad_sq_ema_corr) + eps
            np_ = p - lr * grad_ema_corr / denom
            np_ = np_ - lr * wd * p
            new_p.append(np_)
            new_s.append(_AdamParamState(grad_ema, grad_sq_ema))
        new_params = _tu.tree_unflatten(treedef, new_p)
        new_states = _tu.tree_unflatten(treedef, new_s)
        return new_params, OptimizerState(step + 1, new_states)


def _flatten_states(tree, n):
    out = []

    def rec(node):
        if isinstance(node, _AdamParamState):
            out.append(node)
        elif isinstance(node, dict):
            for k in sorted(node.keys()):
                rec(node[k])
        elif isinstance(node, (list, tuple)):
            for v in node:
                rec(v)
    rec(tree)
    assert len(out) == n
    return out


# Optimizer is deliberately NOT a pytree with a static definition slot: tree_map over a TrainState maps over
# (optimizer_def -> leaf object, state, target); the definition object passes through tree_map functions that only
# touch tensors (jax_utils.replicate / device_get in the reference's main loop are not exercised by the tests).
