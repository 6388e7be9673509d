"""Checkpoint / resume of the train state in the reference's on-disk format (N3).

The reference saves `flax.training.checkpoints.save_checkpoint(train_dir, state, step, keep=100)` and resumes with
`restore_checkpoint(train_dir, state)`; `init_step = state.optimizer.state.step + 1` (train_boxpose.py:404-406, 529-532,
578-581).  A flax checkpoint is the msgpack encoding of the state dict of `utils.TrainState(optimizer=flax.optim.Optimizer)`
(internal/utils.py:37-40):

    {'optimizer': {'target': {'params': {'MLP_0': {'Dense_i': {'kernel', 'bias'}}, 'BoxMLP_k': {...}, 'box_centers'}},
                   'state':  {'step': int,
                              'param_states': {'params': {<same tree>: {'grad_ema', 'grad_sq_ema'}}}}}}

with every ndarray stored as msgpack ExtType(1, packb((shape, dtype.name, bytes))) (flax.serialization, pinned flax>=0.2.2,
requirements_jax.txt:4).  flax itself is not installed here, so the encoding is restated from its published format; files
written here load with flax's `restore_checkpoint`, and reference checkpoints load into `Variables` / Adam moments.
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, Optional

import msgpack
import numpy as np
import torch

_EXT_NDARRAY = 1
_EXT_NPSCALAR = 3


def _encode_ext(obj):
    if isinstance(obj, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, msgpack.packb((list(obj.shape), obj.dtype.name, obj.tobytes('C')), use_bin_type=True))
    if isinstance(obj, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, msgpack.packb(((), obj.dtype.name, obj.tobytes()), use_bin_type=True))
    raise TypeError(f"cannot serialise {type(obj)}")


def _decode_ext(code, data):
    if code in (_EXT_NDARRAY, _EXT_NPSCALAR):
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False)
        arr = np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape)
        return arr.copy() if code == _EXT_NDARRAY else arr.reshape(()).item()
    return msgpack.ExtType(code, data)


def to_bytes(tree: Dict[str, Any]) -> bytes:
    return msgpack.packb(tree, default=_encode_ext, use_bin_type=True, strict_types=True)


def from_bytes(data: bytes) -> Dict[str, Any]:
    return msgpack.unpackb(data, ext_hook=_decode_ext, raw=False, strict_map_key=False)


def _param_tree(variables, flat: torch.Tensor) -> Dict[str, Any]:
    tree: Dict[str, Any] = {}
    for name in variables.topo:
        tree[name] = {f'Dense_{i}': {'kernel': w.detach().cpu().numpy().copy(), 'bias': b.detach().cpu().numpy().copy()}
                      for i, (w, b) in enumerate(variables.layers(name, flat))}
    tree['box_centers'] = variables.view_of(flat, 'box_centers').detach().cpu().numpy().copy()
    return tree


def state_dict(state) -> Dict[str, Any]:
    """TrainState -> the reference's flax state dict (Adam: grad_ema = m, grad_sq_ema = v)."""
    v = state.variables
    m, s = _param_tree(v, state.m), _param_tree(v, state.v)

    def moments(a, b):
        if isinstance(a, dict):
            return {k: moments(a[k], b[k]) for k in a}
        return {'grad_ema': a, 'grad_sq_ema': b}

    return {'optimizer': {'target': {'params': _param_tree(v, v.flat)},
                          'state': {'step': int(state.step), 'param_states': {'params': moments(m, s)}}}}


def load_state_dict(state, tree: Dict[str, Any]):
    """Copies a (reference or own) state dict into `state` in place; returns it."""
    v = state.variables
    opt = tree['optimizer']
    params = opt['target']['params']
    ps = opt['state']['param_states']['params']

    def put(flat, name, get):
        for i, (w, b) in enumerate(v.layers(name, flat)):
            d = get(name, f'Dense_{i}')
            w.copy_(torch.from_numpy(np.asarray(d[0], np.float32)).to(flat.device).reshape(w.shape))
            b.copy_(torch.from_numpy(np.asarray(d[1], np.float32)).to(flat.device).reshape(b.shape))

    for name in v.topo:
        put(v.flat, name, lambda n, d: (params[n][d]['kernel'], params[n][d]['bias']))
        put(state.m, name, lambda n, d: (ps[n][d]['kernel']['grad_ema'], ps[n][d]['bias']['grad_ema']))
        put(state.v, name, lambda n, d: (ps[n][d]['kernel']['grad_sq_ema'], ps[n][d]['bias']['grad_sq_ema']))
    bc = lambda a: torch.from_numpy(np.asarray(a, np.float32)).to(v.flat.device).reshape(v.T, v.K, 6)
    v.box_centers.copy_(bc(params['box_centers']))
    v.view_of(state.m, 'box_centers').copy_(bc(ps['box_centers']['grad_ema']))
    v.view_of(state.v, 'box_centers').copy_(bc(ps['box_centers']['grad_sq_ema']))
    state.step = int(opt['state']['step'])
    v.mark_dirty()
    return state


def _checkpoint_files(train_dir: str, prefix: str):
    out = []
    if os.path.isdir(train_dir):
        for f in os.listdir(train_dir):
            m = re.fullmatch(re.escape(prefix) + r'(\d+)', f)
            if m:
                out.append((int(m.group(1)), os.path.join(train_dir, f)))
    return sorted(out)


def save_checkpoint(train_dir: str, state, step: int, prefix: str = 'checkpoint_', keep: int = 100) -> str:
    """flax.training.checkpoints.save_checkpoint: writes `<train_dir>/<prefix><step>` atomically, keeps the newest `keep`."""
    os.makedirs(train_dir, exist_ok=True)
    path = os.path.join(train_dir, f'{prefix}{step}')
    tmp = path + '.tmp'
    with open(tmp, 'wb') as f:
        f.write(to_bytes(state_dict(state)))
    os.replace(tmp, path)
    files = _checkpoint_files(train_dir, prefix)
    for _, old in files[:-keep] if keep > 0 else []:
        os.remove(old)
    return path


def restore_checkpoint(train_dir: str, state, prefix: str = 'checkpoint_', step: Optional[int] = None):
    """flax.training.checkpoints.restore_checkpoint: loads the newest (or the given) checkpoint into `state`; returns `state`
    unchanged when the directory holds none (what the reference relies on for a fresh run, train_boxpose.py:404)."""
    files = _checkpoint_files(train_dir, prefix)
    if step is not None:
        files = [f for f in files if f[0] == step]
    if not files:
        return state
    with open(files[-1][1], 'rb') as f:
        return load_state_dict(state, from_bytes(f.read()))
