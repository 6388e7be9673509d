"""N3: checkpoint / resume in the reference's flax msgpack layout (train_boxpose.py:404-406, 529-532)."""
import msgpack
import numpy as np
import torch

from durf_b200 import checkpoint as ck
from durf_b200.obbpose_model import MipNerfModel, Variables
from durf_b200.train import TrainState


def _state(seed):
    model = MipNerfModel()
    v = Variables.allocate(model, 2, 5, 'cpu')
    g = torch.Generator().manual_seed(seed)
    v.flat.copy_(torch.randn(v.flat.numel(), generator=g))
    st = TrainState.create(v)
    st.m.copy_(torch.randn(v.flat.numel(), generator=g))
    st.v.copy_(torch.rand(v.flat.numel(), generator=g))
    st.step = 1234 + seed
    return st


def test_round_trip_and_flax_layout(tmp_path):
    a = _state(1)
    path = ck.save_checkpoint(str(tmp_path), a, step=a.step)
    assert path.endswith(f"checkpoint_{a.step}")
    # the file is plain msgpack with flax's ndarray extension (ExtType 1 = packb((shape, dtype, bytes)))
    raw = msgpack.unpackb(open(path, 'rb').read(), raw=False, strict_map_key=False)
    leaf = raw['optimizer']['target']['params']['MLP_0']['Dense_5']['kernel']
    assert isinstance(leaf, msgpack.ExtType) and leaf.code == 1
    shape, dtype, buf = msgpack.unpackb(leaf.data, raw=False)
    assert shape == [316, 256] and dtype == 'float32' and len(buf) == 316 * 256 * 4     # skip layer: 256 + 60 inputs
    tree = ck.from_bytes(open(path, 'rb').read())
    assert set(tree['optimizer']['target']['params']) == {'MLP_0', 'BoxMLP_0', 'BoxMLP_1', 'box_centers'}
    assert tree['optimizer']['target']['params']['box_centers'].shape == (5, 2, 6)
    assert set(tree['optimizer']['state']['param_states']['params']['MLP_0']['Dense_0']['bias']) == {'grad_ema', 'grad_sq_ema'}
    assert tree['optimizer']['state']['step'] == a.step
    b = ck.restore_checkpoint(str(tmp_path), _state(2))
    assert b.step == a.step                                       # init_step = state.optimizer.state.step + 1 in the reference
    for x, y in ((a.variables.flat, b.variables.flat), (a.m, b.m), (a.v, b.v)):
        assert torch.equal(x, y)


def test_restore_picks_latest_and_keeps_n(tmp_path):
    s = _state(3)
    for step in (10, 200, 30):
        s.step = step
        ck.save_checkpoint(str(tmp_path), s, step=step, keep=2)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ['checkpoint_200', 'checkpoint_30']           # numeric order: 10 was dropped
    fresh = ck.restore_checkpoint(str(tmp_path), _state(4))
    assert fresh.step == 200
    untouched = _state(5)
    assert ck.restore_checkpoint(str(tmp_path / 'empty'), untouched) is untouched
