"""N4: durf_b200.obbpose_dataset.Carla against the reference's own loader.

`test_loader_equals_the_live_reference` (only where /root/reference exists) runs internal/obbpose_dataset.py unmodified on the
test-only jax / gin stand-in over a synthetic on-disk scene and compares every array of the first training batches and of the
held-out frames, for the shipped CARLA configuration, with box / yaw noise enabled, for the Waymo loader (configs/waymo.gin, with and
without its box noise) and for the single-camera sequence loader.  The same batches are committed as
tests/golden/ref_dataset.npz (written by this test's generator mode: `python tests/test_dataset.py --write`) so that the
comparison also runs where the reference is absent."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dataset_fixture as F
import refshim_loader as RL

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_dataset.npz')
# case -> (loader, Config overrides on top of the loader's shipped .gin)
CASES = {'shipped': ('carla_dyn', dict()), 'noisy': ('carla_dyn', dict(random_box=True, random_yaw=True)),
         'waymo': ('waymo', dict(random_box=False)), 'waymo_shipped': ('waymo', dict()),      # configs/waymo.gin: random_box = True
         'seq': ('carla_seq', dict(dataset_loader='carla_seq', llffhold=4))}                  # no .gin ships for it: carla_dyn.gin + loader
GIN = {'carla_dyn': 'carla_dyn.gin', 'carla_seq': 'carla_dyn.gin', 'waymo': 'waymo.gin'}
# values of the reference's .gin files that differ from durf_b200.utils.Config's defaults (= configs/carla_dyn.gin)
OURS_BASE = {'carla_dyn': dict(), 'carla_seq': dict(), 'waymo': dict(dataset_loader='waymo', far=40.0, random_box=True)}


def _scene(path, loader):
    if loader == 'carla_seq':
        return F.make_scene(path, T=13, cams=1)
    return F.make_scene(path, waymo=loader == 'waymo')
N_TRAIN, N_TEST = 3, 2


def _splits(overrides):
    """With box / yaw noise only the training split is reproducible in the reference: `_train_init` seeds the global numpy
    generator before the noise is drawn (:208), `_test_init` does not."""
    return (('train', N_TRAIN),) if overrides.get('random_box') else (('train', N_TRAIN), ('test', N_TEST), ('render', 1))


def _flatten(prefix, batch, out):
    for k, v in batch.items():
        if k == 'rays':
            for name, r in zip(('origins', 'directions', 'viewdirs', 'radii', 'lossmult', 'near', 'far'), v):
                out[f'{prefix}/rays/{name}'] = np.asarray(r)
        else:
            out[f'{prefix}/{k}'] = np.asarray(v)


def _ours(root, loader, overrides):
    from durf_b200.obbpose_dataset import get_dataset
    from durf_b200.utils import Config
    kw = dict(OURS_BASE[loader], **overrides)
    out = {}
    for split, n in _splits(kw):
        ds = get_dataset(split, root, Config(**kw))
        first = ds.peek()
        for i in range(n):
            b = next(ds)
            if i == 0:
                assert b is first, "peek() must return the batch the next __next__ yields"
            _flatten(f'{split}{i}', b, out)
    return out


def _reference(root, loader, overrides):
    """The reference's loader class (a thread filling a queue from the GLOBAL numpy generator), run live on the stand-in."""
    import importlib
    ref = RL.load_reference()
    ds_mod = importlib.import_module('internal.obbpose_dataset')
    ref.gin.parse_config_file(os.path.join(RL.REFERENCE_ROOT, 'configs', GIN[loader]))
    cfg = ref.utils.Config()
    for k, v in overrides.items():
        setattr(cfg, k, v)
    kw = dict(OURS_BASE[loader], **overrides)
    out = {}
    for split, n in _splits(kw):
        ds = ds_mod.dataset_dict[loader](split, root, cfg)
        for i in range(n):
            _flatten(f'{split}{i}', ds.queue.get(timeout=60), out)
        # the loader thread keeps drawing from np.random; it is a daemon thread and dies with the process
    return out


def _compare(got, want, what):
    assert set(got) == set(want), f"{what}: keys differ: {sorted(set(got) ^ set(want))}"
    for k in sorted(want):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, f"{what} {k}: shape {a.shape} vs {b.shape}"
        if b.dtype.kind in 'iu':
            assert np.array_equal(a, b), f"{what} {k}"
        else:
            np.testing.assert_allclose(a.astype(np.float64), b.astype(np.float64), rtol=2e-6, atol=2e-6, err_msg=f"{what} {k}")


@pytest.mark.parametrize("case", sorted(CASES))
def test_loader_equals_the_committed_reference_batches(case, tmp_path):
    loader, overrides = CASES[case]
    root = _scene(str(tmp_path / 'scene'), loader)
    want = {k[len(case) + 1:]: v for k, v in np.load(GOLDEN).items() if k.startswith(case + '/')}
    _compare(_ours(root, loader, overrides), want, case)


@pytest.mark.skipif(not RL.reference_available(), reason="needs /root/reference (absent on the GPU box)")
def test_loader_equals_the_live_reference(tmp_path):
    for case, (loader, overrides) in CASES.items():
        root = _scene(str(tmp_path / case), loader)
        _compare(_ours(root, loader, overrides), _reference(root, loader, overrides), case)


def test_loader_contract(tmp_path):
    """Shapes / invariants the train step relies on, and the refusals."""
    from durf_b200.obbpose_dataset import get_dataset
    from durf_b200.utils import Config
    root = F.make_scene(str(tmp_path / 'scene'))
    ds = get_dataset('train', root, Config(batch_size=64))
    b = next(ds)
    assert b['pixels'].shape == (64, 3) and b['depth'].shape == (64, 1) and b['sky'].shape == (64, 1)
    assert all(r.shape[0] == 64 for r in b['rays']) and b['rays'].origins.dtype == np.float32
    assert b['init'].shape == (3, 2, 6) and b['ext'].shape == (2, 3) and 0 <= int(b['ts']) < 3
    assert np.allclose(np.linalg.norm(b['rays'].viewdirs, axis=-1), 1.0, atol=1e-5)
    assert set(np.unique(b['sky'])) <= {0.0, np.float32(0.995)}
    assert ds.size == 13 and get_dataset('test', root, Config()).size == 2          # i_test = [10, 11]
    with pytest.raises(NotImplementedError):
        get_dataset('train', root, Config(spherify=False))
    with pytest.raises(NotImplementedError):
        get_dataset('train', root, Config(dataset_loader='llff'))
    wroot = F.make_scene(str(tmp_path / 'wscene'), waymo=True)
    w = get_dataset('render', wroot, Config(dataset_loader='waymo', far=40.0))
    assert w.size == 15 and w.camera(0)['principal_point'] is not None and get_dataset('test', wroot, Config(dataset_loader='waymo')).size == 2
    with pytest.raises(ValueError):
        get_dataset('train', str(tmp_path / 'missing'), Config())


if __name__ == '__main__' and '--write' in sys.argv:
    import tempfile
    assert RL.reference_available(), "the generator needs /root/reference"
    with tempfile.TemporaryDirectory() as d:
        blob = {}
        for case, (loader, overrides) in CASES.items():
            root = _scene(os.path.join(d, case), loader)
            for k, v in _reference(root, loader, overrides).items():
                blob[f'{case}/{k}'] = v
    np.savez_compressed(GOLDEN, **blob)
    print('wrote', GOLDEN, len(blob), 'arrays', os.path.getsize(GOLDEN), 'bytes')
