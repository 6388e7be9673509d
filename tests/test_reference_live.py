"""Run the reference itself (unmodified files under /root/reference, on tests/_refshim) inside the CPU suite:

* the reference's own unit tests, internal/math_test.py, case by case (only test_ssim_golden is skipped: it compares
  with TensorFlow, which is not installed);
* the committed fixtures tests/golden/ref_*.npz are reproduced bit-for-bit by re-running the generator's functions;
* oracle vs live reference on inputs that are NOT in the fixtures (fresh seeds), so the oracle cannot be fitted to
  the fixtures.

Skipped where /root/reference does not exist (the GPU box).  The shim puts a stand-in `jax` on sys.path, so these
tests run in a subprocess-free but import-isolated way: the stand-in is only ever imported through refshim_loader.
"""
import os
import unittest

import numpy as np
import pytest
import torch

import durf_test_helpers as H
import ref_cases as C
import refshim_loader as L
from oracle import durf_oracle as O

pytestmark = pytest.mark.skipif(not L.reference_available(), reason="reference checkout not present")

MATH_TEST_CASES = [
    'test_sin', 'test_psnr_round_trip', 'test_learning_rate_decay', 'test_delayed_learning_rate_decay',
    'test_ssim_lowerbound', 'test_srgb_linearize', 'test_sorted_piecewise_constant_pdf_train_mode',
    'test_sorted_piecewise_constant_pdf_large_flat', 'test_sorted_piecewise_constant_pdf_sparse_delta',
    'test_sorted_piecewise_constant_pdf_single_bin',
]


@pytest.mark.parametrize("case", MATH_TEST_CASES)
def test_reference_math_test_py(case):
    """internal/math_test.py:<case>, the reference's own test body, against the reference's own internal/math.py."""
    L.load_reference()
    import importlib
    mt = importlib.import_module('internal.math_test')
    result = unittest.TestResult()
    mt.MathUtilsTest(case).run(result)
    problems = result.errors + result.failures
    assert not problems, problems[0][1]
    assert result.testsRun == 1


def test_math_test_py_case_list_is_complete():
    L.load_reference()
    import importlib
    mt = importlib.import_module('internal.math_test')
    have = sorted(n for n in dir(mt.MathUtilsTest) if n.startswith('test_'))
    assert have == sorted(MATH_TEST_CASES + ['test_ssim_golden'])


@pytest.mark.parametrize("fixture", ['math', 'obb', 'mip', 'model', 'train'])
def test_fixtures_regenerate_bit_for_bit(fixture):
    """tests/golden/ref_<fixture>.npz == what the generator produces now from /root/reference."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_ref_golden', os.path.join(os.path.dirname(__file__), 'golden', 'make_ref_golden.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    data = getattr(mod.Gen(), fixture)()
    old = np.load(os.path.join(os.path.dirname(__file__), 'golden', f'ref_{fixture}.npz'))
    assert sorted(old.files) == sorted(data.keys())
    for k, v in data.items():
        a, b = np.asarray(v), old[k]
        if a.dtype.kind in 'US':
            assert str(a) == str(b), k
        else:
            assert a.shape == b.shape and np.array_equal(a, b, equal_nan=a.dtype.kind == 'f'), k


def _T(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a)))


@pytest.mark.parametrize("seed", [1001, 1002])
def test_oracle_vs_live_reference_fresh_inputs(seed):
    """Forward of the full model on a scene that is in no fixture: reference (live) vs oracle."""
    ref = L.load_reference()
    jax, jnp, flax, gin = ref.jax, ref.jnp, ref.flax, ref.gin
    sc = C._scene(B=64, K=3, seed=seed)
    gin.clear_config()
    gin.parse_config_file(os.path.join(ref.root, 'configs', 'carla_dyn.gin'))
    gin.bind_parameter('MipNerfModel.num_objects', 3)
    model = ref.obbpose_model.MipNerfModel()

    def tree(layers):
        return {f'Dense_{i}': {'kernel': jnp.array(k), 'bias': jnp.array(b)} for i, (k, b) in enumerate(layers)}
    p = {'MLP_0': tree(sc['mlp']), 'box_centers': jnp.array(sc['centers'])}
    for k, m in enumerate(sc['box_mlps']):
        p[f'BoxMLP_{k}'] = tree(m)
    V = flax.core.freeze({'params': p})
    rays = ref.utils.BoxRays(*[jnp.array(a) for a in sc['rays']])
    bufs = [sc['t_rand'], sc['u_rand']]
    jax.random.set_provider(lambda kind, path, shape: bufs.pop(0))
    try:
        ret = model.apply(V, jax.random.PRNGKey(seed), rays, jnp.array(sc['centers']), jnp.array(sc['ext']),
                          jnp.array(np.array([4], np.int32)), randomized=True, rand_bkgd=False, white_bkgd=False, alpha=5.5)
    finally:
        jax.random.set_provider(None)
    want = O.model_forward(H.oracle_params(sc), H.oracle_rays(sc), _T(sc['ext']), 4, True, False, False, 5.5,
                           t_rand=_T(sc['t_rand']), u_rand=_T(sc['u_rand']))
    for lvl in range(2):
        rt = 3e-6 if lvl == 0 else 2e-4
        for i, nm in enumerate(('comp_rgb', 'distance', 'acc', 'weights', 't_vals')):
            a, b = _T(ret[lvl][i].detach()).double(), want[lvl][i].double()
            tol = rt * torch.clamp(b.abs(), min=1.0 if nm != 'weights' else 0.1)
            assert bool(((a - b).abs() <= tol).all()), f"L{lvl} {nm}: worst {float(((a - b).abs() - tol).max()):.3e}"
