"""jax.test_util.JaxTestCase stand-in for internal/math_test.py."""
import numpy as _np
from absl.testing import absltest


class JaxTestCase(absltest.TestCase):
    def assertAllClose(self, x, y, atol=None, rtol=None, check_dtypes=False):
        x, y = _np.asarray(x, dtype=_np.float64), _np.asarray(y, dtype=_np.float64)
        tol = 1e-6          # jax's default tolerance for float32
        _np.testing.assert_allclose(x, y, atol=tol if atol is None else atol, rtol=tol if rtol is None else rtol)

    def assertArraysEqual(self, x, y):
        _np.testing.assert_array_equal(_np.asarray(x), _np.asarray(y))
