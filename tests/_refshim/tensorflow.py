"""Import-time stub for internal/math_test.py:26 (only test_ssim_golden uses it; that case is skipped)."""


def __getattr__(name):
    raise AttributeError(f"tensorflow is not installed (test-only stub); attribute {name!r} requested")
