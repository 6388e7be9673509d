"""Import the reference's own source files, unmodified, on top of tests/_refshim (test-only; see tests/_refshim/README.md).

    ref = load_reference()          # None when /root/reference is absent (e.g. on the GPU box)
    ref.mip.cast_rays(...)          # -> internal/mip.py:155, the reference's code, executing on torch-CPU float32

Nothing under durf_b200/ imports this module.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('DURF_REFERENCE_ROOT', '/root/reference')
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_refshim')

_cached = None


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'internal', 'mip.py'))


def load_reference(with_train: bool = True):
    """-> namespace(math, mip, mip360, box_helpers, utils, obbpose_model, train_boxpose, jax, jnp, flax, gin) or None."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        return None
    for p in (REFERENCE_ROOT, SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import jax
    assert getattr(jax, 'IS_REFSHIM', False), "a real jax is importable: use it instead of the shim"
    import jax.numpy as jnp
    import flax
    import gin
    ns = types.SimpleNamespace(jax=jax, jnp=jnp, flax=flax, gin=gin, root=REFERENCE_ROOT)
    for name in ('math', 'mip', 'mip360', 'box_helpers', 'utils', 'obbpose_model'):
        setattr(ns, name, importlib.import_module('internal.' + name))
    if with_train:
        # train_boxpose.py defines absl flags and imports the dataset / visualisation modules at import time; they are
        # importable on the shim (cv2, PIL, scipy are installed; natsort / matplotlib are import-time stubs).
        ns.train_boxpose = importlib.import_module('train_boxpose')
    _cached = ns
    return ns


def source_sha256(relpath: str) -> str:
    import hashlib
    with open(os.path.join(REFERENCE_ROOT, relpath), 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()
