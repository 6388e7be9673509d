"""Test-only stand-in for the parts of `flax` the reference touches (see ../README.md)."""
from . import core, struct, nn, optim, linen, jax_utils  # noqa: F401
from .core import FrozenDict, freeze, unfreeze  # noqa: F401

__version__ = '0.0-refshim'
