"""The deprecated `flax.nn` activations the reference registers with gin (internal/utils.py:32-34)."""
from jax.nn import relu, sigmoid, softplus  # noqa: F401
