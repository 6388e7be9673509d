from . import tensorboard  # noqa: F401
