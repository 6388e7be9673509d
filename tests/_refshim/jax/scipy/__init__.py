from . import signal  # noqa: F401
