"""Import-time stub (internal/obbpose_dataset.py:12)."""


def natsorted(xs, **k):
    return sorted(xs, **k)
