class FrozenDict(dict):
    """Immutable-by-convention dict; a pytree node like any dict (keys sorted)."""

    def __setitem__(self, k, v):
        raise TypeError("FrozenDict is immutable")

    def unfreeze(self):
        return unfreeze(self)


def freeze(d):
    return FrozenDict({k: freeze(v) if isinstance(v, dict) else v for k, v in d.items()})


def unfreeze(d):
    return {k: unfreeze(v) if isinstance(v, dict) else v for k, v in d.items()}
