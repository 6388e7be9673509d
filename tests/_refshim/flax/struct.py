import dataclasses as _dc

from jax import tree_util as _tu


def dataclass(cls):
    """flax.struct.dataclass: frozen dataclass + `.replace` + pytree registration."""
    cls = _dc.dataclass(cls)

    def replace(self, **kw):
        return _dc.replace(self, **kw)

    cls.replace = replace
    _tu.register_dataclass(cls)
    return cls
