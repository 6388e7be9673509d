from jax import tree_util as _tu
from jax._core import asarray


def replicate(tree):
    return _tu.tree_map(lambda v: asarray(v)[None], tree)


def unreplicate(tree):
    return _tu.tree_map(lambda v: v[0], tree)


def prefetch_to_device(it, n):
    return it
