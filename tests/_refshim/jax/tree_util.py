"""Minimal pytree utilities (dict / list / tuple / namedtuple / registered dataclasses), jax leaf order: dict keys sorted."""
import dataclasses as _dc

_registered = {}      # cls -> (flatten, unflatten)


def register_dataclass(cls):
    names = [f.name for f in _dc.fields(cls)]
    _registered[cls] = (lambda obj: [getattr(obj, n) for n in names], lambda kids: cls(**dict(zip(names, kids))))
    return cls


def _children(tree):
    """-> (children, rebuild) or None for a leaf."""
    if tree is None:
        return [], lambda kids: None
    t = type(tree)
    if t in _registered:
        fl, un = _registered[t]
        return fl(tree), un
    if isinstance(tree, dict):
        keys = sorted(tree.keys())
        return [tree[k] for k in keys], lambda kids: t(dict(zip(keys, kids))) if t is not dict else dict(zip(keys, kids))
    if isinstance(tree, tuple) and hasattr(tree, '_fields'):
        return list(tree), lambda kids: t(*kids)
    if isinstance(tree, (list, tuple)):
        return list(tree), lambda kids: t(kids)
    return None


def tree_flatten(tree):
    leaves = []

    def rec(node):
        ch = _children(node)
        if ch is None:
            leaves.append(node)
            return ('leaf',)
        kids, rebuild = ch
        return ('node', rebuild, [rec(k) for k in kids])

    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        if d[0] == 'leaf':
            return next(it)
        return d[1]([rec(k) for k in d[2]])

    return rec(treedef)


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def tree_map(fn, tree, *rest):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [fn(*xs) for xs in zip(leaves, *others)])


tree_multimap = tree_map


def tree_reduce(fn, tree, initializer=None):
    leaves = tree_leaves(tree)
    if initializer is None:
        acc, leaves = leaves[0], leaves[1:]
    else:
        acc = initializer
    for l in leaves:
        acc = fn(acc, l)
    return acc
