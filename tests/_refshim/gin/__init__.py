"""Test-only `gin`: a registry of `Name.field = value` bindings applied as constructor keyword defaults."""
import ast as _ast
import functools as _functools

from . import config  # noqa: F401

_bindings = {}        # configurable name -> {field: value}
_search = []


def add_config_file_search_path(p):
    _search.append(p)


def clear_config():
    _bindings.clear()


def bind_parameter(name, value):
    scope, field = name.rsplit('.', 1)
    _bindings.setdefault(scope.split('.')[-1], {})[field] = value


def parse_config(text):
    for line in text.splitlines():
        line = line.split('#')[0].strip()
        if '=' not in line:
            continue
        key, val = [s.strip() for s in line.split('=', 1)]
        try:
            v = _ast.literal_eval(val)
        except Exception:
            v = val
        bind_parameter(key, v)


def parse_config_file(path):
    with open(path) as f:
        parse_config(f.read())


def parse_config_files_and_bindings(files, bindings=None, **kw):
    for f in files or []:
        parse_config_file(f)
    for b in bindings or []:
        parse_config(b)


def _wrap(cls_or_fn, name):
    if isinstance(cls_or_fn, type):
        orig = cls_or_fn.__init__

        @_functools.wraps(orig)
        def __init__(self, *a, **k):
            merged = dict(_bindings.get(name, {}))
            merged.update(k)
            orig(self, *a, **merged)
        cls_or_fn.__init__ = __init__
        return cls_or_fn

    @_functools.wraps(cls_or_fn)
    def fn(*a, **k):
        merged = dict(_bindings.get(name, {}))
        merged.update(k)
        return cls_or_fn(*a, **merged)
    return fn


def configurable(arg=None, **kw):
    if callable(arg):
        return _wrap(arg, arg.__name__)
    return lambda obj: _wrap(obj, arg if isinstance(arg, str) else obj.__name__)
