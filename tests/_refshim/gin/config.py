def external_configurable(fn, name=None, module=None):
    return fn
