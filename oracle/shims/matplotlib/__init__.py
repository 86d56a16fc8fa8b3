"""Inert `matplotlib` stand-in (rendering is out of scope)."""


class _Inert(object):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __iter__(self):
        return iter(())


def use(*a, **k):
    pass


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Inert()
