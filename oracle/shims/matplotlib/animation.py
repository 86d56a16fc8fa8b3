from matplotlib import _Inert

rcParams = {}


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Inert()
