"""The two gym entry points the reference uses (simulator/__init__.py:1-7, rl/utils/utils.py:28):
`register(id, entry_point)` and `make(id)`.  `gym` itself is not a dependency of the hot path."""
import importlib

_registry = {}


class Env(object):
    pass


def register(id, entry_point, **kwargs):
    _registry[id] = entry_point


def make(env_id):
    module_name, class_name = _registry[env_id].split(":")
    return getattr(importlib.import_module(module_name), class_name)()
