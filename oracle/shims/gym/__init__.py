"""Minimal `gym` stand-in so that the UNMODIFIED reference imports here.
Test infrastructure only.  Used at simulator/__init__.py:1-7, simulator/env.py:2,19,
rl/utils/utils.py:4,28."""
import importlib

_reg = {}


class Env(object):
    pass


def make(env_id):
    module_name, class_name = _reg[env_id].split(":")
    return getattr(importlib.import_module(module_name), class_name)()
