"""Throw-away stand-in for the PyPI `autoray` package (absent here, no network).

GENERATOR SCAFFOLDING ONLY (see multimethod.py beside it)."""
import importlib

import numpy as _np

from . import autoray  # noqa: F401  (submodule with the tables rosnet.extra pokes)


def _resolve(like, args):
    if like is None or like == "numpy":
        return _np
    if isinstance(like, str):
        return importlib.import_module(like)
    return _np


def get_lib_fn(backend, name):
    mod = _resolve(backend, ())
    obj = mod
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj


def do(name, *args, like=None, **kwargs):
    return get_lib_fn(like, name)(*args, **kwargs)
