"""Structural array protocols used as dispatch predicates.

Mirrors the reference's ABCs (``rosnet/core/interface.py:17-109``): a class *is* an
``ArrayConvertable`` if it has ``__array__``, an ``ArrayDispatchable`` if it implements both NumPy
override protocols, and an ``Array`` if it is both and exposes ``shape`` / ``dtype`` as
non-callable attributes.  ``BlockArray.__init__`` relies on these to accept foreign block types
(``rosnet/array/block/__init__.py:80-85``).
"""
import abc
from typing import Any, Tuple, Union

import numpy as np


def hasmethod(cls, name) -> bool:
    return callable(getattr(cls, name, None))


def hasproperty(cls, name) -> bool:
    return hasattr(cls, name) and not callable(getattr(cls, name))


class ArrayConvertable(metaclass=abc.ABCMeta):
    """Anything NumPy can turn into an ndarray (has ``__array__``), plus Python scalars."""

    @classmethod
    def __subclasshook__(cls, subclass: type):
        return hasmethod(subclass, "__array__")


for _t in (bool, int, float, complex, str, bytes):
    ArrayConvertable.register(_t)


class ArrayDispatchable(metaclass=abc.ABCMeta):
    """Implements NEP-13 and NEP-18 (``__array_ufunc__`` and ``__array_function__``)."""

    @classmethod
    def __subclasshook__(cls, subclass: type):
        return hasmethod(subclass, "__array_ufunc__") and hasmethod(subclass, "__array_function__")


class Array(metaclass=abc.ABCMeta):
    """``ArrayConvertable`` + ``ArrayDispatchable`` + ``shape``/``dtype`` attributes."""

    shape: Tuple[int, ...]
    dtype: Union[type, np.dtype]
    data: Any

    @classmethod
    def __subclasshook__(cls, subclass: type):
        return (
            issubclass(subclass, ArrayConvertable)
            and issubclass(subclass, ArrayDispatchable)
            and hasproperty(subclass, "shape")
            and hasproperty(subclass, "dtype")
        )


class Future(metaclass=abc.ABCMeta):
    """Handle to a value that is still being produced (reference ``interface.py:96-105``).
    Here: a device buffer with work pending on a CUDA stream."""


class AsyncArray(metaclass=abc.ABCMeta):
    """Array whose payload may still be in flight (reference ``interface.py:108-109``)."""

    data: Union[Array, Future]
