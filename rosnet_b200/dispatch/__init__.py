"""Dispatch surface: by-name entry points for NumPy's ``__array_function__`` protocol and for
by-name backend lookup (opt_einsum / autoray ``backend="rosnet"``).

Mirrors ``rosnet/dispatch/__init__.py``: ``to_numpy`` (``:5-13``), the NumPy-named dispatchers
(``:16-30``), ``linalg`` (``:32``) and every ``np.ufunc`` re-exported as a module attribute
(``:35-38``)."""
import numpy as _np

from ._registry import Dispatcher
from . import linalg  # noqa: F401
from .numpy import (  # noqa: F401
    block,
    count_nonzero,
    cumsum,
    einsum,
    empty_like,
    full_like,
    ones_like,
    reshape,
    split,
    stack,
    tensordot,
    transpose,
    zeros_like,
)

to_numpy = Dispatcher("to_numpy")


@to_numpy.register(_np.ndarray)
def _to_numpy_ndarray(a):
    return a


@to_numpy.register(_np.generic)
def _to_numpy_scalar(a):
    return a


for _name, _obj in list(vars(_np).items()):
    if isinstance(_obj, _np.ufunc):
        globals()[_name] = _obj
del _name, _obj
