"""NEP-18 glue: route ``np.<func>(BlockArray, ...)`` to ``rosnet_b200.dispatch.<func>``.

Same contract as the reference's ``ArrayFunctionMixin.__array_function__``
(``rosnet/core/mixin.py:17-41``): look the NumPy function up *by name* in the dispatch module
(``dispatch.linalg`` for ``np.linalg.*``) and return ``NotImplemented`` when there is no such
name, so NumPy raises ``TypeError``.  ``np.zeros/ones/full/random.rand`` carry no array argument
to dispatch on and are resolved on the array's own module (``mixin.py:8-13, 26-35``).
"""
import inspect

import numpy as np

EXPLICITLY_DISPATCH = [np.zeros, np.ones, np.full, np.random.rand]


class ArrayFunctionMixin:
    def __array_function__(self, func, types, args, kwargs):
        from rosnet_b200 import dispatch

        if func in EXPLICITLY_DISPATCH:
            module = inspect.getmodule(type(self))
            impl = getattr(module, func.__name__, None)
            return impl(*args, **kwargs) if impl is not None else NotImplemented

        module = dispatch.linalg if inspect.getmodule(func) is np.linalg else dispatch
        impl = getattr(module, func.__name__, None)
        if impl is None:
            return NotImplemented
        return impl(*args, **kwargs)


class ArrayAttributeMixin:
    """Method spellings of the dispatched functions (reference ``mixin.py:44-60``)."""

    def reshape(self, shape, order="C"):
        from rosnet_b200 import dispatch

        return dispatch.reshape(self, shape, order=order)

    def transpose(self, *axes):
        from rosnet_b200 import dispatch

        if len(axes) == 1 and not isinstance(axes[0], (int, np.integer)):
            axes = axes[0]
        return dispatch.transpose(self, axes=tuple(axes) if axes else None)

    @property
    def T(self):
        return self.transpose()
