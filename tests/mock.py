"""Shape/dtype-only block type fulfilling the Array protocol (the role of the reference's
``test/mock.py:6-45``)."""
from math import prod

import numpy as np


class MockArray:
    def __init__(self, shape, **kwargs):
        self._shape = tuple(shape)
        self._dtype = kwargs.get("dtype", np.generic)

    def __array__(self, dtype=None, copy=None):
        return np.zeros(self._shape, dtype=self._dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        return NotImplemented

    def __array_function__(self, function, types, args, kwargs):
        return NotImplemented

    def __getitem__(self, key):
        return 0

    shape = property(lambda self: self._shape)
    dtype = property(lambda self: self._dtype)
    ndim = property(lambda self: len(self._shape))
    size = property(lambda self: prod(self._shape))
    itemsize = property(lambda self: np.dtype(self._dtype).itemsize)
    nbytes = property(lambda self: prod(self._shape) * np.dtype(self._dtype).itemsize)
