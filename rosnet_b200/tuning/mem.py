"""Memory-footprint cost model in bytes (reference ``rosnet/tuning/mem.py:10-45``)."""
from math import prod

import numpy as np

from ..core.util import result_shape


def tensordot(a, b, axes) -> int:
    assert all(a.shape[i] == b.shape[j] for i, j in zip(*axes))
    shape = result_shape(a.shape, b.shape, axes)
    dtype = np.result_type(a.dtype, b.dtype)
    return a.nbytes + b.nbytes + prod(shape) * dtype.itemsize


def sequential(a, b, axes) -> int:
    """Footprint of one output block of the blocked contraction: all its a-blocks, all its
    b-blocks and the result (reference ``mem.py:18-24``, which reads ``.shape`` off the lists by
    mistake; the first blocks' shapes are what is meant)."""
    assert all(a[0].shape == ai.shape for ai in a)
    assert all(b[0].shape == bi.shape for bi in b)
    shape = result_shape(a[0].shape, b[0].shape, axes)
    dtype = np.result_type(a[0].dtype, b[0].dtype)
    return sum(ai.nbytes for ai in a) + sum(bi.nbytes for bi in b) + prod(shape) * dtype.itemsize


def commutative(buffer, a, b, axes) -> int:
    return tensordot(a, b, axes)


def full(shape, fill_value, dtype=None, **kwargs) -> int:
    dtype = np.dtype(dtype) if dtype is not None else np.dtype(type(fill_value))
    return prod(shape) * dtype.itemsize


def reshape(a, shape, **kwargs) -> int:
    return a.nbytes


def transpose(a, axes=None, **kwargs) -> int:
    """Out of place: source + destination (reference ``mem.py:41-45``)."""
    return 2 * a.nbytes
