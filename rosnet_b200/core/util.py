"""Index helpers of the blocked path (reference ``rosnet/core/util.py``)."""
import itertools
from typing import Sequence

import numpy as np

from .interface import Array


def isunique(seq: Sequence) -> bool:
    """True when no element of ``seq`` repeats (reference ``util.py:11-13``)."""
    seq = list(seq)
    return len(set(seq)) == len(seq)


def space(extents: Sequence[int]):
    """Row-major iterator over all multi-indices below ``extents`` (reference ``util.py:16-18``)."""
    return itertools.product(*(range(int(e)) for e in extents))


def normalize_axes(axes, ndim_a: int, ndim_b: int):
    """``axes`` as two equal-length lists of non-negative ints.

    The reference requires the explicit pair form (it indexes ``axes[0]``/``axes[1]``,
    ``rosnet/array/block/__init__.py:291``); NumPy's integer form ``axes=n`` is accepted here
    too (last n of a with first n of b), as ``numpy.tensordot`` itself does.
    """
    if isinstance(axes, (int, np.integer)):
        n = int(axes)
        ax_a, ax_b = list(range(ndim_a - n, ndim_a)), list(range(n))
    else:
        ax_a, ax_b = axes
        ax_a = [ax_a] if isinstance(ax_a, (int, np.integer)) else list(ax_a)
        ax_b = [ax_b] if isinstance(ax_b, (int, np.integer)) else list(ax_b)
    ax_a = [int(x) + ndim_a if int(x) < 0 else int(x) for x in ax_a]
    ax_b = [int(x) + ndim_b if int(x) < 0 else int(x) for x in ax_b]
    if len(ax_a) != len(ax_b):
        raise ValueError("shape-mismatch for sum")
    if not isunique(ax_a) or not isunique(ax_b):
        raise ValueError("repeated axis in tensordot axes")
    if any(x >= ndim_a for x in ax_a) or any(x >= ndim_b for x in ax_b):
        raise ValueError("tensordot axis out of range")
    return ax_a, ax_b


def free_axes(ndim: int, contracted: Sequence[int]):
    """Surviving axes, ascending (reference ``block/__init__.py:291``: ``set`` difference of
    small ints iterates ascending)."""
    drop = set(contracted)
    return [k for k in range(ndim) if k not in drop]


def result_shape(a, b, axes):
    """(Block)shape after ``tensordot``: a's free extents then b's (reference ``util.py:21-24``)."""
    fa = free_axes(len(a), axes[0])
    fb = free_axes(len(b), axes[1])
    return tuple(a[k] for k in fa) + tuple(b[k] for k in fb)


def join_idx(outer, inner, axes):
    """Full grid index from a free-axes index and a contracted-axes index
    (reference ``util.py:27-39``): ``inner[j]`` goes to axis ``axes[j]``, ``outer`` fills the
    rest in ascending axis order."""
    n = len(outer) + len(inner)
    full = [0] * n
    for ax, v in zip(axes, inner):
        full[ax] = int(v)
    rest = free_axes(n, axes)
    for ax, v in zip(rest, outer):
        full[ax] = int(v)
    return tuple(full)


def _levels(x):
    """Yield each nesting level of a nested list / object ndarray of blocks."""
    while True:
        if isinstance(x, list):
            yield x
            x = x[0]
        elif isinstance(x, np.ndarray) and x.dtype == object and x.size and isinstance(x.flat[0], (np.ndarray, Array)):
            yield x
            return
        else:
            return


def nest_level(x) -> int:
    """How many container levels wrap the blocks (reference ``util.py:67-68``).  A plain numeric
    ndarray or a single block has level 0; an object ndarray of blocks has level 1; nested
    lists count one per level."""
    return sum(1 for _ in _levels(x))


def measure_shape(x):
    """Grid shape implied by a nested list (reference ``util.py:71-72``)."""
    out = []
    for level in _levels(x):
        if isinstance(level, np.ndarray):
            out.extend(level.shape)
        else:
            out.append(len(level))
    return tuple(out)
