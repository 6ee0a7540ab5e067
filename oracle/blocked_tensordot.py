"""CPU oracle for rosnet's blocked-tensor contraction path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy-only restatement of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker / the reported CPU baseline.  Nothing under
``rosnet_b200/`` imports it: the product path is the CUDA extension and fails loudly without it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the *unmodified* reference from
``/root/reference`` (under two throw-away stand-ins for the absent ``multimethod`` / ``autoray``
packages) and records its outputs in ``tests/golden/blocked_tensordot.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors bit-exactly,
including the reference's block mis-pairing defect in ``order="K"`` mode.
The reference's own tests pin layout only (``test/array/test_blockarray.py:115-125``), never a
value, so the numerical contract additionally is the dense ``numpy.tensordot`` of the assembled
operands (``order="C"`` mode must agree with it to rounding).

A "grid" here is what the reference calls ``BlockArray.data``: a NumPy object array whose
cells are equally-shaped ndarrays with ``ndim == grid.ndim``
(reference ``rosnet/array/block/__init__.py:30-94``).

Reference lines followed, function by function, are cited in each docstring.
"""
from __future__ import annotations

import itertools
from math import prod
from typing import Iterable, List, Sequence, Tuple

import numpy as np

__all__ = [
    "free_axes",
    "result_shape",
    "join_idx",
    "fold_pairs",
    "fold_pairs_commutative",
    "inner_visit_order",
    "blocked_tensordot",
    "blocked_transpose",
    "blocked_reshape",
    "split_blocks",
    "make_grid",
    "to_numpy",
    "grid_layout",
    "flops_tensordot",
    "mem_tensordot",
    "relative_frobenius",
]


# --------------------------------------------------------------------------------------
# index helpers  (reference rosnet/core/util.py)
# --------------------------------------------------------------------------------------
def free_axes(ndim: int, contracted: Sequence[int]) -> List[int]:
    """Axes of an operand that survive the contraction, ascending.

    Reference: ``outer_axes = list(set(range(ndim)) - set(ax))`` at
    ``rosnet/array/block/__init__.py:291`` and ``rosnet/core/util.py:23`` — small-int sets
    iterate ascending, so the order is ascending axis number.
    """
    drop = {int(c) % ndim if ndim else int(c) for c in contracted}
    return [k for k in range(ndim) if k not in drop]


def result_shape(a_shape: Sequence[int], b_shape: Sequence[int], axes) -> Tuple[int, ...]:
    """Shape of ``tensordot(a, b, axes)``: a's free extents then b's free extents.

    Reference: ``rosnet/core/util.py:21-24``.
    """
    fa = free_axes(len(a_shape), axes[0])
    fb = free_axes(len(b_shape), axes[1])
    return tuple(a_shape[k] for k in fa) + tuple(b_shape[k] for k in fb)


def join_idx(outer: Sequence[int], inner: Sequence[int], axes: Sequence[int]) -> Tuple[int, ...]:
    """Merge a free-axes multi-index and a contracted-axes multi-index into a full grid index.

    ``inner[j]`` lands on axis ``axes[j]``; ``outer`` fills the remaining axes ascending.
    Reference: ``rosnet/core/util.py:27-39``.
    """
    n = len(outer) + len(inner)
    full = [0] * n
    taken = set()
    for ax, v in zip(axes, inner):
        full[ax] = int(v)
        taken.add(ax)
    rest = [k for k in range(n) if k not in taken]
    for ax, v in zip(rest, outer):
        full[ax] = int(v)
    return tuple(full)


# --------------------------------------------------------------------------------------
# the k-block reduction  (reference block/__init__.py:281-283, compss/__init__.py:430-451)
# --------------------------------------------------------------------------------------
def fold_pairs(a_blocks: Iterable[np.ndarray], b_blocks: Iterable[np.ndarray], axes):
    """``sum(np.tensordot(ai, bi, axes) for ai, bi in zip(a, b))``.

    Python's ``sum`` is a left fold that starts from the int ``0``; the first addition
    ``0 + ndarray`` is exact, every later one rounds once.  This is also the body of the
    COMPSs ``sequential`` task (``rosnet/array/compss/task/tensordot.py:17-21``).
    Reference: ``rosnet/array/block/__init__.py:281-283``.
    """
    acc = 0
    for ai, bi in zip(a_blocks, b_blocks):
        acc = acc + np.tensordot(ai, bi, axes)
    return acc


def fold_pairs_commutative(a_blocks, b_blocks, axes, order: Sequence[int] | None = None):
    """The ``commutative`` / ``commutative-but-first`` reduction: ``res += tensordot(a, b)``
    per pair, in whatever order the runtime picks (``order`` = a permutation of the pair list;
    identity if None).  First pair initialises the accumulator (``MaybeArray`` semantics,
    ``rosnet/array/maybe.py:5-38``).

    Reference: ``rosnet/array/compss/__init__.py:441-448``,
    ``rosnet/array/compss/task/tensordot.py:24-28,45-49``.
    """
    pairs = list(zip(a_blocks, b_blocks))
    if order is not None:
        pairs = [pairs[i] for i in order]
    acc = None
    for ai, bi in pairs:
        term = np.tensordot(ai, bi, axes)
        if acc is None:
            acc = term
        else:
            acc += term
    return acc


def inner_visit_order(grid: np.ndarray, contracted: Sequence[int], order: str) -> List[Tuple[int, ...]]:
    """Multi-indices over the contracted grid axes in the order the reference visits them.

    The returned tuples are aligned with ``contracted`` (entry j belongs to axis
    ``contracted[j]``) so they can be fed to :func:`join_idx`.

    ``order="C"``: row-major over the axes *as listed* — the pairing that makes block j of a
    meet block j of b, i.e. the mathematically correct one.

    ``order="K"``: what ``np.nested_iters`` does by default at
    ``rosnet/array/block/__init__.py:292-303`` — the inner axes are walked in *memory* order of
    ``grid`` (largest |stride| outermost), whatever order they were listed in.  For a
    C-contiguous grid that is ascending axis number.  When a and b disagree on the relative
    memory order of two contracted axes with grid extent > 1 this pairs the wrong blocks
    (SURVEY Appendix A1); kept so the restatement can be checked bit-exactly against the
    reference's recorded output.
    """
    contracted = [int(c) for c in contracted]
    if order == "C":
        walk = list(range(len(contracted)))
    elif order == "K":
        # stable sort by decreasing |stride|; ties keep ascending axis number
        walk = sorted(range(len(contracted)), key=lambda j: (-abs(grid.strides[contracted[j]]), contracted[j]))
    else:
        raise ValueError("order must be 'C' or 'K'")
    extents = [grid.shape[contracted[j]] for j in walk]
    out = []
    for idx in itertools.product(*[range(e) for e in extents]):
        aligned = [0] * len(contracted)
        for pos, j in enumerate(walk):
            aligned[j] = idx[pos]
        out.append(tuple(aligned))
    return out


# --------------------------------------------------------------------------------------
# the path itself  (reference block/__init__.py:286-336)
# --------------------------------------------------------------------------------------
def blocked_tensordot(a: np.ndarray, b: np.ndarray, axes, order: str = "C", method: str = "sequential") -> np.ndarray:
    """Blocked contraction of two grids; returns the result grid (object ndarray).

    For every output block — a's free grid axes (ascending) followed by b's free grid axes
    (ascending), row-major — gather the a-blocks and b-blocks along the contracted grid axes
    and reduce their pairwise ``numpy.tensordot`` with :func:`fold_pairs`.

    ``axes`` is the explicit pair form ``(axes_a, axes_b)`` the reference requires
    (it indexes ``axes[0]`` / ``axes[1]``, ``:291``).

    Reference: ``rosnet/array/block/__init__.py:286-336`` (enumeration and output ordering),
    ``:281-283`` (fold), ``rosnet/core/util.py:27-39`` (index merge).
    """
    ax_a = [int(x) for x in axes[0]]
    ax_b = [int(x) for x in axes[1]]
    fa = free_axes(a.ndim, ax_a)
    fb = free_axes(b.ndim, ax_b)
    out_grid = tuple(a.shape[k] for k in fa) + tuple(b.shape[k] for k in fb)
    out = np.empty(out_grid, dtype=object)

    inner_a = inner_visit_order(a, ax_a, order)
    inner_b = inner_visit_order(b, ax_b, order)

    for oa in itertools.product(*[range(a.shape[k]) for k in fa]):
        for ob in itertools.product(*[range(b.shape[k]) for k in fb]):
            a_list = [a[join_idx(oa, ia, ax_a)] for ia in inner_a]
            b_list = [b[join_idx(ob, ib, ax_b)] for ib in inner_b]
            if method == "sequential":
                out[oa + ob] = fold_pairs(a_list, b_list, (ax_a, ax_b))
            elif method in ("commutative", "commutative-but-first"):
                out[oa + ob] = fold_pairs_commutative(a_list, b_list, (ax_a, ax_b))
            else:
                raise ValueError("invalid method")
    return out


def blocked_transpose(a: np.ndarray, axes: Sequence[int]) -> np.ndarray:
    """Intent of ``transpose(a: BlockArray, axes)``: permute every block by ``axes`` and permute
    the grid by the same ``axes``.  The shipped function raises ``AttributeError`` after doing
    exactly these two steps, so this restates the two steps.

    Reference: ``rosnet/array/block/__init__.py:264-278``.
    """
    axes = [int(x) for x in axes]
    if len(set(axes)) != len(axes):
        raise ValueError("'axes' must be a unique list: %s" % (axes,))
    out = np.empty(a.shape, dtype=object)
    for idx in np.ndindex(*a.shape):
        out[idx] = np.transpose(a[idx], axes)
    return np.transpose(out, axes)


def blocked_reshape(a: np.ndarray, shape, order: str = "F") -> np.ndarray:
    """``reshape(a: BlockArray, shape, order="F")``: every block is reshaped to ``shape`` (a new
    *blockshape*); the grid is untouched.  Default order is Fortran, as in the reference.

    Reference: ``rosnet/array/block/__init__.py:251-261``.
    """
    out = np.empty(a.shape, dtype=object)
    for idx in np.ndindex(*a.shape):
        out[idx] = np.reshape(a[idx], shape, order=order)
    return out


# --------------------------------------------------------------------------------------
# construction / conversion
# --------------------------------------------------------------------------------------
def split_blocks(arr: np.ndarray, blockshape: Sequence[int]) -> np.ndarray:
    """Cut a dense array into a grid of C-contiguous blocks (``grid = shape // blockshape``).

    Reference: the commented-out ``array`` constructor,
    ``rosnet/array/block/__init__.py:339-366`` (block ``bidx`` covers
    ``[bidx*bs, (bidx+1)*bs)`` on every axis), and ``full`` at ``:216-228`` for the grid rule.
    """
    blockshape = tuple(int(x) for x in blockshape)
    grid = tuple(s // bs for s, bs in zip(arr.shape, blockshape))
    out = np.empty(grid, dtype=object)
    for bidx in np.ndindex(*grid):
        sl = tuple(slice(i * bs, (i + 1) * bs) for i, bs in zip(bidx, blockshape))
        out[bidx] = np.ascontiguousarray(arr[sl])
    return out


def make_grid(blocks: Sequence[np.ndarray], grid: Sequence[int]) -> np.ndarray:
    """Flat list of blocks + grid shape -> object grid, row-major
    (``BlockArray(list, grid=...)``, reference ``block/__init__.py:52-85``)."""
    out = np.empty(tuple(grid), dtype=object)
    for i, idx in enumerate(np.ndindex(*out.shape)):
        out[idx] = blocks[i]
    return out


def to_numpy(grid: np.ndarray) -> np.ndarray:
    """Assemble the dense array: ``np.block(grid.tolist())``.

    Reference: ``rosnet/array/block/__init__.py:203-205``.  A 0-d grid (full contraction,
    SURVEY Appendix A6) holds a single 0-d block and yields that scalar array.
    """
    if grid.ndim == 0:
        return np.asarray(grid[()])
    return np.block(grid.tolist())


def grid_layout(grid: np.ndarray):
    """(shape, grid, blockshape, dtype) as the reference's properties report them
    (``rosnet/array/block/__init__.py:143-185``)."""
    first = grid.flat[0]
    bs = tuple(first.shape)
    return tuple(int(g * s) for g, s in zip(grid.shape, bs)), tuple(grid.shape), bs, np.dtype(first.dtype)


# --------------------------------------------------------------------------------------
# cost models  (reference rosnet/tuning/flops.py, mem.py)
# --------------------------------------------------------------------------------------
def flops_tensordot(a_shape, b_shape, axes) -> int:
    """Multiply-add count: prod(free extents of a and b) * prod(contracted extents).

    Reference: ``rosnet/tuning/flops.py:8-13`` (pinned by ``test/tuning/test_flops.py:8-16``).
    """
    assert all(a_shape[i] == b_shape[j] for i, j in zip(*axes))
    return prod(result_shape(a_shape, b_shape, axes)) * prod(a_shape[i] for i in axes[0])


def mem_tensordot(a_shape, a_dtype, b_shape, b_dtype, axes) -> int:
    """bytes(a) + bytes(b) + bytes(out).  Reference: ``rosnet/tuning/mem.py:10-15``."""
    out_dtype = np.result_type(a_dtype, b_dtype)
    return (
        prod(a_shape) * np.dtype(a_dtype).itemsize
        + prod(b_shape) * np.dtype(b_dtype).itemsize
        + prod(result_shape(a_shape, b_shape, axes)) * out_dtype.itemsize
    )


def relative_frobenius(x: np.ndarray, ref: np.ndarray) -> float:
    """||x - ref||_F / ||ref||_F (the parity metric of BASELINE.json's north_star)."""
    x = np.asarray(x)
    ref = np.asarray(ref)
    den = float(np.linalg.norm(ref.ravel()))
    num = float(np.linalg.norm((x.astype(np.result_type(x, ref)) - ref).ravel()))
    return num / den if den else num
