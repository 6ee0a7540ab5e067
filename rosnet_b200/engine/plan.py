"""Contraction planning: turn ``tensordot(BlockArray, BlockArray, axes)`` into one grouped launch.

Pure host logic (NumPy index arithmetic only, no device access) so it is testable on a CPU box.
It replaces the Python double loop over output blocks and the per-block index merging of the
reference (``rosnet/array/block/__init__.py:286-336``, ``rosnet/core/util.py:27-39``):

* output blocks are enumerated a's free grid axes (ascending) then b's free grid axes
  (ascending), row-major — the reference's output ordering (``:291, :305, :309``);
* for every output block the contracted grid multi-index runs row-major over the axes in the
  order they are *listed*, which pairs block j of a with block j of b.  The reference's
  ``np.nested_iters`` walks them in memory order instead and mis-pairs blocks when the two
  operands list their contracted axes in different relative order;
  the plan implements the mathematically correct pairing (= dense ``numpy.tensordot``);
* the order of the contracted axis *pairs* is free (it is a summation index), so it is chosen
  to let as many operands as possible be consumed in place by the GEMM (no permute pass);
* an operand block is usable in place when its memory order is [free..., contracted...]
  ("K-major") or [contracted..., free...] ("MN-major") and the TMA alignment rules of
  ``csrc/gemm_f64.cu`` hold; otherwise all its blocks are matricized by ONE permute launch into a
  zero-padded K-major workspace.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field
from math import prod
from typing import Optional, Tuple

import numpy as np

from ..core.util import free_axes, normalize_axes
from . import lib as _lib

SUPPORTED = ("float32", "float64", "complex64", "complex128")


@dataclass(frozen=True)
class OperandPlan:
    direct: bool                 # consumed in place (no permute pass)
    major: int                   # RNB_MAJOR_K / RNB_MAJOR_MN as the GEMM sees it
    rows: int                    # M (for a) or N (for b)
    k: int                       # contracted extent of one block
    ld: int                      # leading dimension, elements
    block_stride: int            # elements between consecutive blocks (storage or workspace)
    nblocks: int
    perm: Tuple[int, ...] = ()   # block-axis order written to the workspace (matricize only)
    ws_elems: int = 0            # workspace size in elements (matricize only)
    zero_fill: bool = False      # workspace has padding columns that must be zero


@dataclass(frozen=True)
class ContractionPlan:
    dtype: str
    ax_a: Tuple[int, ...]        # contracted axes of a, in the chosen pair order
    ax_b: Tuple[int, ...]
    free_a: Tuple[int, ...]
    free_b: Tuple[int, ...]
    out_grid: Tuple[int, ...]
    out_blockshape: Tuple[int, ...]
    m: int
    n: int
    k: int
    a: OperandPlan
    b: OperandPlan
    n_out: int
    n_pairs: int
    pair_a: np.ndarray = field(compare=False, repr=False)   # int32 [n_out, n_pairs]
    pair_b: np.ndarray = field(compare=False, repr=False)
    # small contractions (rnb_contract_nd): merged (extent, stride) lists of a's free, b's free and the contracted axes,
    # as (ext_m, sa_m, ext_n, sb_n, ext_k, sa_k, sb_k); None when the contraction is not small
    nd: Optional[tuple] = field(default=None, compare=False, repr=False)
    # nd lists of a contraction with one TINY operand and one large one (csrc/skinny.cu: apply_nd_kernel): no offset tables
    nd_apply: bool = field(default=False, compare=False, repr=False)

    @property
    def flops(self) -> int:
        """Algorithmic flops of the whole contraction (2 per real multiply-add, 8 per complex)."""
        per = 8 if self.dtype.startswith("complex") else 2
        return per * self.m * self.n * self.k * self.n_out * self.n_pairs

    @property
    def out_block_elems(self) -> int:
        return self.m * self.n


def _is_identity_layout(order, extents) -> bool:
    """True when visiting axes in ``order`` walks memory linearly, ignoring extent-1 axes."""
    kept = [ax for ax in order if extents[ax] != 1]
    return all(kept[i] < kept[i + 1] for i in range(len(kept) - 1))


def _direct_ok(major: int, rows: int, k: int, ld: int, block_stride: int, nblocks: int, itemsize: int) -> bool:
    """Alignment rules of the in-place TMA path (csrc/gemm_f64.cu: make_operand_map)."""
    if (ld * itemsize) % 16:
        return False
    if nblocks > 1 and (block_stride * itemsize) % 16:
        return False
    if major == _lib.RNB_MAJOR_K:
        return k % 4 == 0 and k >= 4
    group = 2 if itemsize == 16 else 4   # MN-major slabs group 32 bytes of rows
    return rows % group == 0 and rows >= group


def _plan_operand(blockshape, free, contracted, nblocks, itemsize, force_matricize=False, single=False) -> OperandPlan:
    rows = prod(blockshape[ax] for ax in free)
    k = prod(blockshape[ax] for ax in contracted)
    blk = prod(blockshape)
    if single:
        # float32 / complex64 (csrc/gemm_tf32.cu): the split pre-pass reads K-major operands with any
        # pitch and writes its own aligned hi/lo planes, so only the storage order matters
        if _is_identity_layout(list(free) + list(contracted), blockshape):
            return OperandPlan(True, _lib.RNB_MAJOR_K, rows, k, k, blk, nblocks)
        return OperandPlan(False, _lib.RNB_MAJOR_K, rows, k, k, rows * k, nblocks, tuple(free) + tuple(contracted), rows * k * nblocks, False)
    if force_matricize:
        pass
    elif _is_identity_layout(list(free) + list(contracted), blockshape) and _direct_ok(_lib.RNB_MAJOR_K, rows, k, k, blk, nblocks, itemsize):
        return OperandPlan(True, _lib.RNB_MAJOR_K, rows, k, k, blk, nblocks)
    elif _is_identity_layout(list(contracted) + list(free), blockshape) and _direct_ok(_lib.RNB_MAJOR_MN, rows, k, rows, blk, nblocks, itemsize):
        return OperandPlan(True, _lib.RNB_MAJOR_MN, rows, k, rows, blk, nblocks)
    # matricize every block to [rows][kp], K-major, kp a multiple of 4 (zero padded)
    kp = -(-k // 4) * 4
    stride = rows * kp
    if (stride * itemsize) % 16:
        stride += 1  # only for 8-byte elements with odd rows*kp: impossible since kp % 4 == 0; kept for clarity
    return OperandPlan(False, _lib.RNB_MAJOR_K, rows, kp, kp, stride, nblocks, tuple(free) + tuple(contracted), stride * nblocks, kp != k)


def _pair_tables(a_grid, b_grid, fa, fb, ax_a, ax_b):
    """Block numbers (row-major over each operand's grid) feeding every output block."""
    def strides(grid):
        s, acc = [0] * len(grid), 1
        for i in range(len(grid) - 1, -1, -1):
            s[i] = acc
            acc *= grid[i]
        return s

    sa, sb = strides(a_grid), strides(b_grid)

    def offsets(grid, axes, st):
        """row-major enumeration over `axes` (in the given order) -> linear block offsets"""
        off = np.zeros((), dtype=np.int64)
        for ax in axes:
            off = np.add.outer(off, np.arange(grid[ax], dtype=np.int64) * st[ax])
        return np.asarray(off).reshape(-1)

    outer_a = offsets(a_grid, fa, sa)
    outer_b = offsets(b_grid, fb, sb)
    inner_a = offsets(a_grid, ax_a, sa)
    inner_b = offsets(b_grid, ax_b, sb)
    n_oa, n_ob, n_k = outer_a.size, outer_b.size, inner_a.size
    pair_a = np.broadcast_to(outer_a[:, None, None] + inner_a[None, None, :], (n_oa, n_ob, n_k))
    pair_b = np.broadcast_to(outer_b[None, :, None] + inner_b[None, None, :], (n_oa, n_ob, n_k))
    return (
        np.ascontiguousarray(pair_a.reshape(n_oa * n_ob, n_k), dtype=np.int32),
        np.ascontiguousarray(pair_b.reshape(n_oa * n_ob, n_k), dtype=np.int32),
    )


def plan_tensordot(a_grid, a_blockshape, b_grid, b_blockshape, dtype, axes) -> ContractionPlan:
    """Plan ``tensordot`` of two block grids.  Raises ``ValueError`` for mismatched extents
    (NumPy's "shape-mismatch for sum") and ``NotImplementedError`` for unsupported dtypes."""
    # first-level cache on the arguments as given (tuples of ints from BlockArray properties): a
    # tensor-network path asks for the same few hundred plans once per slice
    try:
        key = (a_grid, a_blockshape, b_grid, b_blockshape, dtype, (tuple(axes[0]), tuple(axes[1])))
        hit = _fast_cache.get(key)
        if hit is not None:
            return hit
    except TypeError:
        key = None
    plan = _plan_cached(tuple(int(x) for x in a_grid), tuple(int(x) for x in a_blockshape),
                        tuple(int(x) for x in b_grid), tuple(int(x) for x in b_blockshape),
                        str(np.dtype(dtype)), _freeze_axes(axes, len(a_grid), len(b_grid)))
    if key is not None:
        if len(_fast_cache) > 8192:
            _fast_cache.clear()
        _fast_cache[key] = plan
    return plan


_fast_cache = {}


def _freeze_axes(axes, nda, ndb):
    ax_a, ax_b = normalize_axes(axes, nda, ndb)
    return (tuple(ax_a), tuple(ax_b))


@functools.lru_cache(maxsize=4096)
def _plan_cached(a_grid, a_bs, b_grid, b_bs, dtype, axes) -> ContractionPlan:
    if dtype not in SUPPORTED:
        raise NotImplementedError(f"blocked tensordot supports {SUPPORTED}, not {dtype}")
    if len(a_grid) != len(a_bs) or len(b_grid) != len(b_bs):
        raise ValueError("blocks and grid should have same ndim")
    ax_a, ax_b = axes
    for i, j in zip(ax_a, ax_b):
        if a_bs[i] != b_bs[j] or a_grid[i] != b_grid[j]:
            raise ValueError("shape-mismatch for sum")
    itemsize = np.dtype(dtype).itemsize
    fa = tuple(free_axes(len(a_grid), ax_a))
    fb = tuple(free_axes(len(b_grid), ax_b))
    na, nb = prod(a_grid), prod(b_grid)

    # candidate orders of the contracted pairs: as listed, ascending in a, ascending in b
    npairs = len(ax_a)
    orders = [tuple(range(npairs)), tuple(sorted(range(npairs), key=lambda j: ax_a[j])), tuple(sorted(range(npairs), key=lambda j: ax_b[j]))]
    best = None
    for order in dict.fromkeys(orders):
        ca = tuple(ax_a[j] for j in order)
        cb = tuple(ax_b[j] for j in order)
        single = dtype in ("float32", "complex64")
        pa = _plan_operand(a_bs, fa, ca, na, itemsize, single=single)
        pb = _plan_operand(b_bs, fb, cb, nb, itemsize, single=single)
        if pa.k != pb.k:  # one side was zero-padded in k: both must present the same padded k
            if pa.direct:
                pa = _plan_operand(a_bs, fa, ca, na, itemsize, force_matricize=True)
            if pb.direct:
                pb = _plan_operand(b_bs, fb, cb, nb, itemsize, force_matricize=True)
        cost = (0 if pa.direct else prod(a_bs) * na) + (0 if pb.direct else prod(b_bs) * nb)
        if best is None or cost < best[0]:
            best = (cost, ca, cb, pa, pb)
    _, ca, cb, pa, pb = best

    pair_a, pair_b = _pair_tables(a_grid, b_grid, fa, fb, ca, cb)
    nd = _nd_lists(a_bs, b_bs, fa, fb, ca, cb, pair_a.shape[0], pair_a.shape[1])
    out_grid = tuple(a_grid[ax] for ax in fa) + tuple(b_grid[ax] for ax in fb)
    out_bs = tuple(a_bs[ax] for ax in fa) + tuple(b_bs[ax] for ax in fb)
    return ContractionPlan(
        dtype=dtype, ax_a=ca, ax_b=cb, free_a=fa, free_b=fb, out_grid=out_grid, out_blockshape=out_bs,
        m=pa.rows, n=pb.rows, k=pa.k, a=pa, b=pb, n_out=pair_a.shape[0], n_pairs=pair_a.shape[1],
        pair_a=pair_a, pair_b=pair_b,
        nd=nd, nd_apply=nd is not None and not _nd_small(a_bs, b_bs, fa, ca, fb, pair_a.shape[0], pair_a.shape[1]),
    )


@functools.lru_cache(maxsize=1024)
def plan_batched(batch_grid, a_grid, a_bs, b_grid, b_bs, dtype, axes) -> ContractionPlan:
    """Batched blocked contraction ``C[batch, free_a, free_b] = sum_k A[batch, free_a, k] B[batch, free_b, k]`` for
    operands whose LEADING axes are the batch axes with block extent 1 (grid extent ``batch_grid``): the plan of the
    batch-free contraction (``a_grid`` etc. are the operands' grids WITHOUT the batch axes) repeated per batch block —
    output block ``(beta, g)`` pairs a-block ``beta * na + pair_a[g][p]`` with b-block ``beta * nb + pair_b[g][p]`` — so one
    grouped launch runs the whole batch (numpy.einsum with a shared, kept index; the reference issues one einsum task
    per block, ``rosnet/array/compss/__init__.py:554-568``)."""
    import dataclasses

    p = _plan_cached(a_grid, a_bs, b_grid, b_bs, dtype, axes)
    nbatch = prod(batch_grid)
    na, nb = prod(a_grid), prod(b_grid)
    beta = np.arange(nbatch, dtype=np.int64)
    pair_a = (p.pair_a[None, :, :].astype(np.int64) + (beta * na)[:, None, None]).reshape(nbatch * p.n_out, p.n_pairs).astype(np.int32)
    pair_b = (p.pair_b[None, :, :].astype(np.int64) + (beta * nb)[:, None, None]).reshape(nbatch * p.n_out, p.n_pairs).astype(np.int32)
    pa = dataclasses.replace(p.a, nblocks=p.a.nblocks * nbatch, ws_elems=p.a.ws_elems * nbatch)
    pb = dataclasses.replace(p.b, nblocks=p.b.nblocks * nbatch, ws_elems=p.b.ws_elems * nbatch)
    return dataclasses.replace(p, a=pa, b=pb, n_out=p.n_out * nbatch, pair_a=np.ascontiguousarray(pair_a), pair_b=np.ascontiguousarray(pair_b),
                               out_grid=tuple(batch_grid) + p.out_grid, out_blockshape=(1,) * len(batch_grid) + p.out_blockshape, nd=None, nd_apply=False)


ND_MAX_K, ND_MAX_OUT, ND_MAX_MADDS = 512, 1 << 20, 1 << 22   # the "small" route of csrc/skinny.cu (simt_route)


def _merged_axes(extents, *stride_lists):
    """Drop extent-1 axes and merge neighbours that are contiguous in EVERY stride list (row-major, last axis fastest)."""
    ext, strides = [], [[] for _ in stride_lists]
    for d, e in enumerate(extents):
        if e == 1:
            continue
        if ext and all(sl[-1] == e * st[d] for sl, st in zip(strides, stride_lists)):
            ext[-1] *= e
            for sl, st in zip(strides, stride_lists):
                sl[-1] = st[d]
        else:
            ext.append(e)
            for sl, st in zip(strides, stride_lists):
                sl.append(st[d])
    return (ext,) + tuple(strides)


APPLY_MAX_SMALL, APPLY_MAX_K, APPLY_MAX_TILE = 16, 64, 1024   # csrc/skinny.cu: apply_eligible


def _nd_small(a_bs, b_bs, fa, ca, fb, n_out, n_pairs) -> bool:
    """The "small" route of rnb_contract_nd: one thread per output element, offsets tabulated by the caller."""
    m, n, k = prod(a_bs[x] for x in fa), prod(b_bs[x] for x in fb), prod(a_bs[x] for x in ca)
    outs, k_total = n_out * m * n, n_pairs * k
    return k_total <= ND_MAX_K and outs <= ND_MAX_OUT and outs * k_total <= ND_MAX_MADDS


def _nd_apply(a_bs, b_bs, fa, ca, fb, n_out, n_pairs) -> bool:
    """The "apply" route of rnb_contract_nd: one operand of <= 16 rows and k <= 64 (a gate), the other one large and
    left in its n-d layout; HBM-bound, one read of the large operand and one write of the result (no matricize pass)."""
    m, n, k = prod(a_bs[x] for x in fa), prod(b_bs[x] for x in fb), prod(a_bs[x] for x in ca)
    ns = min(m, n)
    pad = 2 if ns <= 2 else 4 if ns <= 4 else 8 if ns <= 8 else 16
    return (ns <= APPLY_MAX_SMALL and k <= APPLY_MAX_K and n_pairs * pad * k <= APPLY_MAX_TILE and n_out <= 65535
            and max(prod(a_bs), prod(b_bs)) <= (1 << 30))


def _nd_lists(a_bs, b_bs, fa, fb, ca, cb, n_out, n_pairs):
    """The rnb_contract_nd axis lists of a small contraction or of a gate application, or None if it is neither / has too many axes."""
    if not (_nd_small(a_bs, b_bs, fa, ca, fb, n_out, n_pairs) or _nd_apply(a_bs, b_bs, fa, ca, fb, n_out, n_pairs)):
        return None
    sa = [prod(a_bs[i + 1:]) for i in range(len(a_bs))]
    sb = [prod(b_bs[i + 1:]) for i in range(len(b_bs))]
    ext_m, sa_m = _merged_axes([a_bs[x] for x in fa], [sa[x] for x in fa])
    ext_n, sb_n = _merged_axes([b_bs[x] for x in fb], [sb[x] for x in fb])
    ext_k, sa_k, sb_k = _merged_axes([a_bs[x] for x in ca], [sa[x] for x in ca], [sb[x] for x in cb])
    if max(len(ext_m), len(ext_n), len(ext_k)) > _lib.RNB_MAX_RANK:
        return None
    return (tuple(ext_m), tuple(sa_m), tuple(ext_n), tuple(sb_n), tuple(ext_k), tuple(sa_k), tuple(sb_k))


def matricize_desc(blockshape, n_free, perm, ld, block_stride, nblocks):
    """Extents / source strides / destination strides (elements) of the ONE permute launch that
    matricizes all blocks of an operand: leading axis = block number, then the block axes.
    Destination: block-major, each block [rows][ld]; the free axes (first ``n_free`` entries of
    ``perm``) index the rows, the contracted axes form the contiguous part of a row."""
    r = len(blockshape)
    blk = prod(blockshape)
    extents = [nblocks] + [blockshape[i] for i in range(r)]
    src = [blk] + [prod(blockshape[i + 1:]) for i in range(r)]
    dst_of_axis = {}
    acc = 1
    for ax in reversed(perm[n_free:]):      # contracted axes: contiguous inside a row
        dst_of_axis[ax] = acc
        acc *= blockshape[ax]
    acc = ld                                  # free axes: rows of pitch ld
    for ax in reversed(perm[:n_free]):
        dst_of_axis[ax] = acc
        acc *= blockshape[ax]
    dst = [block_stride] + [dst_of_axis[i] for i in range(r)]
    return extents, src, dst
