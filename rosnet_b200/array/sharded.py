"""``ShardedBlockArray`` and the distributed ``tensordot``: the single-box scheduler behind the array API.

In the reference the *library* fans a blocked contraction out — ``tensordot`` on COMPSs-backed blocks submits one task
per block pair and a task tree for the sum (``rosnet/array/compss/__init__.py:430-451``, ``rosnet/tuning/task.py:49-63``) —
and the user only writes ``np.tensordot``.  Here one process drives one GPU (torchrun) and, once
``rosnet_b200.engine.scheduler.init()`` has been called, the same ``np.tensordot`` deals the work over the ranks:

* ``tensordot(BlockArray, BlockArray)`` with replicated operands: ``scheduler.choose_tier`` picks the output-block
  partition (tier 1, no exchange), the contracted-index partition (tier 2) or a local contraction from
  ``tuning.configure(n_gpus=, threshold_k=)``;
* ``tensordot(ShardedBlockArray, ShardedBlockArray)`` with both operands sharded along a pair of contracted block axes
  (every rank holds a k-slab, BASELINE config 4): each rank contracts its slab into a partial of EVERY output block and
  the partials are summed either inside the GEMM epilogue over NVLink peer memory (float64 / complex128) or by one NCCL
  reduce-scatter — both leave rank r with output blocks ``[r * chunk, (r + 1) * chunk)``;
* ``tensordot(ShardedBlockArray, BlockArray)`` with ``a`` sharded along a free block axis: a local contraction per
  rank, the result is sharded along the same output axis.

The result is a ``ShardedBlockArray``: every rank owns a contiguous range of blocks; ``np.asarray`` / ``gather()``
assemble the whole array on every rank (one all-gather).
"""
from __future__ import annotations

from math import prod
from typing import Optional, Tuple

import numpy as np

from .. import dispatch as dispatcher
from ..core.mixin import ArrayFunctionMixin
from ..core.util import normalize_axes
from ..engine import lib as _lib
from ..engine import scheduler as _sched
from .block import BlockArray
from .device import DeviceStorage, _torch, current_stream_ptr


class ShardedBlockArray(np.lib.mixins.NDArrayOperatorsMixin, ArrayFunctionMixin):
    """A BlockArray whose blocks are dealt over the ranks of a ``scheduler.Context``.

    ``grid`` / ``blockshape`` / ``dtype`` describe the WHOLE array.  Two layouts:

    * ``axis = a``: this rank holds the slab ``[first, first + count)`` of grid axis ``a`` as ``local`` (a device
      BlockArray with that slab's grid) — how k-sharded operands come in;
    * ``axis = None``: this rank holds the row-major block range ``[first, first + count)`` in ``storage`` (``chunk``
      blocks allocated, equal on every rank) — how distributed contractions hand their result back.
    """

    def __init__(self, ctx, grid, blockshape, dtype, axis: Optional[int], first: int, count: int,
                 local: Optional[BlockArray] = None, storage: Optional[DeviceStorage] = None, chunk: Optional[int] = None):
        self.ctx = ctx
        self.grid = tuple(int(g) for g in grid)
        self.blockshape = tuple(int(b) for b in blockshape)
        self.dtype = np.dtype(dtype)
        self.axis, self.first, self.count = axis, int(first), int(count)
        self.local, self.storage = local, storage
        self.chunk = int(chunk) if chunk is not None else None

    # ---- construction -----------------------------------------------------------------
    @classmethod
    def from_local_slab(cls, local: BlockArray, axis: int, ctx=None) -> "ShardedBlockArray":
        """Every rank passes its own slab (equal grids); the whole array is their concatenation along grid ``axis``."""
        ctx = ctx or _sched.context()
        if ctx is None:
            raise RuntimeError("no scheduler context: call rosnet_b200.engine.scheduler.init() first")
        axis = int(axis) % len(local.grid)
        grid = list(local.grid)
        per = grid[axis]
        grid[axis] = per * ctx.world
        return cls(ctx, grid, local.blockshape, local.dtype, axis, ctx.rank * per, per, local=local.to_device())

    @classmethod
    def shard(cls, arr: BlockArray, axis: int, ctx=None) -> "ShardedBlockArray":
        """Keep this rank's slab of a replicated BlockArray along grid ``axis`` (extent divisible by the ranks)."""
        ctx = ctx or _sched.context()
        if ctx is None:
            raise RuntimeError("no scheduler context: call rosnet_b200.engine.scheduler.init() first")
        axis = int(axis) % len(arr.grid)
        if arr.grid[axis] % ctx.world:
            raise ValueError(f"grid axis {axis} of extent {arr.grid[axis]} does not divide over {ctx.world} ranks")
        per = arr.grid[axis] // ctx.world
        sl = tuple(slice(ctx.rank * per, (ctx.rank + 1) * per) if i == axis else slice(None) for i in range(len(arr.grid)))
        host = arr.to_host() if arr.is_device else arr
        return cls(ctx, arr.grid, arr.blockshape, arr.dtype, axis, ctx.rank * per, per, local=BlockArray(host.data[sl]).to_device())

    # ---- array face --------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(g * b for g, b in zip(self.grid, self.blockshape))

    @property
    def ndim(self) -> int:
        return len(self.grid)

    @property
    def nblock(self) -> int:
        return prod(self.grid)

    @property
    def blocksize(self) -> int:
        return prod(self.blockshape)

    @property
    def nbytes(self) -> int:
        return self.nblock * self.blocksize * self.dtype.itemsize

    def __repr__(self):
        where = f"axis {self.axis}" if self.axis is not None else "flat"
        return (f"ShardedBlockArray(shape={self.shape}, grid={self.grid}, blockshape={self.blockshape}, dtype={self.dtype}, "
                f"{where} blocks [{self.first}, {self.first + self.count}) on rank {self.ctx.rank}/{self.ctx.world})")

    def owned(self) -> BlockArray:
        """The blocks this rank owns as a device BlockArray: the slab (axis layout) or a ``(count, 1, ...)`` stack."""
        if self.axis is not None:
            return self.local
        nd = len(self.grid)
        grid = ((self.count,) + (1,) * (nd - 1)) if nd else ()
        return BlockArray._from_storage(self.storage, grid, self.blockshape, self.dtype)

    def owned_indices(self):
        """Grid multi-indices of the owned blocks, in storage order."""
        if self.axis is None:
            idx = list(np.ndindex(*self.grid)) if self.grid else [()]
            return idx[self.first:self.first + self.count]
        lg = list(self.grid)
        lg[self.axis] = self.count
        return [tuple(i + (self.first if d == self.axis else 0) for d, i in enumerate(ix)) for ix in np.ndindex(*lg)]

    def gather(self) -> BlockArray:
        """The whole array on every rank: one all-gather (rank-major), then — for an axis layout whose axis is not the
        leading one — one permute launch into row-major grid order."""
        ctx = self.ctx
        torch = _torch()
        itemsize = self.dtype.itemsize
        blk = self.blocksize
        if self.axis is None:
            per = self.chunk
            src = self.storage
        else:
            per = self.local.nblock
            src = self.local._storage
        out = DeviceStorage(ctx.world * per * blk * itemsize, src.device)
        with torch.cuda.device(src.device):
            _all_gather(ctx, src, out, per * blk, self.dtype)
            if self.axis is None or self.axis == 0 or all(g == 1 for g in self.grid[:self.axis]):
                return BlockArray._from_storage(out, self.grid, self.blockshape, self.dtype)
            # gathered order: (rank, local grid..., block elements) -> row-major global grid
            lg = list(self.local.grid)
            extents = [ctx.world] + lg + [blk]
            src_strides, acc = [0] * len(extents), 1
            for i in range(len(extents) - 1, -1, -1):
                src_strides[i] = acc
                acc *= extents[i]
            gstr, acc = [0] * len(self.grid), blk
            for i in range(len(self.grid) - 1, -1, -1):
                gstr[i] = acc
                acc *= self.grid[i]
            dst_strides = [gstr[self.axis] * lg[self.axis]] + gstr + [1]
            final = DeviceStorage(out.nbytes, src.device)
            _lib.permute(out.ptr, final.ptr, itemsize, extents, src_strides, dst_strides, current_stream_ptr(src.device))
        return BlockArray._from_storage(final, self.grid, self.blockshape, self.dtype)

    def __array__(self, dtype=None, copy=None):
        out = np.asarray(self.gather())
        return out if dtype is None else out.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        raise NotImplementedError("__array_ufunc__ is not implemented yet")


def _all_gather(ctx, src: DeviceStorage, dst: DeviceStorage, count: int, dtype):
    """dst[r * count : (r + 1) * count] = rank r's src[:count] (elements of ``dtype``)."""
    comm = ctx.comm
    _lib._pre()      # queued small contractions first: the gather reads their results
    if isinstance(comm, _sched.NcclComm):
        stream = current_stream_ptr(src.device)
        _lib.check(comm._L.rnb_all_gather(src.ptr, dst.ptr, int(count), _lib.DTYPE_CODES[str(np.dtype(dtype))], comm._comm, stream), "rnb_all_gather")
        return
    torch = _torch()
    nb = count * np.dtype(dtype).itemsize
    comm.all_gather(src.tensor[:nb], dst.tensor[: nb * ctx.world])


# ------------------------------------------------------------------------------------------
# distributed tensordot
# ------------------------------------------------------------------------------------------
def _flat_result(ctx, plan, storage, chunk) -> ShardedBlockArray:
    first = ctx.rank * chunk
    count = max(0, min(chunk, plan.n_out - first))
    return ShardedBlockArray(ctx, plan.out_grid, plan.out_blockshape, plan.dtype, None, first, count, storage=storage, chunk=chunk)


def _exchange_partials(ctx, plan, run_partial, device) -> ShardedBlockArray:
    """Tier 2 tail: ``run_partial(out_storage=None | peers=...)`` computes this rank's partial of every output block;
    the partials are summed across ranks and rank r keeps blocks [r * chunk, (r + 1) * chunk)."""
    torch = _torch()
    itemsize = np.dtype(plan.dtype).itemsize
    blk_bytes = plan.out_block_elems * itemsize
    chunk = _sched.padded_chunk(plan.n_out, ctx.world)
    stream = current_stream_ptr(device)
    if ctx.use_fused(plan.dtype):
        peers = ctx.peer_buffers(chunk * blk_bytes)
        peers.zero(stream)
        ctx.barrier()                 # every owner buffer is zero before any rank adds into it
        dummy = DeviceStorage(blk_bytes, device)
        run_partial(out_storage=dummy, peers=peers.ptrs, blocks_per_peer=chunk)
        ctx.barrier()                 # all ranks' adds have landed
        mine = DeviceStorage(chunk * blk_bytes, device)
        _lib.memcpy_d2d(mine.ptr, peers.own, chunk * blk_bytes, stream)   # the peer buffer is reused by the next call
        ctx.stats["fused"] += 1
        return _flat_result(ctx, plan, mine, chunk)
    partial = DeviceStorage(ctx.world * chunk * blk_bytes, device, zero=(ctx.world * chunk != plan.n_out))
    run_partial(out_storage=partial)
    recv = DeviceStorage(chunk * blk_bytes, device)
    tdt = {"complex128": torch.complex128, "float64": torch.float64, "complex64": torch.complex64, "float32": torch.float32}[str(np.dtype(plan.dtype))]
    send_t = partial.tensor[: ctx.world * chunk * blk_bytes].view(tdt)
    recv_t = recv.tensor[: chunk * blk_bytes].view(tdt)
    ctx.comm.reduce_scatter_sum(send_t, recv_t)
    ctx.stats["nccl"] += 1
    return _flat_result(ctx, plan, recv, chunk)


def distributed_tensordot(a: BlockArray, b: BlockArray, axes, ctx=None, tier: Optional[str] = None):
    """``tensordot`` of REPLICATED operands dealt over the ranks (see the module docstring).  Returns a
    ``ShardedBlockArray`` (tiers 1 and 2) or a plain ``BlockArray`` (local)."""
    from ..engine import plan as _plan
    from ..engine import runtime
    from ..tuning import blocking

    ctx = ctx or _sched.context()
    world = _sched.active_world() if ctx is not None else 1
    dtype = np.result_type(a.dtype, b.dtype)
    p = _plan.plan_tensordot(a.grid, a.blockshape, b.grid, b.blockshape, dtype, axes)
    tier = tier or _sched.choose_tier(p.n_out, p.n_pairs, p.k, world, blocking.threshold_k())
    if ctx is None or world <= 1 or tier == "local":
        if ctx is not None:
            ctx.stats["local"] += 1
        from .block import _local_tensordot

        return _local_tensordot(a, b, axes)
    if world != ctx.world:
        raise NotImplementedError("tuning.configure(n_gpus=) must be 1 or the size of the scheduler context")
    sa, sb = a._device_storage(dtype), b._device_storage(dtype)
    itemsize = np.dtype(dtype).itemsize
    if tier == "output":
        chunk = _sched.padded_chunk(p.n_out, world)
        first = ctx.rank * chunk
        count = max(0, min(chunk, p.n_out - first))
        out = DeviceStorage(chunk * p.out_block_elems * itemsize, sa.device)
        if count:
            runtime.execute_tensordot(p, sa, a.blockshape, sb, b.blockshape, out_storage=out, n_out=count, pair_offset=first)
        ctx.stats["output"] += 1
        return _flat_result(ctx, p, out, chunk)
    if tier == "ksplit":
        mine = _sched.ksplit_plan(p, world, ctx.rank)
        ctx.stats["ksplit"] += 1
        return _exchange_partials(ctx, p, lambda **kw: runtime.execute_tensordot(mine, sa, a.blockshape, sb, b.blockshape, **kw), sa.device)
    raise ValueError(f"unknown tier {tier!r}")


@dispatcher.tensordot.register
def tensordot_sharded(a: ShardedBlockArray, b: ShardedBlockArray, axes=2):
    """Both operands sharded along a PAIR of contracted block axes: tier 2 (partials + exchange)."""
    from ..engine import plan as _plan
    from ..engine import runtime

    ctx = a.ctx
    if a.axis is None or b.axis is None:
        raise NotImplementedError("tensordot of flat-sharded results: gather() them first")
    ax_a, ax_b = normalize_axes(axes, len(a.grid), len(b.grid))
    if a.axis not in ax_a or ax_b[list(ax_a).index(a.axis)] != b.axis:
        raise NotImplementedError("operands must be sharded along axes that are contracted with each other")
    dtype = np.result_type(a.dtype, b.dtype)
    la, lb = a.local, b.local
    p = _plan.plan_tensordot(la.grid, la.blockshape, lb.grid, lb.blockshape, dtype, (list(ax_a), list(ax_b)))
    sa, sb = la._device_storage(dtype), lb._device_storage(dtype)
    ctx.stats["ksplit"] += 1
    return _exchange_partials(ctx, p, lambda **kw: runtime.execute_tensordot(p, sa, la.blockshape, sb, lb.blockshape, **kw), sa.device)


@dispatcher.tensordot.register
def tensordot_sharded_left(a: ShardedBlockArray, b: BlockArray, axes=2):
    """``a`` sharded along a FREE block axis, ``b`` replicated: tier 1, the result is sharded along that output axis."""
    from .block import _local_tensordot

    if a.axis is None:
        raise NotImplementedError("tensordot of flat-sharded results: gather() them first")
    ax_a, ax_b = normalize_axes(axes, len(a.grid), len(b.grid))
    if a.axis in ax_a:
        raise NotImplementedError("a is sharded along a contracted axis: shard b along the paired axis as well")
    local = _local_tensordot(a.local, b, (list(ax_a), list(ax_b)))
    out_axis = sorted(set(range(len(a.grid))) - set(ax_a)).index(a.axis)
    grid = list(local.grid)
    grid[out_axis] = a.grid[a.axis]
    a.ctx.stats["output"] += 1
    return ShardedBlockArray(a.ctx, grid, local.blockshape, local.dtype, out_axis, a.first, a.count, local=local)


@dispatcher.to_numpy.register
def to_numpy_sharded(arr: ShardedBlockArray) -> np.ndarray:
    return np.asarray(arr.gather())
