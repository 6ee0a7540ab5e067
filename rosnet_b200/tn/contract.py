"""Contraction driver: walks a pairwise path issuing ``rosnet_b200.tensordot`` (one grouped CUDA
launch per step, plus a permute launch when an operand is not in GEMM layout) and a final
``transpose``.  This is the role opt_einsum's ``contract`` plays for the reference with
``backend="rosnet"``.  Two slicing modes:

* ``"loop"``  — independent contractions over the values of the sliced indices, summed at the end
  (the unit the multi-GPU scheduler deals round-robin, BASELINE config 5);
* ``"blocks"`` — *slicing as blocks* (``docs/examples/tn-slicing-as-blocks.ipynb`` cell 6,
  ``benchmark/bench_qft.py:94-102``): sliced indices get block extent 1, so slices become grid
  entries of BlockArrays and the sum over a sliced index is the fused k-block sum of the GEMM.
"""
from __future__ import annotations

import copy
from math import prod
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .. import dispatch as dispatcher
from ..array.block import BlockArray, array as block_array, enqueue_to_host, tensordot as block_tensordot, transpose as block_transpose
from ..array.device import pinned_empty
from .network import TensorNetwork
from .path import greedy_path, path_cost, search_path

SEARCH_TARGET_BYTES = 2 << 30     # largest intermediate a searched tree may hold (contract_network(path="search")): 2 GiB


def _sync_current_stream():
    from ..array.device import _torch, current_stream_ptr
    from ..engine import lib as _lib

    _torch()
    _lib.stream_synchronize(current_stream_ptr())


class _SliceStager:
    """Uploads all input tensors of one slice with ONE host-to-device copy.

    A pageable ``cudaMemcpyAsync`` makes the driver synchronise with the stream before it stages the
    source, i.e. one hidden stream sync per input tensor (259 per Sycamore slice, measured) and the
    host can never run ahead of the GPU.  Here the tensors are packed into a pinned staging buffer
    (a ring, so slice i+1 can be packed while the copy of slice i is still in flight), copied once
    into one device arena, and handed to the contraction as single-block BlockArrays that are views
    of that arena."""

    ALIGN = 256

    def __init__(self, depth: int = 3):
        self.depth = depth
        self.slots = []   # (pinned uint8 ndarray, torch event or None)
        self.turn = 0

    @classmethod
    def layout(cls, arrays):
        offsets, total = [], 0
        for a in arrays:
            offsets.append(total)
            total += -(-max(a.nbytes, 1) // cls.ALIGN) * cls.ALIGN
        return offsets, total

    def copy_in(self, arrays, arena=None):
        """Pack ``arrays`` into the next pinned slot and queue ONE H2D copy into ``arena`` (a fresh one
        if None).  Returns (arena, offsets)."""
        from ..array.device import DeviceStorage, _torch, current_stream_ptr
        from ..engine import lib as _lib

        torch = _torch()
        offsets, total = self.layout(arrays)
        slot = self.turn % self.depth
        self.turn += 1
        if slot >= len(self.slots):
            self.slots.append([pinned_empty((total,), np.uint8), None])
        buf, ev = self.slots[slot]
        if ev is not None:
            ev.synchronize()          # the copy that last read this slot (depth slices ago) has finished
        if buf.nbytes < total:
            buf = pinned_empty((total,), np.uint8)
            self.slots[slot][0] = buf
        for a, off in zip(arrays, offsets):
            if a.nbytes:
                buf[off: off + a.nbytes].view(a.dtype).reshape(a.shape)[...] = a
        if arena is None:
            arena = DeviceStorage(total)
        _lib.memcpy_h2d(arena.ptr, buf.ctypes.data, total, current_stream_ptr(arena.device))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(arena.device))
        self.slots[slot][1] = ev
        return arena, offsets

    @staticmethod
    def views(arena, offsets, arrays, labels=None, block_sliced=()):
        """BlockArrays over the arena.  ``labels`` / ``block_sliced``: slicing-as-blocks — axes whose label
        is in ``block_sliced`` get block extent 1 (one device permute dense -> block-major per such
        tensor, no synchronisation).  ``arrays`` only supplies shapes and dtypes (0-d: the values)."""
        from ..array.device import DeviceStorage
        from ..engine import runtime

        out = []
        for n, (a, off) in enumerate(zip(arrays, offsets)):
            if a.ndim == 0:
                out.append(BlockArray(np.asarray(a)))
                continue
            st = DeviceStorage.view(arena, off, a.nbytes)
            bs = a.shape
            if labels is not None and block_sliced:
                bs = tuple(1 if lab in block_sliced else e for lab, e in zip(labels[n], a.shape))
            if bs == a.shape:
                out.append(BlockArray._from_storage(st, (1,) * a.ndim, a.shape, a.dtype))
            else:
                grid = tuple(e // b for e, b in zip(a.shape, bs))
                blocked = DeviceStorage(a.nbytes, arena.device)
                runtime.dense_to_blocks(st.ptr, blocked, grid, bs, a.dtype.itemsize)
                out.append(BlockArray._from_storage(blocked, grid, bs, a.dtype))
        return out

    def upload(self, arrays, labels=None, block_sliced=()):
        arena, offsets = self.copy_in(arrays)
        out = self.views(arena, offsets, arrays, labels, block_sliced)
        self.last_arena = arena
        return out


class _Resident(dict):
    """``{slice id: [BlockArray, ...]}`` plus the device arena each list is a view of."""

    def __init__(self):
        super().__init__()
        self.arenas = {}


class _GraphedPass:
    """One pass over the path captured as a CUDA graph (B200 playbook: capture launch-bound inner
    loops).  Every slice of a network issues the same ~700 launches with the same shapes; only the
    contents of the small input tensors differ.  The inputs live in ONE static device arena, the
    pass is captured once (after an eager pass has warmed the plan / table / attribute caches, so
    nothing inside the capture touches pageable memory) and every further slice is: one H2D into the
    arena, ``graph.replay()``, one D2H of the result — three driver calls instead of ~2000 Python-level
    operations.  Anything that prevents capture falls back to the eager pass."""

    def __init__(self):
        self.graph = None
        self.arena = None
        self.offsets = None
        self.result = None
        self.failed = False
        self.warm = False      # an eager pass has run (caches filled): the next pass may be captured

    def capture(self, arena, offsets, arrays, labels, output, path, inner, block_sliced=()):
        from ..array.device import _torch

        torch = _torch()
        if any(a.ndim == 0 for a in arrays):
            self.failed = True      # 0-d operands are host constants: they would be frozen into the graph
            return False
        from ..engine import lib as _lib

        from ..engine import runtime as _rt

        before = dict(_lib.launch_counter)
        _rt.capture_refs = refs = []     # cached device tables the recorded launches point at: the graph keeps them alive
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread_local: CUDA calls of other threads (the NCCL watchdog of torch.distributed) do not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                views = _SliceStager.views(arena, offsets, arrays, labels, block_sliced)
                res = _contract_once(labels, output, views, path, inner, block_sliced=set(block_sliced))
            if not res.is_device:
                raise RuntimeError("result is not device resident")
        except Exception:
            self.failed = True
            _rt.capture_refs = None
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            return False
        _rt.capture_refs = None
        self.graph, self.arena, self.offsets, self.result, self.refs = g, arena, offsets, res, refs
        # capture only records the launches: nothing ran, so they are not counted until a replay
        self.counts = {k: _lib.launch_counter[k] - before.get(k, 0) for k in _lib.launch_counter}
        for k, v in self.counts.items():
            _lib.launch_counter[k] -= v
        return True

    def replay(self):
        from ..engine import lib as _lib

        self.graph.replay()
        for k, v in self.counts.items():
            _lib.launch_counter[k] += v
        return self.result


def _clone_storage(storage):
    from ..array.device import DeviceStorage

    out = DeviceStorage(storage.tensor.numel(), storage.device)
    out.tensor.copy_(storage.tensor, non_blocking=True)
    return out


_graph_cache = {}


def _graphed_pass(tn, path, sliced, mode) -> _GraphedPass:
    from ..engine import runtime

    # the arithmetic settings are frozen into a captured graph (kernel variants, descriptors)
    key = (id(tn), mode, tuple(sliced), len(path), tuple(sorted(runtime.settings.items())))
    hit = _graph_cache.get(key)
    if hit is None or hit[0] is not tn or hit[1] != tuple(map(tuple, path)):
        if len(_graph_cache) >= 4:     # every graph pins the memory pool of one whole pass
            _graph_cache.clear()
        hit = _graph_cache[key] = (tn, tuple(map(tuple, path)), _GraphedPass())
    return hit[2]


_stager = None


def _shared_stager() -> "_SliceStager":
    """One long-lived stager: its pinned ring must outlive the asynchronous copies that read it."""
    global _stager
    if _stager is None:
        _stager = _SliceStager(depth=3)
    return _stager


def _to_block(arr, labels, sliced, inner):
    if isinstance(arr, BlockArray):
        return arr
    arr = np.asarray(arr)
    if arr.ndim == 0:
        return BlockArray(arr)
    bs = tuple(1 if l in sliced else s for l, s in zip(labels, arr.shape))
    return block_array(arr, blockshape=bs, inner=inner)


def _contract_once(inputs, output, arrays, path, inner, block_sliced=()):
    """One pass over the path.  Returns the result BlockArray.  The tiny steps are queued and launched one dependency level
    at a time (``engine.runtime.batching``)."""
    from ..engine import runtime

    with runtime.batching():
        return _contract_once_unbatched(inputs, output, arrays, path, inner, block_sliced)


def _contract_once_unbatched(inputs, output, arrays, path, inner, block_sliced=()):
    tensors: Dict[int, Tuple[BlockArray, Tuple[int, ...]]] = {}
    count: Dict[int, int] = {}
    for t, (ix, arr) in enumerate(zip(inputs, arrays)):
        tensors[t] = (_to_block(arr, ix, block_sliced, inner), tuple(ix))
        for i in ix:
            count[i] = count.get(i, 0) + 1
    for i in output:
        count[i] = count.get(i, 0) + 1
    next_id = len(inputs)
    for a, b in path:
        (A, ia), (B, ib) = tensors.pop(a), tensors.pop(b)
        shared = [i for i in ia if i in ib]
        contracted = [i for i in shared if count[i] == 2]
        if len(contracted) != len(shared):
            raise NotImplementedError("indices shared by more than two tensors (hyper-edges / batch indices) need einsum-style batching")
        for i in contracted:
            count[i] -= 2
        ax_a = [ia.index(i) for i in contracted]
        ax_b = [ib.index(i) for i in contracted]
        C = block_tensordot(A, B, (ax_a, ax_b))
        if not isinstance(C, BlockArray):     # dealt over the ranks by the scheduler: the next step needs it whole
            C = C.gather()
        labels = tuple(i for i in ia if i not in contracted) + tuple(i for i in ib if i not in contracted)
        tensors[next_id] = (C, labels)
        next_id += 1
    if len(tensors) != 1:
        raise ValueError("path does not contract the network to a single tensor")
    (res, labels), = tensors.values()
    if tuple(labels) != tuple(output):
        perm = [labels.index(i) for i in output]
        res = block_transpose(res, perm, inplace=False)
    return res


def _slice_inputs(tn: TensorNetwork, sliced, extents, sid):
    """Host arrays and index labels of slice number ``sid`` (sliced indices fixed to its values)."""
    values = np.unravel_index(sid, extents) if extents else ()
    sub_inputs, sub_arrays = [], []
    for ix, arr in zip(tn.inputs, tn.arrays):
        arr = np.asarray(arr)
        keep = []
        for lab in ix:
            if lab in sliced:
                arr = np.take(arr, int(values[sliced.index(lab)]), axis=len(keep))
            else:
                keep.append(lab)
        sub_inputs.append(tuple(keep))
        sub_arrays.append(np.ascontiguousarray(arr))
    return sub_inputs, sub_arrays


def stage_slices(tn: TensorNetwork, sliced: Sequence[int], slice_ids: Sequence[int]):
    """Upload the input tensors of the given slices once and keep them resident in HBM: returns
    ``{slice id: [BlockArray, ...]}`` for ``contract_network(..., resident=...)``.  Used to time
    the contraction alone (operands already on the device) next to the host-to-host path."""
    sliced = list(sliced)
    extents = [tn.sizes[i] for i in sliced]
    stager = _SliceStager(depth=max(len(set(slice_ids)), 1))
    out = _Resident()
    for sid in dict.fromkeys(slice_ids):
        _, arrays = _slice_inputs(tn, sliced, extents, sid)
        out[sid] = stager.upload(arrays)
        out.arenas[sid] = stager.last_arena
    _sync_current_stream()
    return out


_search_cache: Dict[tuple, tuple] = {}      # (network structure, dtype, seed, memory limit) -> (path, sliced): the search costs seconds of host time


def _searched(tn, rdtype, seed):
    path, sliced, _info = search_path(tn.inputs, tn.output, tn.sizes, trials=64, seed=seed, target_size=SEARCH_TARGET_BYTES // rdtype.itemsize,
                                      minimize="time", itemsize=rdtype.itemsize,
                                      dtype=str(rdtype) if str(rdtype) in ("float32", "float64", "complex64", "complex128") else None)
    return path, sliced


def contract_network(tn: TensorNetwork, path: Optional[Sequence[Tuple[int, int]]] = None, sliced: Sequence[int] = (),
                     slice_mode: str = "loop", slice_ids: Optional[Sequence[int]] = None, inner: str = "cuda", seed: int = 0,
                     resident=None, fetch: bool = True, graph: Optional[bool] = None, comm=None):
    """Contract ``tn`` (arrays attached) along ``path`` (the plain greedy tree if None; ``"search"``: the best of 64 seeded
    greedy trees under the time model, sliced by the search itself until every intermediate fits ``SEARCH_TARGET_BYTES``).

    ``sliced``: index labels to slice.  ``slice_mode="loop"``: returns the SUM over the slices in
    ``slice_ids`` (all if None) as a dense ndarray — callers that split slices over ranks add the
    per-rank sums.  ``slice_mode="blocks"``: one pass with the sliced indices blocked at extent 1.
    ``resident``: inputs already on the device (``stage_slices``); ``fetch=False`` leaves every
    slice's result on the device and returns the list of them without synchronising.
    ``graph``: replay the pass as a CUDA graph (``_GraphedPass``).  Default: on for slice loops of three
    or more slices (first slice eager, second captured, the rest replayed), off for single passes
    (``graph=True`` there pays off when the same network is contracted repeatedly).
    ``comm``: a ``rosnet_b200.engine.scheduler.Context`` (``scheduler.init()`` under torchrun, one process per GPU).
    ``slice_mode="loop"``: the slices (``slice_ids``, all if None) are dealt round-robin over the ranks, every rank
    contracts its share on its own GPU and ONE all-reduce (NCCL) of the summed result closes the call: every rank
    returns the same total.  (``slice_mode="blocks"`` with an active context: every pairwise ``tensordot`` is dealt by the
    scheduler itself, ``array/sharded.py``.)
    Returns (result, stats) with stats = dict(flops=..., steps=..., largest=..., n_slices=...)."""
    if isinstance(path, str):
        if path != "search":
            raise ValueError("path must be a list of SSA pairs, None (plain greedy tree) or 'search'")
        if not tn.arrays:
            raise ValueError("network has no arrays attached")
        # searched tree + its own slicing (tn/path.py: search_path, objective: modelled time); every rank finds the same one
        if sliced:
            raise ValueError("path='search' picks the sliced indices itself: pass `sliced` only with an explicit path")
        rdtype = np.result_type(*[a.dtype for a in tn.arrays])
        key = (tuple(tuple(ix) for ix in tn.inputs), tuple(tn.output), tuple(sorted(tn.sizes.items())), str(rdtype), seed, SEARCH_TARGET_BYTES)
        hit = _search_cache.get(key)
        if hit is not None:
            path, sliced = hit
        else:
            path, sliced = _searched(tn, rdtype, seed)
            if len(_search_cache) >= 64:
                _search_cache.pop(next(iter(_search_cache)))
            _search_cache[key] = (path, sliced)
    if comm is not None and slice_mode == "loop" and sliced:
        return _contract_slices_distributed(tn, path, sliced, slice_ids, inner, seed, resident, fetch, graph, comm)
    if not tn.arrays:
        raise ValueError("network has no arrays attached")
    if path is None:
        path = greedy_path(tn.inputs, tn.output, tn.sizes, seed=seed)
    dtype = np.result_type(*[a.dtype for a in tn.arrays])
    per_madd = 8 if dtype.kind == "c" else 2
    sliced = list(sliced)
    if any(i in tn.output for i in sliced):
        raise ValueError("open indices cannot be sliced")
    if slice_mode == "blocks" or not sliced:
        madds, largest, steps = path_cost(tn.inputs, tn.output, tn.sizes, path)
        arrays = tn.arrays
        stats = dict(flops=per_madd * madds, steps=len(steps), largest=largest, n_slices=1)
        if inner == "cuda" and not any(isinstance(a, BlockArray) for a in arrays):
            # one pinned H2D for all inputs (a pageable copy per tensor is a hidden stream sync each)
            host = [np.ascontiguousarray(a) for a in arrays]
            gp = _graphed_pass(tn, path, sliced, "blocks") if graph else None
            if gp is not None and gp.graph is not None:
                _shared_stager().copy_in(host, arena=gp.arena)
                return copy.deepcopy(gp.replay()), stats      # the graph's result buffer is re-used by the next replay
            if gp is not None and gp.warm and not gp.failed:
                arena, offsets = _shared_stager().copy_in(host)
                if gp.capture(arena, offsets, host, tn.inputs, tn.output, path, inner, block_sliced=set(sliced)):
                    return copy.deepcopy(gp.replay()), stats
            if gp is not None:
                gp.warm = True
            arrays = _shared_stager().upload(host, labels=tn.inputs, block_sliced=set(sliced))
        res = _contract_once(tn.inputs, tn.output, arrays, path, inner, block_sliced=set(sliced))
        return res, stats
    if slice_mode != "loop":
        raise ValueError("slice_mode must be 'loop' or 'blocks'")
    extents = [tn.sizes[i] for i in sliced]
    n_slices = prod(extents)
    ids = range(n_slices) if slice_ids is None else slice_ids
    madds, largest, steps = path_cost(tn.inputs, tn.output, tn.sizes, path, sliced)
    total = None
    done = 0
    ids = list(ids)
    # Slices are independent, so nothing forces the host to wait for slice i before issuing slice
    # i+1: results are copied asynchronously into a ring of pinned slots and added (in slice order,
    # in double precision) only when the ring is full or at the end.  The GPU then never idles
    # while Python walks the next slice's path.
    out_shape = tuple(tn.sizes[i] for i in tn.output)
    wide = np.result_type(dtype, np.complex128 if dtype.kind == "c" else np.float64)
    per_slice = max(prod(out_shape) * dtype.itemsize, 1)
    window = int(max(1, min(len(ids), (256 << 20) // per_slice)))
    ring = None
    pending: List[int] = []
    stager = _shared_stager() if inner == "cuda" else None

    def drain():
        nonlocal total
        if not pending:
            return
        _sync_current_stream()
        for slot in pending:
            val = np.asarray(ring[slot]).astype(wide)
            total = val if total is None else total + val
        pending.clear()

    kept = []
    sub_labels = None
    use_graph = (len(ids) >= 3) if graph is None else bool(graph)
    gp = _graphed_pass(tn, path, sliced, "loop") if (use_graph and inner == "cuda") else None
    for sid in ids:
        src_arena = None
        if resident is not None and sid in resident:
            if sub_labels is None:
                sub_labels = [tuple(l for l in ix if l not in sliced) for ix in tn.inputs]
            sub_inputs, sub_arrays = sub_labels, resident[sid]
            src_arena = getattr(resident, "arenas", {}).get(sid)
        else:
            sub_inputs, sub_arrays = _slice_inputs(tn, sliced, extents, sid)
        is_host = src_arena is None and not (resident is not None and sid in resident)
        res = None
        if gp is not None and not gp.failed and (is_host or src_arena is not None):
            if gp.graph is None and gp.warm:
                if is_host:
                    arena, offsets = stager.copy_in(sub_arrays)
                else:
                    arena, offsets = _clone_storage(src_arena), _SliceStager.layout(sub_arrays)[0]
                gp.capture(arena, offsets, sub_arrays, sub_inputs, tn.output, path, inner)
                if gp.graph is not None:
                    res = gp.replay()
            elif gp.graph is not None:
                if is_host:
                    stager.copy_in(sub_arrays, arena=gp.arena)
                else:
                    n = min(gp.arena.tensor.numel(), src_arena.tensor.numel())
                    gp.arena.tensor[:n].copy_(src_arena.tensor[:n], non_blocking=True)
                res = gp.replay()
            if res is not None and not fetch:
                res = copy.deepcopy(res)          # the graph's result buffer is re-used by the next replay
        if res is None:
            if is_host and stager is not None:
                sub_arrays = stager.upload(sub_arrays)
            res = _contract_once(sub_inputs, tn.output, sub_arrays, path, inner)
            if gp is not None:
                gp.warm = True
        if not fetch:
            kept.append(res)
            done += 1
            continue
        if res.is_device and res.dtype == dtype and res.shape == out_shape:
            if ring is None:
                ring = pinned_empty((window,) + out_shape, dtype)
            if len(pending) == window:
                drain()
            slot = len(pending)
            enqueue_to_host(res, ring[slot:slot + 1].reshape(out_shape))   # a view also when the result is 0-d
            pending.append(slot)
        else:
            drain()
            val = np.asarray(res).astype(wide)
            total = val if total is None else total + val
        done += 1
    drain()
    if not fetch:
        total = kept
    return total, dict(flops=per_madd * madds * done, steps=len(steps), largest=largest, n_slices=done, total_slices=n_slices)


def _contract_slices_distributed(tn, path, sliced, slice_ids, inner, seed, resident, fetch, graph, ctx):
    """Tier 3 of the single-box scheduler: independent slices dealt round-robin, one all-reduce of the sum."""
    from .. import tuning
    from ..array.device import _torch
    from ..engine import scheduler as _sched

    if not fetch:
        raise ValueError("fetch=False keeps per-slice results on the device and cannot be combined with comm=")
    if path is None:
        path = greedy_path(tn.inputs, tn.output, tn.sizes, seed=seed)
    n_slices = prod(tn.sizes[i] for i in sliced)
    ids = list(range(n_slices)) if slice_ids is None else list(slice_ids)
    mine = [ids[i] for i in _sched.round_robin(len(ids), ctx.world, ctx.rank)]
    dtype = np.result_type(*[a.dtype for a in tn.arrays])
    wide = np.result_type(dtype, np.complex128 if dtype.kind == "c" else np.float64)
    out_shape = tuple(tn.sizes[i] for i in tn.output)
    total, stats = None, dict(flops=0, steps=0, largest=0, n_slices=0, total_slices=n_slices)
    if mine:
        with tuning.configure(n_gpus=1):      # the pairwise steps of a slice stay on this rank's GPU
            total, stats = contract_network(tn, path, sliced=sliced, slice_mode="loop", slice_ids=mine, inner=inner, seed=seed,
                                            resident=resident, fetch=True, graph=graph)
    if total is None:
        total = np.zeros(out_shape, dtype=wide)
    torch = _torch()
    host = np.ascontiguousarray(np.asarray(total, dtype=wide)).reshape(-1)
    dev = torch.from_numpy(host.copy()).to("cuda" if torch.cuda.is_available() else "cpu")
    ctx.comm.all_reduce_sum(dev)
    total = dev.cpu().numpy().reshape(out_shape)
    ctx.stats["slices"] += len(mine)
    stats = dict(stats)
    stats["n_slices_all_ranks"] = len(ids)
    return total, stats


# ------------------------------------------------------------------------------------------
# einsum for BlockArray operands (the reference registers none: np.einsum(BlockArray...) raises
# NotImplementedError there)
# ------------------------------------------------------------------------------------------
def _parse(subscripts: str, n_ops: int):
    subscripts = subscripts.replace(" ", "")
    if "..." in subscripts:
        raise NotImplementedError("ellipsis in einsum subscripts")
    if "->" in subscripts:
        lhs, out = subscripts.split("->")
    else:
        lhs = subscripts
        flat = lhs.replace(",", "")
        out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
    terms = lhs.split(",")
    if len(terms) != n_ops:
        raise ValueError("number of einsum subscripts does not match the number of operands")
    return terms, out


def _to_device_blocks(x: BlockArray) -> BlockArray:
    return x if x.is_device else x.to_device()


def _diagonal(op: BlockArray, p: int, q: int) -> BlockArray:
    """Collapse the repeated axes p < q of one operand onto p (``np.einsum('..i..i..->..i..')``)."""
    from ..engine import runtime

    op = _to_device_blocks(op)
    if op.grid[p] != op.grid[q] or op.blockshape[p] != op.blockshape[q]:
        op = _reblock(op, {p: 1, q: 1})
    st = runtime.diagonal_blocks(op._storage, op.grid, op.blockshape, p, q, op.dtype.itemsize)
    keep = [i for i in range(op.ndim) if i != q]
    return BlockArray._from_storage(st, tuple(op.grid[i] for i in keep), tuple(op.blockshape[i] for i in keep), op.dtype)


def _reblock(op: BlockArray, new_extent: Dict[int, int]) -> BlockArray:
    """Same array, block extent ``new_extent[axis]`` along the given axes."""
    from ..engine import runtime

    nbs = tuple(new_extent.get(i, b) for i, b in enumerate(op.blockshape))
    if nbs == tuple(op.blockshape):
        return op
    op = _to_device_blocks(op)
    st = runtime.reblock(op._storage, op.grid, op.blockshape, nbs, op.dtype.itemsize)
    return BlockArray._from_storage(st, tuple(s // b for s, b in zip(op.shape, nbs)), nbs, op.dtype)


def _sum_axis(op: BlockArray, axis: int) -> BlockArray:
    """Sum over one axis: a contraction with a vector of ones blocked like that axis."""
    from ..array.block import full

    ones = full((op.shape[axis],), 1, dtype=op.dtype, blockshape=(op.blockshape[axis],), inner="cuda")
    return block_tensordot(op, ones, ([axis], [0]))


def _match_blocks(A: BlockArray, ia, B: BlockArray, ib, contracted) -> BlockArray:
    """``B`` re-cut so that its block extents along the contracted indices equal ``A``'s (operands of an einsum need not
    agree on their blocking; ``tensordot`` itself insists on it, like the reference)."""
    want = {ib.index(i): A.blockshape[ia.index(i)] for i in contracted if B.blockshape[ib.index(i)] != A.blockshape[ia.index(i)]}
    return _reblock(B, want) if want else B


def _batched_tensordot(A: BlockArray, ia, B: BlockArray, ib, batch, contracted):
    """One pairwise einsum step with kept shared indices (``batch``): both operands are transposed so the batch axes
    lead, cut to block extent 1 along them, and ONE grouped launch runs every batch block
    (``engine.plan.plan_batched``).  Returns (C, labels) with labels = batch + free(a) + free(b)."""
    from ..engine import plan as _plan
    from ..engine import runtime

    def lead(X, ix):
        order = [ix.index(i) for i in batch] + [d for d, i in enumerate(ix) if i not in batch]
        if order != list(range(len(ix))):
            X = block_transpose(_to_device_blocks(X), order, inplace=False)
        X = _reblock(_to_device_blocks(X), {d: 1 for d in range(len(batch))})
        return X, tuple(ix[d] for d in order)

    A, ia = lead(A, list(ia))
    B, ib = lead(B, list(ib))
    B = _match_blocks(A, ia, B, ib, contracted)
    nb = len(batch)
    if A.grid[:nb] != B.grid[:nb]:
        raise ValueError("operands disagree on the extent of a shared index")
    dtype = np.result_type(A.dtype, B.dtype)
    ax_a = tuple(ia.index(i) - nb for i in contracted)
    ax_b = tuple(ib.index(i) - nb for i in contracted)
    p = _plan.plan_batched(tuple(A.grid[:nb]), tuple(A.grid[nb:]), tuple(A.blockshape[nb:]), tuple(B.grid[nb:]), tuple(B.blockshape[nb:]),
                           str(np.dtype(dtype)), (ax_a, ax_b))
    out = runtime.execute_tensordot(p, A._device_storage(dtype), tuple(A.blockshape[nb:]), B._device_storage(dtype), tuple(B.blockshape[nb:]))
    C = BlockArray._from_storage(out, p.out_grid, p.out_blockshape, dtype)
    labels = tuple(batch) + tuple(i for i in ia[nb:] if i not in contracted) + tuple(i for i in ib[nb:] if i not in contracted)
    return C, labels


@dispatcher.einsum.register
def einsum(subscripts: str, *operands, optimize="greedy", **kwargs):
    """``np.einsum`` over BlockArrays (the reference registers none for BlockArray - ``np.einsum(BlockArray...)`` raises
    there - while its COMPSs arrays hand any pattern to one ``np.einsum`` task per block,
    ``rosnet/array/compss/__init__.py:554-568``, ``task/einsum.py:8-11``).  Any explicit or implicit pattern without
    ellipsis: an index repeated INSIDE an operand takes the diagonal (one strided-copy launch); an index carried by a
    single operand and absent from the output is summed; along a greedy pairwise path, shared indices that nobody else
    needs are contracted (one blocked ``tensordot`` per step) and shared indices that are kept - output / batch indices,
    hyper-edges - run as ONE batched grouped launch per step; a final ``transpose`` orders the output."""
    ops = [o if isinstance(o, BlockArray) else BlockArray(np.asarray(o)) for o in operands]
    terms, out = _parse(subscripts, len(ops))
    if len(set(out)) != len(out):
        raise NotImplementedError("repeated index in the einsum output")
    labels = {c: k for k, c in enumerate(sorted(set("".join(terms) + out)))}
    if any(c not in "".join(terms) for c in out):
        raise ValueError("einsum output names an index no operand has")
    inputs: List[Tuple[int, ...]] = []
    sizes: Dict[int, int] = {}
    for n, (term, op) in enumerate(zip(terms, ops)):
        if len(term) != op.ndim:
            raise ValueError(f"operand has {op.ndim} dimensions but subscripts '{term}' name {len(term)}")
        ix = [labels[c] for c in term]
        for i, s_ in zip(ix, op.shape):
            if sizes.setdefault(i, s_) != s_:
                raise ValueError("operands disagree on the extent of an index")
        while len(set(ix)) != len(ix):          # diagonal of a repeated index
            q = next(d for d in range(len(ix)) if ix[d] in ix[:d])
            op = _diagonal(op, ix.index(ix[q]), q)
            del ix[q]
        ops[n] = op
        inputs.append(tuple(ix))
    output = tuple(labels[c] for c in out)
    # indices only one operand carries and the output does not want: sum them out first
    flat = [i for ix in inputs for i in ix]
    for n in range(len(ops)):
        for i in [i for i in inputs[n] if flat.count(i) == 1 and i not in output]:
            ops[n] = _sum_axis(ops[n], inputs[n].index(i))
            inputs[n] = tuple(j for j in inputs[n] if j != i)
    if len(ops) == 1:
        perm = [inputs[0].index(i) for i in output]
        return block_transpose(ops[0], perm, inplace=False)
    path = greedy_path(inputs, output, sizes, seed=0)
    tensors: Dict[int, Tuple[BlockArray, Tuple[int, ...]]] = {t: (op, tuple(ix)) for t, (op, ix) in enumerate(zip(ops, inputs))}
    count: Dict[int, int] = {}
    for ix in inputs:
        for i in ix:
            count[i] = count.get(i, 0) + 1
    for i in output:
        count[i] = count.get(i, 0) + 1
    next_id = len(ops)
    for a, b in path:
        (A, ia), (B, ib) = tensors.pop(a), tensors.pop(b)
        shared = [i for i in ia if i in ib]
        contracted = [i for i in shared if count[i] == 2]
        batch = [i for i in shared if count[i] > 2]
        for i in contracted:
            count[i] -= 2
        for i in batch:
            count[i] -= 1                         # two carriers become one
        if batch:
            C, lab = _batched_tensordot(A, ia, B, ib, batch, contracted)
        else:
            B = _match_blocks(A, list(ia), B, list(ib), contracted)
            C = block_tensordot(A, B, ([ia.index(i) for i in contracted], [ib.index(i) for i in contracted]))
            if not isinstance(C, BlockArray):
                C = C.gather()
            lab = tuple(i for i in ia if i not in contracted) + tuple(i for i in ib if i not in contracted)
        tensors[next_id] = (C, lab)
        next_id += 1
    (res, lab), = tensors.values()
    for i in [i for i in lab if i not in output]:   # a hyper-index whose last carrier is the result itself
        res = _sum_axis(res, lab.index(i))
        lab = tuple(j for j in lab if j != i)
    if tuple(lab) != tuple(output):
        res = block_transpose(res, [lab.index(i) for i in output], inplace=False)
    return res
