"""Execution of planned contractions and layout moves on the current CUDA stream.

Everything here is asynchronous with respect to the host, like COMPSs task submission in the
reference (``rosnet/tuning/task.py:49-63``): results are device buffers with work pending on
the stream; the synchronisation points are ``__array__`` / ``to_numpy`` only.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from math import prod

import numpy as np

from ..array.device import DeviceStorage, _torch, current_stream_ptr
from . import lib as _lib
from .plan import ContractionPlan, OperandPlan, matricize_desc

# device copies of the pair tables, keyed by (id(plan), device index)
_table_cache = {}
# tunables (rosnet_b200.tuning.configure)
settings = {"complex_mode": _lib.RNB_CPLX_AUTO, "precision": _lib.RNB_PREC_DEFAULT}
if __import__("os").environ.get("RNB_PRECISION") == "tf32_bf16x2":   # whole-process opt-in (tests / experiments)
    settings["precision"] = _lib.RNB_PREC_TF32_BF16X2
if __import__("os").environ.get("RNB_COMPLEX_MODE") in ("4m", "3m", "auto"):
    settings["complex_mode"] = {"4m": _lib.RNB_CPLX_4M, "3m": _lib.RNB_CPLX_3M, "auto": _lib.RNB_CPLX_AUTO}[__import__("os").environ["RNB_COMPLEX_MODE"]]


# While a CUDA graph is being captured (tn/contract.py: _GraphedPass) this is a list that collects every cached device
# tensor whose ADDRESS the recorded launches carry; the graph keeps the list, so evicting the cache later cannot free
# memory a replay still reads.
capture_refs = None


def _nd_offsets(ext, stride) -> np.ndarray:
    """Offsets an (extent, stride) list addresses, row-major (last axis fastest)."""
    off = np.zeros((), dtype=np.int64)
    for e, s in zip(ext, stride):
        off = np.add.outer(off, np.arange(e, dtype=np.int64) * s)
    return np.asarray(off).reshape(-1)


def _device_tables(plan: ContractionPlan, device):
    """Device copies of the plan's tables: (pair_a, pair_b, nd) with nd = the four offset tables of rnb_contract_nd in
    one int32 tensor (m entries, n entries, k entries of a, k entries of b) for small contractions, else None."""
    torch = _torch()
    key = (id(plan), device.index)
    hit = _table_cache.get(key)
    if hit is None or hit[0] is not plan:
        ta = torch.from_numpy(plan.pair_a.reshape(-1)).to(device)
        tb = torch.from_numpy(plan.pair_b.reshape(-1)).to(device)
        nd = None
        if plan.nd is not None and not plan.nd_apply:
            ext_m, sa_m, ext_n, sb_n, ext_k, sa_k, sb_k = plan.nd
            tabs = np.concatenate([_nd_offsets(ext_m, sa_m), _nd_offsets(ext_n, sb_n), _nd_offsets(ext_k, sa_k), _nd_offsets(ext_k, sb_k)])
            if tabs.max(initial=0) < 2 ** 31:
                nd = torch.from_numpy(tabs.astype(np.int32)).to(device)
        while len(_table_cache) >= 4096:      # evict the oldest entries one by one (dicts keep insertion order)
            _table_cache.pop(next(iter(_table_cache)))
        hit = _table_cache[key] = (plan, ta, tb, nd)
    if capture_refs is not None:
        capture_refs.append(hit)
    return hit[1], hit[2], hit[3]


_matricize_cache = {}


def _matricize_desc(op: OperandPlan, blockshape, n_free, itemsize):
    key = (id(op), blockshape, n_free, itemsize)
    hit = _matricize_cache.get(key)
    if hit is None or hit[0] is not op:
        extents, src, dst = matricize_desc(blockshape, n_free, op.perm, op.ld, op.block_stride, op.nblocks)
        if len(_matricize_cache) > 8192:
            _matricize_cache.clear()
        hit = _matricize_cache[key] = (op, _lib.permute_desc(itemsize, extents, src, dst))
    return hit[1]


def _matricize(op: OperandPlan, blockshape, n_free, storage: DeviceStorage, itemsize, stream) -> DeviceStorage:
    ws = DeviceStorage(op.ws_elems * itemsize, storage.device, zero=op.zero_fill)
    _lib.permute_launch(_matricize_desc(op, blockshape, n_free, itemsize), storage.ptr, ws.ptr, stream)
    return ws


def execute_tensordot(plan: ContractionPlan, a_storage: DeviceStorage, a_blockshape, b_storage: DeviceStorage, b_blockshape,
                      out_storage: DeviceStorage | None = None, accumulate: bool = False, out_index=None, n_out=None,
                      pair_offset: int = 0, out_first_block: int = 0, peers=None, blocks_per_peer: int = 0) -> DeviceStorage:
    """Run one planned blocked contraction; returns the block-major output storage.

    ``n_out`` / ``pair_offset`` run a subset of the output blocks (rows ``pair_offset ..
    pair_offset + n_out`` of the pair tables) — used by the multi-GPU scheduler and by the
    pipelined host-operand path; they are written to output blocks ``out_first_block + g`` (or
    ``out_index[g]`` when given)."""
    torch = _torch()
    device = a_storage.device
    itemsize = np.dtype(plan.dtype).itemsize
    if plan.nd is not None and peers is None and os.environ.get("RNB_SIMT") != "0" and not (plan.nd_apply and os.environ.get("RNB_APPLY") == "0"):
        return _execute_nd(plan, a_storage, b_storage, out_storage, accumulate, out_index, n_out, pair_offset, out_first_block)
    single = plan.dtype in ("float32", "complex64")
    # float32 / complex64: an operand that needs the matricize pass gets it FUSED with the GEMM's split pre-pass
    # (rnb_permute_split writes the hi / lo planes straight into the contraction workspace); decided below, once the
    # descriptor exists.  RNB_FUSE_SPLIT=0 keeps the two-pass route (experiments, A/B timing).
    fuse = single and peers is None and not (plan.a.direct and plan.b.direct) and os.environ.get("RNB_FUSE_SPLIT") != "0"
    with torch.cuda.device(device):
        stream = current_stream_ptr(device)
        a_buf, b_buf = a_storage, b_storage
        if not plan.a.direct and not fuse:
            a_buf = _matricize(plan.a, a_blockshape, len(plan.free_a), a_storage, itemsize, stream)
        if not plan.b.direct and not fuse:
            b_buf = _matricize(plan.b, b_blockshape, len(plan.free_b), b_storage, itemsize, stream)
        n_out_total = plan.n_out
        if out_storage is None:
            out_storage = DeviceStorage((n_out_total if peers is None else 1) * plan.out_block_elems * itemsize, device)
        ta, tb, _ = _device_tables(plan, device)
        d = _lib.ContractDesc()
        d.dtype = _lib.DTYPE_CODES[plan.dtype]
        d.precision = settings["precision"]
        d.complex_mode = settings["complex_mode"]
        d.accumulate = int(bool(accumulate))
        d.m, d.n, d.k = plan.m, plan.n, plan.k
        d.a, d.a_ld, d.a_block_stride, d.a_major, d.a_nblocks = a_buf.ptr, plan.a.ld, plan.a.block_stride, plan.a.major, plan.a.nblocks
        d.b, d.b_ld, d.b_block_stride, d.b_major, d.b_nblocks = b_buf.ptr, plan.b.ld, plan.b.block_stride, plan.b.major, plan.b.nblocks
        d.c, d.c_ld, d.c_block_stride = out_storage.ptr + out_first_block * plan.out_block_elems * itemsize, plan.n, plan.out_block_elems
        d.c_nblocks = max(n_out_total, 1)
        d.n_out = plan.n_out if n_out is None else int(n_out)
        d.n_pairs = plan.n_pairs
        d.pair_a = ta.data_ptr() + 4 * pair_offset * plan.n_pairs
        d.pair_b = tb.data_ptr() + 4 * pair_offset * plan.n_pairs
        d.out_index = None if out_index is None else out_index.data_ptr()
        if peers is not None:  # fused exchange: reduce into the owners' peer-mapped buffers instead of storing
            arr = (ctypes.c_void_p * len(peers))(*[int(x) for x in peers])
            d.c_peers, d.n_peers, d.blocks_per_peer = arr, len(peers), int(blocks_per_peer)
        ws_bytes = _lib.contract_workspace_bytes(d)
        ws = None
        if ws_bytes:
            ws = DeviceStorage(ws_bytes, device)
            d.workspace, d.workspace_bytes = ws.ptr, ws_bytes
        if fuse:
            planes_a, planes_b = _lib.contract_planes(d) if ws is not None else (None, None)
            for which, op, bs, n_free, st, planes, flag in (("a", plan.a, a_blockshape, len(plan.free_a), a_storage, planes_a, _lib.RNB_FLAG_A_PRESPLIT),
                                                            ("b", plan.b, b_blockshape, len(plan.free_b), b_storage, planes_b, _lib.RNB_FLAG_B_PRESPLIT)):
                if op.direct:
                    continue
                done = False
                if planes is not None and planes.mode != _lib.RNB_SPLIT_NONE:
                    done = _lib.permute_split(_matricize_desc(op, bs, n_free, itemsize), st.ptr, planes, ws.ptr, op.rows, op.k, op.nblocks, stream)
                if done:
                    d.flags |= flag
                else:   # layout the fused kernel does not take: the two-pass route for this operand
                    buf = _matricize(op, bs, n_free, st, itemsize, stream)
                    if which == "a":
                        a_buf, d.a = buf, buf.ptr
                    else:
                        b_buf, d.b = buf, buf.ptr
        _lib.contract_grouped(d, stream)
        # keep temporaries alive until the stream has consumed them: torch's caching allocator
        # reuses a freed block only for work enqueued later on the same stream, which is ordered
        # after this launch, so dropping the Python references here is safe.
        del a_buf, b_buf, ws
    return out_storage


_nd_cache = {}


def _execute_nd(plan: ContractionPlan, a_storage, b_storage, out_storage, accumulate, out_index, n_out, pair_offset, out_first_block):
    """Small contraction whose operands are not in GEMM layout: one ``rnb_contract_nd`` launch on the blocks as they lie
    (the two matricize launches and the zero-padded workspaces of the general route are skipped)."""
    torch = _torch()
    device = a_storage.device
    itemsize = np.dtype(plan.dtype).itemsize
    with torch.cuda.device(device):
        stream = current_stream_ptr(device)
        if out_storage is None:
            out_storage = DeviceStorage(plan.n_out * plan.out_block_elems * itemsize, device)
        ta, tb, tabs = _device_tables(plan, device)
        key = id(plan)
        hit = _nd_cache.get(key)
        if hit is None or hit[0] is not plan:
            d = _lib.ContractNdDesc()
            ext_m, sa_m, ext_n, sb_n, ext_k, sa_k, sb_k = plan.nd
            d.dtype = _lib.DTYPE_CODES[plan.dtype]
            d.rank_m, d.rank_n, d.rank_k = len(ext_m), len(ext_n), len(ext_k)
            for i, (e, st) in enumerate(zip(ext_m, sa_m)):
                d.ext_m[i], d.a_stride_m[i] = e, st
            for i, (e, st) in enumerate(zip(ext_n, sb_n)):
                d.ext_n[i], d.b_stride_n[i] = e, st
            for i, (e, s1, s2) in enumerate(zip(ext_k, sa_k, sb_k)):
                d.ext_k[i], d.a_stride_k[i], d.b_stride_k[i] = e, s1, s2
            d.n_pairs = plan.n_pairs
            d.a_block_stride, d.b_block_stride = plan.a.block_stride if plan.a.direct else _block_elems(plan, "a"), \
                plan.b.block_stride if plan.b.direct else _block_elems(plan, "b")
            d.c_block_stride = plan.out_block_elems
            if len(_nd_cache) > 8192:
                _nd_cache.clear()
            hit = _nd_cache[key] = (plan, d)
        d = hit[1]
        d.accumulate = int(bool(accumulate))
        d.a, d.b = a_storage.ptr, b_storage.ptr
        d.c = out_storage.ptr + out_first_block * plan.out_block_elems * itemsize
        d.n_out = plan.n_out if n_out is None else int(n_out)
        d.pair_a = ta.data_ptr() + 4 * pair_offset * plan.n_pairs
        d.pair_b = tb.data_ptr() + 4 * pair_offset * plan.n_pairs
        d.out_index = None if out_index is None else out_index.data_ptr()
        if tabs is not None:      # offsets tabulated once per plan: the kernel does no index arithmetic
            m, n, k = plan.m, plan.n, prod(plan.nd[4])
            base = tabs.data_ptr()
            d.tab_m, d.tab_n, d.tab_ka, d.tab_kb = base, base + 4 * m, base + 4 * (m + n), base + 4 * (m + n + k)
        b = _batch
        if b is not None and tabs is not None and out_index is None and (b.stream is None or b.stream == stream):
            # a tensor-network driver is walking its path: queue the step; steps of one dependency level share a launch
            level = 1 + max(b.level_of.get(a_storage.ptr, 0), b.level_of.get(b_storage.ptr, 0), b.level_of.get(out_storage.ptr, 0) if accumulate else 0)
            b.tasks.append((level, _lib.ContractNdDesc.from_buffer_copy(d)))
            b.keep.append((a_storage, b_storage, out_storage, tabs, ta, tb))
            b.level_of[out_storage.ptr] = level
            b.stream, b.device = stream, device
            _lib._flush = flush_pending
            return out_storage
        _lib.contract_nd(d, stream)
    return out_storage


class _Batch:
    __slots__ = ("tasks", "keep", "level_of", "stream", "device")

    def __init__(self):
        self.tasks, self.keep, self.level_of, self.stream, self.device = [], [], {}, None, None


_batch = None


@contextlib.contextmanager
def batching():
    """Scope in which small contractions (``rnb_contract_nd`` route) are QUEUED instead of launched: the queue is flushed
    - one ``rnb_contract_nd_batch`` launch per dependency level - as soon as anything else is about to touch the stream
    (``engine.lib._pre``) and when the scope ends.  Operand and result storages of queued steps are kept alive until then."""
    global _batch
    if _batch is not None or os.environ.get("RNB_BATCH_SMALL") == "0":
        yield
        return
    _batch = _Batch()
    try:
        yield
    finally:
        try:
            flush_pending()
        finally:
            _batch = None


def flush_pending():
    b = _batch
    _lib._flush = None
    if b is None or not b.tasks:
        return
    tasks, keep = b.tasks, b.keep
    b.tasks, b.keep = [], []
    b.level_of.clear()
    torch = _torch()
    levels = {}
    for level, d in tasks:
        levels.setdefault(level, []).append(d)
    with torch.cuda.device(b.device):
        for level in sorted(levels):
            _lib.contract_nd_batch(levels[level], b.stream)
    del keep     # every queued launch is on the stream now: the allocator may reuse the temporaries for LATER work


def _block_elems(plan: ContractionPlan, which: str) -> int:
    """Elements of one STORED block of an operand (rows * true k: the matricize plan holds the padded k)."""
    ext_m, _sa, ext_n, _sb, ext_k, _s1, _s2 = plan.nd
    return (prod(ext_m) if which == "a" else prod(ext_n)) * prod(ext_k)


def blockify_strides(grid, blockshape):
    """Element strides of the dense C-order tensor and of block-major storage, both indexed by
    the interleaved axis list (g0, b0, g1, b1, ...)."""
    r = len(grid)
    shape = [g * b for g, b in zip(grid, blockshape)]
    dense_axis = [prod(shape[i + 1:]) for i in range(r)]
    blk = prod(blockshape)
    extents, dense, blocked = [], [], []
    for i in range(r):
        extents += [grid[i], blockshape[i]]
        dense += [dense_axis[i] * blockshape[i], dense_axis[i]]
        blocked += [prod(grid[i + 1:]) * blk, prod(blockshape[i + 1:])]
    return extents, dense, blocked


def dense_to_blocks(dense_ptr, storage: DeviceStorage, grid, blockshape, itemsize):
    """Device ingest: dense C-order tensor -> block-major storage (one permute launch).
    The step the commented-out ``array`` constructor did on the host
    (``rosnet/array/block/__init__.py:339-366``)."""
    extents, dense, blocked = blockify_strides(grid, blockshape)
    _lib.permute(dense_ptr, storage.ptr, itemsize, extents, dense, blocked, current_stream_ptr(storage.device))


def blocks_to_dense(storage: DeviceStorage, dense_ptr, grid, blockshape, itemsize):
    """``np.block`` on the device: block-major storage -> dense C-order tensor
    (``to_numpy``, ``rosnet/array/block/__init__.py:203-205``)."""
    extents, dense, blocked = blockify_strides(grid, blockshape)
    _lib.permute(storage.ptr, dense_ptr, itemsize, extents, blocked, dense, current_stream_ptr(storage.device))


def transpose_blocks(storage: DeviceStorage, grid, blockshape, axes, itemsize) -> DeviceStorage:
    """Permute every block by ``axes`` and the grid by ``axes`` in one launch
    (intent of ``rosnet/array/block/__init__.py:264-278``)."""
    r = len(grid)
    blk = prod(blockshape)
    new_grid = [grid[a] for a in axes]
    new_bs = [blockshape[a] for a in axes]
    out = DeviceStorage(storage.nbytes, storage.device)
    extents = list(grid) + list(blockshape)
    src = [prod(grid[i + 1:]) * blk for i in range(r)] + [prod(blockshape[i + 1:]) for i in range(r)]
    dst = [0] * (2 * r)
    for new_pos, old in enumerate(axes):
        dst[old] = prod(new_grid[new_pos + 1:]) * blk
        dst[r + old] = prod(new_bs[new_pos + 1:])
    _lib.permute(storage.ptr, out.ptr, itemsize, extents, src, dst, current_stream_ptr(storage.device))
    return out


def reshape_blocks_f(storage: DeviceStorage, nblocks, blockshape, new_blockshape, itemsize) -> DeviceStorage:
    """Fortran-order per-block reshape: ``block.reshape(new, order='F')`` equals
    ``block.T.reshape(new[::-1]).T`` in C order, i.e. two axis-reversal permutes."""
    r = len(blockshape)
    blk = prod(blockshape)
    stream = current_stream_ptr(storage.device)
    tmp = DeviceStorage(storage.nbytes, storage.device)
    # 1) reverse the axes of every block
    extents = [nblocks] + list(blockshape)
    src = [blk] + [prod(blockshape[i + 1:]) for i in range(r)]
    dst = [blk] + [prod(blockshape[:i]) for i in range(r)]
    _lib.permute(storage.ptr, tmp.ptr, itemsize, extents, src, dst, stream)
    # 2) reinterpret as reversed(new) in C order, reverse again
    out = DeviceStorage(storage.nbytes, storage.device)
    nr = len(new_blockshape)
    rev = list(new_blockshape)[::-1]
    extents = [nblocks] + rev
    src = [blk] + [prod(rev[i + 1:]) for i in range(nr)]
    dst = [blk] + [prod(rev[:i]) for i in range(nr)]
    _lib.permute(tmp.ptr, out.ptr, itemsize, extents, src, dst, stream)
    return out


def reblock(storage: DeviceStorage, grid, blockshape, new_blockshape, itemsize) -> DeviceStorage:
    """Change the block extents of a block-major array (same dense tensor, new grid).  One permute launch when every new
    block extent divides the old one (e.g. batch axes cut to block extent 1); otherwise through the dense layout."""
    grid, bs, nbs = [int(g) for g in grid], [int(b) for b in blockshape], [int(b) for b in new_blockshape]
    shape = [g * b for g, b in zip(grid, bs)]
    if any(s % nb for s, nb in zip(shape, nbs)):
        raise ValueError(f"blockshape {tuple(nbs)} does not tile shape {tuple(shape)}")
    ngrid = [s // nb for s, nb in zip(shape, nbs)]
    out = DeviceStorage(prod(shape) * itemsize, storage.device)
    stream = current_stream_ptr(storage.device)
    r = len(grid)
    if all(b % nb == 0 for b, nb in zip(bs, nbs)):
        blk, nblk = prod(bs), prod(nbs)
        extents, src, dst = [], [], []
        for i in range(r):
            q = bs[i] // nbs[i]
            gs, bst = prod(grid[i + 1:]) * blk, prod(bs[i + 1:])
            ngs, nbst = prod(ngrid[i + 1:]) * nblk, prod(nbs[i + 1:])
            extents += [grid[i], q, nbs[i]]
            src += [gs, nbs[i] * bst, bst]
            dst += [q * ngs, ngs, nbst]
        try:
            _lib.permute(storage.ptr, out.ptr, itemsize, extents, src, dst, stream)
            return out
        except _lib.RosnetB200Error:
            pass   # more than RNB_MAX_RANK axes after merging: take the two-step route
    dense = DeviceStorage(prod(shape) * itemsize, storage.device)
    blocks_to_dense(storage, dense.ptr, grid, bs, itemsize)
    dense_to_blocks(dense.ptr, out, ngrid, nbs, itemsize)
    return out


def diagonal_blocks(storage: DeviceStorage, grid, blockshape, p: int, q: int, itemsize) -> DeviceStorage:
    """``out[..., i, ...] = in[..., i, ..., i, ...]``: axes ``p < q`` (same grid and block extents) collapse onto axis
    ``p``, axis ``q`` disappears.  One strided-copy launch: the source strides of the two axes are added."""
    grid, bs = [int(g) for g in grid], [int(b) for b in blockshape]
    if not (p < q and grid[p] == grid[q] and bs[p] == bs[q]):
        raise ValueError("diagonal needs p < q with equal grid and block extents")
    r = len(grid)
    blk = prod(bs)
    keep = [i for i in range(r) if i != q]
    ngrid, nbs = [grid[i] for i in keep], [bs[i] for i in keep]
    nblk = prod(nbs)
    gs = [prod(grid[i + 1:]) * blk for i in range(r)]
    bst = [prod(bs[i + 1:]) for i in range(r)]
    extents = ngrid + nbs
    src = [gs[i] + (gs[q] if i == p else 0) for i in keep] + [bst[i] + (bst[q] if i == p else 0) for i in keep]
    dst = [prod(ngrid[j + 1:]) * nblk for j in range(len(keep))] + [prod(nbs[j + 1:]) for j in range(len(keep))]
    out = DeviceStorage(prod(ngrid) * nblk * itemsize, storage.device)
    _lib.permute(storage.ptr, out.ptr, itemsize, extents, src, dst, current_stream_ptr(storage.device))
    return out
