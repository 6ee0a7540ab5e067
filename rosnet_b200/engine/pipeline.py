"""Pipelined contraction for HOST-resident operands.

When both operands are BlockArrays of host ndarray blocks (what the reference holds) the naive
sequence is H2D(a) -> H2D(b) -> launch -> D2H(c), fully serialised on PCIe.  Output blocks only
need the a-blocks of their grid row and the b-blocks of their grid column, so this module streams
the blocks in the order the output needs them and overlaps three things on three CUDA streams:

    copy-in stream : a-row 0, then b column by column, then the remaining a-rows
    compute stream : one grouped launch per chunk of output blocks, as soon as its inputs landed
    copy-out stream: finished output blocks -> pinned host mirror (PCIe is full duplex)

The result is a device-resident BlockArray that also carries the host mirror, so ``to_host()``
only waits for the copy-out stream.  This is the analogue of COMPSs moving task inputs/outputs in
the background while other tasks run (``rosnet/array/compss/__init__.py:430-451``).
Only used when both operands are consumable in place (no matricize pass) and there is more than
one output block; otherwise the simple path in ``array/block.py`` runs.
"""
from __future__ import annotations

import numpy as np

from ..array.device import DeviceStorage, _torch, pinned_empty
from . import lib as _lib
from . import runtime

_streams = {}
MIN_BYTES = 8 << 20


def _side_streams(device):
    torch = _torch()
    key = device.index
    if key not in _streams:
        _streams[key] = tuple(torch.cuda.Stream(device=device) for _ in range(3))
    return _streams[key]


def _tiles_per_block(plan) -> int:
    """CTA tiles one output block gives the grouped GEMM (tile sizes of csrc/gemm_*.cu)."""
    if plan.dtype in ("float32", "complex64"):
        f = 2 if plan.dtype == "complex64" else 1
        return -(-plan.m // 128) * -(-(plan.n * f) // 256)
    bn = 128 if plan.dtype == "float64" else 64
    return -(-plan.m // 128) * -(-plan.n // bn)


def eligible(plan, a, b) -> bool:
    if a.is_device or b.is_device or not (plan.a.direct and plan.b.direct):
        return False
    if plan.dtype in ("float32", "complex64"):
        # the split pre-pass of csrc/gemm_tf32.cu converts ALL operand blocks on every call: per chunk it would repeat that
        # traffic and read blocks whose host-to-device copy has not landed yet
        return False
    if plan.n_out < 2:
        return False
    return (a.nbytes + b.nbytes) >= MIN_BYTES


def _host_block_ptrs(arr, dtype):
    """(list of host pointers, keepalive list) of C-contiguous blocks in row-major grid order."""
    ptrs, keep = [], []
    for blk in arr.data.flat:
        blk = np.asarray(blk)
        if blk.dtype != dtype or not blk.flags.c_contiguous:
            blk = np.ascontiguousarray(blk, dtype=dtype)
        keep.append(blk)
        ptrs.append(blk.ctypes.data)
    return ptrs, keep


def _copy_blocks(ids, ptrs, dev_base, nbytes, stream):
    """H2D a set of blocks, merging runs that are contiguous on both sides into one copy."""
    ids = sorted(ids)
    i = 0
    while i < len(ids):
        j = i
        while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1 and ptrs[ids[j + 1]] == ptrs[ids[j]] + nbytes:
            j += 1
        _lib.memcpy_h2d(dev_base + ids[i] * nbytes, ptrs[ids[i]], (j - i + 1) * nbytes, stream)
        i = j + 1


def pipelined_tensordot(plan, a, b, dtype):
    """Returns (out_storage, host_mirror, done_event).  ``host_mirror`` is a pinned ndarray of shape
    (n_out,) + out_blockshape filled asynchronously; ``done_event`` fires when it is complete."""
    torch = _torch()
    dtype = np.dtype(dtype)
    device = torch.device("cuda", torch.cuda.current_device())
    s_in, s_out, s_c1 = _side_streams(device)
    compute = torch.cuda.current_stream(device)
    a_nb = a.blocksize * dtype.itemsize
    b_nb = b.blocksize * dtype.itemsize
    o_nb = plan.out_block_elems * dtype.itemsize
    sa = DeviceStorage(a_nb * a.nblock, device)
    sb = DeviceStorage(b_nb * b.nblock, device)
    out = DeviceStorage(o_nb * plan.n_out, device)
    mirror = pinned_empty((plan.n_out,) + tuple(plan.out_blockshape), dtype)
    a_ptrs, keep_a = _host_block_ptrs(a, dtype)
    b_ptrs, keep_b = _host_block_ptrs(b, dtype)

    # Chunks of consecutive output blocks: the first output row one block at a time (compute starts
    # after one a-row and one b-column have landed), the remaining rows whole.  Measured on C1
    # (B200, PCIe-bound): 7 chunks 8.2 ms/step; 16 single-block chunks on two compute lanes 10.6 ms
    # (per-chunk host cost and half-empty launches outweigh the shorter tail); serial 11.3 ms.
    n_ob = int(np.prod([b.grid[ax] for ax in plan.free_b])) if plan.free_b else 1
    n_oa = plan.n_out // n_ob
    chunks = [(j, 1) for j in range(n_ob)] + [(r * n_ob, n_ob) for r in range(1, n_oa)]
    lanes = (compute,)
    have_a, have_b = set(), set()
    # the side streams must not run ahead of work already queued on the compute stream that
    # still uses recycled memory
    for st in (s_in, s_out, s_c1):
        st.wait_stream(compute)
    for n, (first, count) in enumerate(chunks):
        lane = lanes[n % len(lanes)]
        need_a = set(int(x) for x in plan.pair_a[first:first + count].reshape(-1)) - have_a
        need_b = set(int(x) for x in plan.pair_b[first:first + count].reshape(-1)) - have_b
        if need_a or need_b:
            _copy_blocks(need_a, a_ptrs, sa.ptr, a_nb, s_in.cuda_stream)
            _copy_blocks(need_b, b_ptrs, sb.ptr, b_nb, s_in.cuda_stream)
            have_a |= need_a
            have_b |= need_b
            ev = torch.cuda.Event()
            ev.record(s_in)
            for ln in lanes:
                ln.wait_event(ev)
        with torch.cuda.stream(lane):
            runtime.execute_tensordot(plan, sa, a.blockshape, sb, b.blockshape, out_storage=out, n_out=count, pair_offset=first,
                                      out_first_block=first)
        done = torch.cuda.Event()
        done.record(lane)
        s_out.wait_event(done)
        _lib.memcpy_d2h(mirror.ctypes.data + first * o_nb, out.ptr + first * o_nb, count * o_nb, s_out.cuda_stream)
    if len(lanes) > 1:
        compute.wait_stream(s_c1)  # later work on the caller's stream sees the whole result
    finished = torch.cuda.Event()
    finished.record(s_out)
    # storages are used on side streams too: tell the caching allocator
    for st in (sa, sb):
        st.tensor.record_stream(s_in)
        st.tensor.record_stream(s_c1)
    out.tensor.record_stream(s_out)
    out.tensor.record_stream(s_c1)
    return out, mirror, finished, (keep_a, keep_b)
