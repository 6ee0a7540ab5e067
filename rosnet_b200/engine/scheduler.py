"""Single-box scheduler: how one blocked contraction (or a set of tensor-network slices) is dealt
to the GPUs of one node.  Replaces the pyCOMPSs task graph for this path
(``rosnet/tuning/task.py:15-81``; task-per-output-block / task-per-pair submission at
``rosnet/array/compss/__init__.py:430-451``).

One process per GPU (``torchrun``); ``torch.distributed`` is the plumbing.  Three partitions,
chosen by what is split:

1. **output blocks** — independent units, *no exchange*: rank r computes a contiguous range of the
   output-block list (rows of the plan's pair tables);
2. **contracted block index** — every rank holds a slice of the k-block list, produces a
   full-shape partial, and ONE reduce-scatter(sum) leaves each rank with a disjoint range of
   finished output blocks (block-major storage makes a range of blocks one contiguous chunk);
3. **slices** of a tensor network — independent contractions dealt round-robin, scalar/small
   all-reduce at the end.

The communicator is abstract so the host logic runs under ``gloo`` on CPU in the tests.

The scheduler sits BEHIND the array API: ``scheduler.init()`` (once per process, under torchrun)
makes ``np.tensordot`` on BlockArrays / ShardedBlockArrays (``array/sharded.py``) and
``tn.contract_network(..., comm=ctx)`` deal the work themselves — the user writes the same lines as
on one GPU, as with the reference where ``tensordot`` submits the COMPSs tasks itself.
``tuning.configure(n_gpus=, threshold_k=)`` steers the choice (``choose_tier``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def _flush_queued():
    """Small contractions queued by ``engine.runtime.batching`` must be on the stream before a collective reads their results."""
    from . import lib

    lib._pre()


def block_range(n_blocks: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of the contiguous range owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(n_blocks), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def padded_chunk(n_blocks: int, world: int) -> int:
    """Blocks per rank when every rank must receive the same count (NCCL reduce-scatter)."""
    return -(-int(n_blocks) // int(world))


def round_robin(n_items: int, world: int, rank: int) -> List[int]:
    """Items (tensor-network slices) dealt to ``rank``."""
    return list(range(rank, int(n_items), int(world)))


def split_contracted_axis(grid: Sequence[int], axis: int, world: int, rank: int) -> slice:
    """Slice of grid ``axis`` (a contracted block axis) held by ``rank``."""
    first, count = block_range(grid[axis], world, rank)
    return slice(first, first + count)


class TorchComm:
    """Collectives through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in CPU tests).
    Complex tensors are reduced through their real views (NCCL has no complex dtype)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    @staticmethod
    def _real(t):
        import torch

        return torch.view_as_real(t).reshape(-1) if t.is_complex() else t.reshape(-1)

    def reduce_scatter_sum(self, send, recv):
        """``send``: world * n elements, ``recv``: n elements; recv = sum over ranks of chunk[rank]."""
        _flush_queued()
        s, r = self._real(send), self._real(recv)
        backend = self.dist.get_backend(self.group)
        if backend == "gloo":  # gloo has no reduce_scatter: all-reduce then keep our chunk
            tmp = s.clone()
            self.dist.all_reduce(tmp, group=self.group)
            n = r.numel()
            r.copy_(tmp[self.rank * n:(self.rank + 1) * n])
        else:
            self.dist.reduce_scatter_tensor(r, s, group=self.group)
        return recv

    def all_reduce_sum(self, t):
        _flush_queued()
        self.dist.all_reduce(self._real(t), group=self.group)
        return t

    def all_gather(self, send, recv):
        _flush_queued()
        self.dist.all_gather_into_tensor(self._real(recv), self._real(send), group=self.group)
        return recv

    def barrier(self):
        _flush_queued()
        self.dist.barrier(group=self.group)


class NcclComm:
    """Own NCCL communicator driven through the C ABI (``rnb_nccl_*``, ``rnb_reduce_scatter_sum``):
    the unique id is created on rank 0 and broadcast with ``torch.distributed``."""

    def __init__(self, rank: int, world: int):
        import ctypes

        import torch
        import torch.distributed as dist

        from . import lib

        L = lib.load()
        self.rank, self.world = rank, world
        uid = (ctypes.c_char * lib.RNB_NCCL_UNIQUE_ID_BYTES)()
        if rank == 0:
            lib.check(L.rnb_nccl_get_unique_id(uid), "rnb_nccl_get_unique_id")
        payload = [bytes(uid.raw)]
        dist.broadcast_object_list(payload, src=0)
        uid = (ctypes.c_char * lib.RNB_NCCL_UNIQUE_ID_BYTES).from_buffer_copy(payload[0])
        self._comm = ctypes.c_void_p()
        lib.check(L.rnb_nccl_comm_init_rank(ctypes.byref(self._comm), world, uid, rank), "rnb_nccl_comm_init_rank")
        self._lib, self._L, self._torch = lib, L, torch

    def _code(self, t):
        return self._lib.DTYPE_CODES[str(t.dtype).replace("torch.", "")]

    def reduce_scatter_sum(self, send, recv):
        _flush_queued()
        stream = self._torch.cuda.current_stream().cuda_stream
        self._lib.check(self._L.rnb_reduce_scatter_sum(send.data_ptr(), recv.data_ptr(), recv.numel(), self._code(recv), self._comm, stream),
                        "rnb_reduce_scatter_sum")
        return recv

    def all_reduce_sum(self, t):
        _flush_queued()
        stream = self._torch.cuda.current_stream().cuda_stream
        self._lib.check(self._L.rnb_all_reduce_sum(t.data_ptr(), t.data_ptr(), t.numel(), self._code(t), self._comm, stream), "rnb_all_reduce_sum")
        return t

    def close(self):
        if self._comm:
            self._L.rnb_nccl_comm_destroy(self._comm)
            self._comm = None


@dataclass
class OwnedBlocks:
    """Result of a distributed contraction on one rank: a contiguous range of finished output
    blocks (block-major), plus the global layout they belong to."""

    out_grid: Tuple[int, ...]
    out_blockshape: Tuple[int, ...]
    dtype: np.dtype
    first: int
    count: int
    data: object  # flat buffer (torch tensor on the device, or ndarray in CPU tests) of count blocks

    def block_indices(self):
        idx = list(np.ndindex(*self.out_grid)) if len(self.out_grid) else [()]
        return idx[self.first:self.first + self.count]


def contract_output_partition(plan, local_exec: Callable, rank: int, world: int) -> OwnedBlocks:
    """Partition 1: this rank computes output blocks [first, first+count) — no communication.
    ``local_exec(pair_offset, n_out)`` runs the grouped contraction for that range of pair-table
    rows and returns the flat buffer of those blocks."""
    first, count = block_range(plan.n_out, world, rank)
    data = local_exec(first, count)
    return OwnedBlocks(plan.out_grid, plan.out_blockshape, np.dtype(plan.dtype), first, count, data)


def contract_k_partition(plan, local_partial, comm, alloc: Callable) -> OwnedBlocks:
    """Partition 2: ``local_partial`` is this rank's full-shape partial result (flat, block-major,
    padded to ``world * padded_chunk`` blocks); one reduce-scatter(sum) gives every rank its
    disjoint range of finished blocks.  ``alloc(n_elems)`` allocates the receive buffer."""
    chunk = padded_chunk(plan.n_out, comm.world)
    blk = plan.out_block_elems
    recv = alloc(chunk * blk)
    comm.reduce_scatter_sum(local_partial, recv)
    first = comm.rank * chunk
    count = max(0, min(chunk, plan.n_out - first))
    return OwnedBlocks(plan.out_grid, plan.out_blockshape, np.dtype(plan.dtype), first, count, recv)


class PeerBuffers:
    """One cudaMalloc'ed receive buffer per rank, exported to every other rank of the box with CUDA
    IPC (``rnb_ipc_*``): ``ptrs[r]`` is rank r's buffer as seen from THIS process (peer mapped over
    NVLink for r != rank).  Used by the fused GEMM + exchange (``rnb_contract_desc.c_peers``): every
    rank's epilogue adds its partial of output block g straight into its owner's buffer."""

    def __init__(self, nbytes: int, rank: int, world: int):
        import ctypes

        import torch.distributed as dist

        from . import lib

        L = lib.load()
        self._lib, self._L = lib, L
        self.rank, self.world, self.nbytes = rank, world, int(nbytes)
        own = ctypes.c_void_p()
        lib.check(L.rnb_device_malloc(self.nbytes, ctypes.byref(own)), "rnb_device_malloc")
        self.own = own.value
        handle = (ctypes.c_char * lib.RNB_IPC_HANDLE_BYTES)()
        lib.check(L.rnb_ipc_get_handle(own, handle), "rnb_ipc_get_handle")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        self.ptrs, self._opened = [], []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.own)
                continue
            p = ctypes.c_void_p()
            h = (ctypes.c_char * lib.RNB_IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
            lib.check(L.rnb_ipc_open_handle(h, ctypes.byref(p)), "rnb_ipc_open_handle")
            self.ptrs.append(p.value)
            self._opened.append(p.value)

    def zero(self, stream):
        _flush_queued()
        self._lib.check(self._L.rnb_memset_async(self.own, 0, self.nbytes, stream), "rnb_memset_async")

    def close(self):
        for p in self._opened:
            self._L.rnb_ipc_close_handle(p)
        self._opened = []
        if self.own:
            self._L.rnb_device_free(self.own)
            self.own = None


# ------------------------------------------------------------------------------------------
# the scheduler behind the API
# ------------------------------------------------------------------------------------------
def choose_tier(n_out: int, n_pairs: int, k: int, world: int, threshold_k: int) -> str:
    """Which partition a contraction of REPLICATED operands gets on ``world`` GPUs:

    * ``"output"`` — at least one output block per GPU: deal output blocks, no exchange (tier 1);
    * ``"ksplit"`` — fewer output blocks than GPUs, but the contracted block index has ``>= world``
      entries of extent ``>= threshold_k`` each: split the pair list, exchange partial sums (tier 2);
    * ``"local"``  — neither: every rank computes the whole (small) contraction, no exchange.
    """
    if world <= 1:
        return "local"
    if n_out >= world:
        return "output"
    if n_pairs >= world and k >= threshold_k:
        return "ksplit"
    return "local"


def ksplit_plan(plan, world: int, rank: int):
    """``plan`` restricted to this rank's share of the contracted block index: columns
    ``block_range(n_pairs, world, rank)`` of the pair tables (every output block keeps a partial sum)."""
    import dataclasses

    first, count = block_range(plan.n_pairs, world, rank)
    if count == 0:
        raise ValueError("rank %d of %d gets no contracted block (n_pairs=%d)" % (rank, world, plan.n_pairs))
    return dataclasses.replace(plan, n_pairs=count,
                               pair_a=np.ascontiguousarray(plan.pair_a[:, first:first + count]),
                               pair_b=np.ascontiguousarray(plan.pair_b[:, first:first + count]))


class Context:
    """The ranks of one box that take part in distributed contractions (one process per GPU).

    ``exchange``: how tier 2 sums the partials — ``"nccl"`` (reduce-scatter through the C ABI),
    ``"fused"`` (the FP64 GEMM epilogue adds into the owner's buffer over NVLink peer memory; float64 /
    complex128 only) or ``"auto"`` (fused where the kernel has it, NCCL otherwise)."""

    def __init__(self, rank: int, world: int, comm=None, exchange: str = "auto"):
        if exchange not in ("auto", "nccl", "fused"):
            raise ValueError("exchange must be 'auto', 'nccl' or 'fused'")
        self.rank, self.world, self.comm, self.exchange = int(rank), int(world), comm, exchange
        self._peers = None          # PeerBuffers, grown on demand
        self.stats = {"output": 0, "ksplit": 0, "local": 0, "fused": 0, "nccl": 0, "slices": 0}

    # ---- control plane (torch.distributed) ----
    def barrier(self):
        import torch
        import torch.distributed as dist

        _flush_queued()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier()

    def peer_buffers(self, nbytes: int) -> "PeerBuffers":
        """A receive buffer of at least ``nbytes`` on every rank, mapped into every rank (collective call)."""
        if self._peers is None or self._peers.nbytes < nbytes:
            if self._peers is not None:
                self.barrier()
                self._peers.close()
            self._peers = PeerBuffers(int(nbytes), self.rank, self.world)
        return self._peers

    def use_fused(self, dtype) -> bool:
        if self.exchange == "nccl":
            return False
        ok = str(np.dtype(dtype)) in ("float64", "complex128") and self.world <= 8
        if self.exchange == "fused" and not ok:
            raise ValueError("the fused GEMM + exchange epilogue exists for float64 / complex128 on up to 8 ranks")
        return ok

    def close(self):
        if self._peers is not None:
            self._peers.close()
            self._peers = None
        if self.comm is not None and hasattr(self.comm, "close"):
            self.comm.close()


_active: Optional[Context] = None


def init(exchange: str = "auto", comm=None) -> Context:
    """Join the ranks ``torch.distributed`` already connects (torchrun + ``init_process_group``).
    On GPUs the data path gets its own NCCL communicator through the C ABI (``NcclComm``); under gloo
    (CPU tests) collectives go through ``TorchComm``."""
    global _active
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("scheduler.init() needs an initialised torch.distributed process group (run under torchrun)")
    rank, world = dist.get_rank(), dist.get_world_size()
    if comm is None:
        comm = NcclComm(rank, world) if dist.get_backend() == "nccl" else TorchComm()
    if _active is not None:
        _active.close()
    _active = Context(rank, world, comm, exchange)
    return _active


def context() -> Optional[Context]:
    return _active


def shutdown():
    global _active
    if _active is not None:
        _active.close()
        _active = None


def active_world() -> int:
    """GPUs a contraction may be dealt to: the active context's ranks, capped by ``tuning.configure(n_gpus=)``."""
    if _active is None:
        return 1
    from ..tuning import blocking

    n = blocking._state.get("n_gpus")
    return _active.world if n is None else max(1, min(int(n), _active.world))
