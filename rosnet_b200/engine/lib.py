"""ctypes binding of ``librosnet_b200.so`` (the C ABI declared in ``include/rosnet_b200.h``).

There is no CPU fallback: if the shared library is missing or a CUDA call fails, the product
path raises ``RosnetB200Error``.  The library is built in-tree by ``rosnet_b200/csrc/build.py``
(``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "lib", "librosnet_b200.so")

RNB_MAX_RANK = 32
RNB_F32, RNB_F64, RNB_C64, RNB_C128 = 0, 1, 2, 3
RNB_MAJOR_K, RNB_MAJOR_MN = 0, 1
RNB_PREC_DEFAULT, RNB_PREC_TF32X1, RNB_PREC_TF32_BF16X2 = 0, 1, 2
RNB_CPLX_4M, RNB_CPLX_3M, RNB_CPLX_AUTO = 0, 1, 2
RNB_SPLIT_NONE, RNB_SPLIT_PLANES, RNB_SPLIT_C64_B4M, RNB_SPLIT_C64_3M = 0, 1, 2, 3
RNB_FLAG_A_PRESPLIT, RNB_FLAG_B_PRESPLIT = 1, 2
RNB_ERR_UNSUPPORTED = 3
RNB_NCCL_UNIQUE_ID_BYTES = 128

DTYPE_CODES = {"float32": RNB_F32, "float64": RNB_F64, "complex64": RNB_C64, "complex128": RNB_C128}

# every symbol include/rosnet_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "rnb_abi_version", "rnb_reload_env", "rnb_last_error_string", "rnb_get_device_props", "rnb_permute", "rnb_permute_split", "rnb_contract_planes",
    "rnb_contract_workspace_bytes", "rnb_contract_grouped", "rnb_contract_nd", "rnb_contract_nd_batch", "rnb_contract_route", "rnb_block_sum", "rnb_memcpy_h2d_async",
    "rnb_memcpy_d2h_async", "rnb_memcpy_d2d_async", "rnb_stream_synchronize", "rnb_nccl_get_unique_id", "rnb_nccl_comm_init_rank",
    "rnb_nccl_comm_destroy", "rnb_reduce_scatter_sum", "rnb_all_reduce_sum", "rnb_all_gather",
    "rnb_device_malloc", "rnb_device_free", "rnb_memset_async", "rnb_ipc_get_handle", "rnb_ipc_open_handle", "rnb_ipc_close_handle",
]
RNB_MAX_PEERS = 8
RNB_IPC_HANDLE_BYTES = 64


class RosnetB200Error(RuntimeError):
    pass


class PermuteDesc(Structure):
    _fields_ = [
        ("elem_bytes", c_int32),
        ("rank", c_int32),
        ("extent", c_int64 * RNB_MAX_RANK),
        ("src_stride", c_int64 * RNB_MAX_RANK),
        ("dst_stride", c_int64 * RNB_MAX_RANK),
        ("src", c_void_p),
        ("dst", c_void_p),
    ]


class ContractDesc(Structure):
    _fields_ = [
        ("dtype", c_int32), ("precision", c_int32), ("complex_mode", c_int32), ("accumulate", c_int32),
        ("m", c_int64), ("n", c_int64), ("k", c_int64),
        ("a", c_void_p), ("a_ld", c_int64), ("a_block_stride", c_int64), ("a_major", c_int32), ("a_nblocks", c_int32),
        ("b", c_void_p), ("b_ld", c_int64), ("b_block_stride", c_int64), ("b_major", c_int32), ("b_nblocks", c_int32),
        ("c", c_void_p), ("c_ld", c_int64), ("c_block_stride", c_int64), ("c_nblocks", c_int32),
        ("n_out", c_int32), ("n_pairs", c_int32), ("flags", c_int32),
        ("pair_a", c_void_p), ("pair_b", c_void_p), ("out_index", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("c_peers", POINTER(c_void_p)), ("n_peers", c_int32), ("blocks_per_peer", c_int32),
    ]


class OperandPlanes(Structure):
    _fields_ = [("mode", c_int32), ("reserved", c_int32), ("hi_offset", c_size_t), ("lo_offset", c_size_t)]


class ContractNdDesc(Structure):
    _fields_ = [
        ("dtype", c_int32), ("accumulate", c_int32), ("rank_m", c_int32), ("rank_n", c_int32), ("rank_k", c_int32),
        ("n_out", c_int32), ("n_pairs", c_int32), ("reserved", c_int32),
        ("ext_m", c_int64 * RNB_MAX_RANK), ("a_stride_m", c_int64 * RNB_MAX_RANK),
        ("ext_n", c_int64 * RNB_MAX_RANK), ("b_stride_n", c_int64 * RNB_MAX_RANK),
        ("ext_k", c_int64 * RNB_MAX_RANK), ("a_stride_k", c_int64 * RNB_MAX_RANK), ("b_stride_k", c_int64 * RNB_MAX_RANK),
        ("a", c_void_p), ("b", c_void_p), ("c", c_void_p),
        ("a_block_stride", c_int64), ("b_block_stride", c_int64), ("c_block_stride", c_int64),
        ("pair_a", c_void_p), ("pair_b", c_void_p), ("out_index", c_void_p),
        ("tab_m", c_void_p), ("tab_n", c_void_p), ("tab_ka", c_void_p), ("tab_kb", c_void_p),
    ]


class DeviceProps(Structure):
    _fields_ = [
        ("sm_count", c_int32), ("cc_major", c_int32), ("cc_minor", c_int32), ("max_smem_per_block", c_int32),
        ("total_mem_bytes", c_int64), ("l2_bytes_mb", c_int32), ("reserved", c_int32),
    ]


_lib = None


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RosnetB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(rosnet_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.rnb_abi_version.restype = c_int
    lib.rnb_last_error_string.restype = c_char_p
    lib.rnb_get_device_props.argtypes = [c_int, POINTER(DeviceProps)]
    lib.rnb_permute.argtypes = [POINTER(PermuteDesc), c_void_p]
    lib.rnb_contract_workspace_bytes.argtypes = [POINTER(ContractDesc), POINTER(c_size_t)]
    lib.rnb_contract_grouped.argtypes = [POINTER(ContractDesc), c_void_p]
    lib.rnb_contract_nd.argtypes = [POINTER(ContractNdDesc), c_void_p]
    lib.rnb_contract_nd_batch.argtypes = [POINTER(ContractNdDesc), c_int32, c_void_p]
    lib.rnb_contract_planes.argtypes = [POINTER(ContractDesc), POINTER(OperandPlanes), POINTER(OperandPlanes)]
    lib.rnb_permute_split.argtypes = [POINTER(PermuteDesc), POINTER(OperandPlanes), c_void_p, c_int64, c_int64, c_int32, c_void_p]
    lib.rnb_contract_route.argtypes = [POINTER(ContractDesc), POINTER(c_int32)]
    lib.rnb_block_sum.argtypes = [c_void_p, POINTER(c_void_p), c_int32, c_int64, c_int32, c_int32, c_void_p]
    lib.rnb_memcpy_h2d_async.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.rnb_memcpy_d2h_async.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.rnb_memcpy_d2d_async.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
    lib.rnb_stream_synchronize.argtypes = [c_void_p]
    lib.rnb_nccl_get_unique_id.argtypes = [ctypes.c_char * RNB_NCCL_UNIQUE_ID_BYTES]
    lib.rnb_nccl_comm_init_rank.argtypes = [POINTER(c_void_p), c_int, ctypes.c_char * RNB_NCCL_UNIQUE_ID_BYTES, c_int]
    lib.rnb_nccl_comm_destroy.argtypes = [c_void_p]
    lib.rnb_reduce_scatter_sum.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.rnb_all_reduce_sum.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.rnb_all_gather.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.rnb_device_malloc.argtypes = [c_size_t, POINTER(c_void_p)]
    lib.rnb_device_free.argtypes = [c_void_p]
    lib.rnb_memset_async.argtypes = [c_void_p, c_int, c_size_t, c_void_p]
    lib.rnb_ipc_get_handle.argtypes = [c_void_p, ctypes.c_char * RNB_IPC_HANDLE_BYTES]
    lib.rnb_ipc_open_handle.argtypes = [ctypes.c_char * RNB_IPC_HANDLE_BYTES, POINTER(c_void_p)]
    lib.rnb_ipc_close_handle.argtypes = [c_void_p]
    for name in EXPORTED_SYMBOLS:
        if name not in ("rnb_last_error_string",):
            getattr(lib, name).restype = c_int
    lib.rnb_last_error_string.restype = c_char_p
    if lib.rnb_abi_version() != 1:
        raise RosnetB200Error("librosnet_b200.so ABI version mismatch")
    _lib = lib
    return lib


def reload_env():
    """Make the library re-read its RNB_* environment knobs (they are cached at first use)."""
    load().rnb_reload_env()


def check(status: int, what: str):
    if status != 0:
        msg = load().rnb_last_error_string().decode("utf-8", "replace")
        raise RosnetB200Error(f"{what} failed with status {status}: {msg}")


# engine.runtime queues small contractions while a tensor-network driver walks its path (one launch per dependency
# level instead of one per step); anything else that touches the stream flushes the queue first through this hook
_flush = None


def _pre():
    if _flush is not None:
        _flush()


# counts launches of OUR kernels (bench.py reports it as gpu_launches)
launch_counter = {"permute": 0, "contract": 0, "block_sum": 0}


_perm_cache = {}


def permute_desc(elem_bytes, extents, src_strides, dst_strides) -> PermuteDesc:
    """Descriptor of a strided n-d copy (pointers unset), cached by shape: tensor-network paths issue
    the same few hundred layouts for every slice, and filling a 3 x 32-entry ctypes struct from
    Python costs more host time than the small kernels take."""
    key = (int(elem_bytes), tuple(extents), tuple(src_strides), tuple(dst_strides))
    d = _perm_cache.get(key)
    if d is not None:
        return d
    d = PermuteDesc()
    d.elem_bytes = int(elem_bytes)
    # drop extent-1 axes and merge axes adjacent in both layouts (the library does it too; doing it
    # here keeps tensor-network operands with 2 x 32 grid+block axes inside RNB_MAX_RANK)
    dims = sorted(((int(s), int(e), int(t)) for e, s, t in zip(extents, src_strides, dst_strides) if int(e) != 1), reverse=True)
    merged = []
    for s_, e_, t_ in dims:
        if merged and merged[-1][0] == s_ * e_ and merged[-1][2] == t_ * e_:
            ps, pe, pt = merged[-1]
            merged[-1] = (s_, pe * e_, t_)
        else:
            merged.append((s_, e_, t_))
    ext = [m[1] for m in merged]
    sst = [m[0] for m in merged]
    dst = [m[2] for m in merged]
    rank = len(ext)
    if rank == 0:
        ext, sst, dst, rank = [1], [1], [1], 1
    if rank > RNB_MAX_RANK:
        raise RosnetB200Error(f"permute rank {rank} exceeds RNB_MAX_RANK={RNB_MAX_RANK}")
    d.rank = rank
    for i in range(rank):
        d.extent[i] = int(ext[i])
        d.src_stride[i] = int(sst[i])
        d.dst_stride[i] = int(dst[i])
    if len(_perm_cache) > 4096:
        _perm_cache.clear()
    _perm_cache[key] = d
    return d


def permute_launch(d: PermuteDesc, src_ptr, dst_ptr, stream):
    _pre()
    d.src = src_ptr
    d.dst = dst_ptr
    check(load().rnb_permute(ctypes.byref(d), c_void_p(stream)), "rnb_permute")
    launch_counter["permute"] += 1


def contract_planes(desc: ContractDesc):
    """(planes of a, planes of b) of this contraction's workspace: where ``permute_split`` has to write and in which mode."""
    pa, pb = OperandPlanes(), OperandPlanes()
    check(load().rnb_contract_planes(ctypes.byref(desc), ctypes.byref(pa), ctypes.byref(pb)), "rnb_contract_planes")
    return pa, pb


def permute_split(d: PermuteDesc, src_ptr, planes: OperandPlanes, workspace_ptr, rows, k, nblocks, stream) -> bool:
    """Matricize fused with the TF32 split: False (nothing launched) when the fused kernel cannot take the layout."""
    _pre()
    d.src = src_ptr
    d.dst = workspace_ptr
    status = load().rnb_permute_split(ctypes.byref(d), ctypes.byref(planes), c_void_p(workspace_ptr), int(rows), int(k), int(nblocks), c_void_p(stream))
    if status == RNB_ERR_UNSUPPORTED:
        return False
    check(status, "rnb_permute_split")
    launch_counter["permute"] += 1
    return True


def permute(src_ptr, dst_ptr, elem_bytes, extents, src_strides, dst_strides, stream):
    """dst[sum idx*dst_stride] = src[sum idx*src_stride]; strides in elements."""
    permute_launch(permute_desc(elem_bytes, extents, src_strides, dst_strides), src_ptr, dst_ptr, stream)


def contract_grouped(desc: ContractDesc, stream):
    _pre()
    check(load().rnb_contract_grouped(ctypes.byref(desc), c_void_p(stream)), "rnb_contract_grouped")
    launch_counter["contract"] += 1


def contract_nd(desc: ContractNdDesc, stream):
    """Small contraction on operands left in their own n-d layout (no matricize launches)."""
    _pre()
    check(load().rnb_contract_nd(ctypes.byref(desc), c_void_p(stream)), "rnb_contract_nd")
    launch_counter["contract"] += 1


def contract_nd_batch(descs, stream):
    """``descs``: a list of ContractNdDesc of INDEPENDENT small contractions (offset tables set): one launch per 340 tasks."""
    n = len(descs)
    arr = (ContractNdDesc * n)(*descs)
    check(load().rnb_contract_nd_batch(arr, n, c_void_p(stream)), "rnb_contract_nd_batch")
    launch_counter["contract"] += -(-n // 340)


ROUTE_TENSOR, ROUTE_SKINNY, ROUTE_SMALL = 0, 1, 2


def contract_route(desc: ContractDesc) -> int:
    """Kernel family ``rnb_contract_grouped`` picks for this descriptor (ROUTE_*)."""
    out = c_int32(0)
    check(load().rnb_contract_route(ctypes.byref(desc), ctypes.byref(out)), "rnb_contract_route")
    return int(out.value)


def contract_workspace_bytes(desc: ContractDesc) -> int:
    out = c_size_t(0)
    check(load().rnb_contract_workspace_bytes(ctypes.byref(desc), ctypes.byref(out)), "rnb_contract_workspace_bytes")
    return int(out.value)


def block_sum(dst_ptr, src_ptrs, count, dtype_code, accumulate, stream):
    _pre()
    arr = (c_void_p * len(src_ptrs))(*src_ptrs)
    check(load().rnb_block_sum(c_void_p(dst_ptr), arr, len(src_ptrs), int(count), int(dtype_code), int(bool(accumulate)), c_void_p(stream)), "rnb_block_sum")
    launch_counter["block_sum"] += 1


def memcpy_h2d(dst_ptr, src_ptr, nbytes, stream):
    _pre()
    check(load().rnb_memcpy_h2d_async(c_void_p(dst_ptr), c_void_p(src_ptr), int(nbytes), c_void_p(stream)), "rnb_memcpy_h2d_async")


def memcpy_d2h(dst_ptr, src_ptr, nbytes, stream):
    _pre()
    check(load().rnb_memcpy_d2h_async(c_void_p(dst_ptr), c_void_p(src_ptr), int(nbytes), c_void_p(stream)), "rnb_memcpy_d2h_async")


def memcpy_d2d(dst_ptr, src_ptr, nbytes, stream):
    _pre()
    check(load().rnb_memcpy_d2d_async(c_void_p(dst_ptr), c_void_p(src_ptr), int(nbytes), c_void_p(stream)), "rnb_memcpy_d2d_async")


def stream_synchronize(stream):
    _pre()
    check(load().rnb_stream_synchronize(c_void_p(stream)), "rnb_stream_synchronize")


def device_props(device: int = 0) -> DeviceProps:
    p = DeviceProps()
    check(load().rnb_get_device_props(int(device), ctypes.byref(p)), "rnb_get_device_props")
    return p
