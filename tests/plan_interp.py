"""NumPy interpreter of a ContractionPlan — TEST INFRASTRUCTURE.  It performs, on the host, exactly
the data movement and arithmetic the plan asks the CUDA library for (matricize per OperandPlan,
pair tables, majors, leading dimensions, zero padding), so the host planning logic can be checked
against dense numpy.tensordot on a CPU-only box.  The product never imports it."""
from math import prod

import numpy as np

from rosnet_b200.engine import lib
from rosnet_b200.engine.plan import matricize_desc


def strided_copy(src_flat, dst_flat, extents, src_strides, dst_strides):
    """What rnb_permute does: dst[sum idx*dst_stride] = src[sum idx*src_stride]."""
    idx = np.indices(extents).reshape(len(extents), -1)
    so = (idx * np.array(src_strides)[:, None]).sum(0)
    do = (idx * np.array(dst_strides)[:, None]).sum(0)
    dst_flat[do] = src_flat[so]


def operand_matrices(op, blockshape, n_free, blocks_flat):
    """Return a list (per block) of 2-D views [rows][k] regardless of storage order."""
    blk = prod(blockshape)
    out = []
    if op.direct:
        for j in range(op.nblocks):
            raw = blocks_flat[j * op.block_stride: j * op.block_stride + blk]
            if op.major == lib.RNB_MAJOR_K:
                out.append(raw.reshape(op.rows, op.ld)[:, : op.k])
            else:
                out.append(raw.reshape(op.k, op.ld)[:, : op.rows].T)
        return out
    ws = np.zeros(op.ws_elems, dtype=blocks_flat.dtype)
    extents, src, dst = matricize_desc(blockshape, n_free, op.perm, op.ld, op.block_stride, op.nblocks)
    strided_copy(blocks_flat, ws, extents, src, dst)
    for j in range(op.nblocks):
        out.append(ws[j * op.block_stride: j * op.block_stride + op.rows * op.ld].reshape(op.rows, op.ld)[:, : op.k])
    return out


def run_plan(plan, a_blocks_flat, a_blockshape, b_blocks_flat, b_blockshape):
    """Block-major output storage (flat) computed the way the grouped kernel is asked to."""
    A = operand_matrices(plan.a, a_blockshape, len(plan.free_a), a_blocks_flat)
    B = operand_matrices(plan.b, b_blockshape, len(plan.free_b), b_blocks_flat)
    out = np.zeros(plan.n_out * plan.m * plan.n, dtype=np.result_type(a_blocks_flat, b_blocks_flat))
    for g in range(plan.n_out):
        acc = np.zeros((plan.m, plan.n), dtype=out.dtype)
        for p in range(plan.n_pairs):
            acc += A[plan.pair_a[g, p]] @ B[plan.pair_b[g, p]].T
        out[g * plan.m * plan.n: (g + 1) * plan.m * plan.n] = acc.reshape(-1)
    return out


def block_major(dense, blockshape):
    grid = tuple(s // b for s, b in zip(dense.shape, blockshape))
    parts = []
    for bidx in np.ndindex(*grid):
        sl = tuple(slice(i * b, (i + 1) * b) for i, b in zip(bidx, blockshape))
        parts.append(np.ascontiguousarray(dense[sl]).reshape(-1))
    return np.concatenate(parts) if parts else np.zeros(0, dense.dtype), grid


def from_block_major(flat, grid, blockshape):
    blk = prod(blockshape)
    shape = tuple(g * b for g, b in zip(grid, blockshape))
    out = np.empty(shape, dtype=flat.dtype)
    for i, bidx in enumerate(np.ndindex(*grid)):
        sl = tuple(slice(j * b, (j + 1) * b) for j, b in zip(bidx, blockshape))
        out[sl] = flat[i * blk: (i + 1) * blk].reshape(blockshape)
    return out
