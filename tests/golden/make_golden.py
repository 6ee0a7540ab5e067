#!/usr/bin/env python
"""Generate tests/golden/blocked_tensordot.npz by running the UNMODIFIED reference.

Run in the build container only (it reads /root/reference, which does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.

The reference cannot be imported as shipped (``multimethod`` and ``autoray`` are not
installed and there is no network), so two throw-away stand-ins under ``tests/golden/_standins``
are put on ``sys.path`` first.  They only provide type-dispatch and ``autoray.do`` name lookup;
all arithmetic executed is the reference's own code calling NumPy:
``rosnet/array/block/__init__.py:281-336``.

Per case the file stores the inputs, the dense output of the reference
(``np.asarray(np.tensordot(BlockArray, BlockArray, axes))``), its grid / blockshape / shape /
dtype, and the dense ``numpy.tensordot`` of the assembled operands.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "_standins"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, HERE)

import numpy as np  # noqa: E402

import rosnet  # noqa: E402  (the reference)
from rosnet import BlockArray  # noqa: E402

from cases import CASES, case_inputs  # noqa: E402

assert rosnet.__file__.startswith("/root/reference"), rosnet.__file__


def to_blockarray(dense, blockshape):
    grid = tuple(s // bs for s, bs in zip(dense.shape, blockshape))
    blocks = []
    for bidx in np.ndindex(*grid):
        sl = tuple(slice(i * bs, (i + 1) * bs) for i, bs in zip(bidx, blockshape))
        blocks.append(np.ascontiguousarray(dense[sl]))
    return BlockArray(blocks, grid=grid)


def main():
    out = {}
    for i, case in enumerate(CASES):
        a, b = case_inputs(case, i)
        A = to_blockarray(a, case["ab"])
        B = to_blockarray(b, case["bb"])
        axes = [tuple(case["axes"][0]), tuple(case["axes"][1])]
        C = np.tensordot(A, B, axes)  # -> reference __array_function__ -> blocked tensordot
        assert isinstance(C, BlockArray)
        dense_ref = np.asarray(C)
        while dense_ref.dtype == object:  # full contraction: 0-d object grid holding the scalar
            dense_ref = np.asarray(dense_ref.item())
        dense_np = np.tensordot(a, b, axes)
        n = case["name"]
        out[f"{n}/a"] = a
        out[f"{n}/b"] = b
        out[f"{n}/ref"] = dense_ref
        out[f"{n}/dense"] = dense_np
        out[f"{n}/grid"] = np.array(C.grid, dtype=np.int64)
        out[f"{n}/blockshape"] = np.array(C.blockshape, dtype=np.int64)
        out[f"{n}/shape"] = np.array([int(s) for s in C.shape], dtype=np.int64)
        out[f"{n}/dtype"] = np.array(str(C.dtype))
        den = np.linalg.norm(dense_np.ravel()) or 1.0
        err = np.linalg.norm((dense_ref - dense_np).ravel()) / den
        print(f"{n:28s} grid={C.grid} blockshape={C.blockshape} dtype={C.dtype} ref-vs-dense={err:.3e}")
    # reference cost models (rosnet/tuning/flops.py:8-13) through the same by-name lookup the
    # reference test uses (test/tuning/test_flops.py:8-16)
    from autoray import do

    class Shape:
        def __init__(self, shape):
            self.shape = shape

    fl = []
    for m in (1, 2, 21, 32768):
        for n_ in (1, 2, 21, 32768):
            for k in (1, 2, 21, 32768):
                fl.append((m, n_, k, do("tensordot", Shape((m, k)), Shape((k, n_)), [(1,), (0,)], like="rosnet.tuning.flops")))
    out["flops_matmul"] = np.array(fl, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "blocked_tensordot.npz"), **out)
    print("wrote", os.path.join(HERE, "blocked_tensordot.npz"))


if __name__ == "__main__":
    main()
