"""Block sizing for B200 (the chooser the reference's benchmarks expect but v0.3.0 never
defined: ``rn.tuning.configure(threshold_k=...)`` at ``benchmark/bench_contraction.py:44``;
SURVEY section 2.1 row 10).

The rule is hardware-derived rather than cluster-derived: a block pair should be a GEMM big
enough to fill whole 128x128 CTA tiles (FP64 DMMA) with k deep enough to amortise the pipeline
fill, small enough that a, b and the output of one contraction fit the HBM budget, and the
number of output blocks should cover the GPUs the scheduler deals them to.
"""
from __future__ import annotations

import contextlib
from math import prod

import numpy as np

from ..engine import lib as _lib

HBM_BYTES = 180 * 10**9          # B200
TILE_M = 128                      # CTA tile of csrc/gemm_f64.cu
MIN_K = 256                       # k per block below which pipeline fill / epilogue dominate

_state = {"threshold_k": None, "hbm_fraction": 0.8, "n_gpus": None}   # n_gpus None: every rank of the active scheduler context


@contextlib.contextmanager
def configure(threshold_k=None, hbm_fraction=None, n_gpus=None, complex_mode=None, precision=None):
    """Scoped tuning knobs.  ``threshold_k``: smallest contracted extent per block worth
    splitting across GPUs (below it the k index stays whole and only output blocks are dealt);
    ``n_gpus``: GPUs a contraction may be dealt to (default: every rank of ``engine.scheduler.init()``;
    1 keeps everything on the calling rank) — both are read by ``engine.scheduler.choose_tier``;
    ``complex_mode``: "auto" (default: complex128 3M from a contracted extent of 64 up, complex64 3M when the contracted extent is deep enough to repay its workspace pass, else 4M) | "4m" | "3m"; ``precision``: "default" (3xTF32) | "tf32x1" | "tf32_bf16x2" (TF32 hi.hi + two BF16 cross products, opt-in)."""
    from ..engine import runtime

    old, old_rt = dict(_state), dict(runtime.settings)
    if threshold_k is not None:
        _state["threshold_k"] = int(threshold_k)
    if hbm_fraction is not None:
        _state["hbm_fraction"] = float(hbm_fraction)
    if n_gpus is not None:
        _state["n_gpus"] = int(n_gpus)
    if complex_mode is not None:
        runtime.settings["complex_mode"] = {"4m": _lib.RNB_CPLX_4M, "3m": _lib.RNB_CPLX_3M, "auto": _lib.RNB_CPLX_AUTO}[complex_mode]
    if precision is not None:
        runtime.settings["precision"] = {"default": _lib.RNB_PREC_DEFAULT, "tf32x1": _lib.RNB_PREC_TF32X1, "tf32_bf16x2": _lib.RNB_PREC_TF32_BF16X2}[precision]
    try:
        yield dict(_state)
    finally:
        _state.clear()
        _state.update(old)
        runtime.settings.clear()
        runtime.settings.update(old_rt)


def threshold_k() -> int:
    return _state["threshold_k"] if _state["threshold_k"] is not None else MIN_K


def choose_blockshape(shape, dtype, n_gpus=None, target_block_bytes=None):
    """Pick a blockshape that tiles ``shape`` exactly.

    Greedy: start from the whole tensor as one block and halve the axis with the largest block
    extent (ties: leading axis first, which keeps trailing = contiguous axes long for coalesced
    permutes) until (i) a block is at most ``target_block_bytes`` (default: 1/64 of the HBM
    budget, capped at 1 GiB) and (ii) there are at least ``n_gpus`` blocks — never going below
    128 elements per axis that is split, so blocks still fill whole CTA tiles."""
    shape = tuple(int(s) for s in shape)
    itemsize = np.dtype(dtype).itemsize
    if n_gpus is None:
        from ..engine import scheduler

        n_gpus = scheduler.active_world()
    n_gpus = int(n_gpus)
    if target_block_bytes is None:
        target_block_bytes = min(int(HBM_BYTES * _state["hbm_fraction"]) // 64, 1 << 30)
    bs = list(shape)

    def splittable(i):
        return bs[i] % 2 == 0 and bs[i] // 2 >= min(TILE_M, shape[i])

    while prod(bs) * itemsize > target_block_bytes or prod(s // b for s, b in zip(shape, bs)) < n_gpus:
        cand = [i for i in range(len(bs)) if splittable(i)]
        if not cand:
            break
        i = max(cand, key=lambda j: (bs[j], -j))
        bs[i] //= 2
    return tuple(bs)
