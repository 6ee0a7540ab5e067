"""Deterministic contraction-path search and slicing for tensor networks.

Pure integer bookkeeping on index labels and extents (no tensor data).  The reference delegates
this to ``opt_einsum.contract_path(..., optimize="greedy")`` and ``cotengra.SliceFinder``
(``benchmark/bench_random_tn.py:80-110``); neither is installed here, so the build carries its
own: paths must be reproducible (seeded tie-breaks) so that the CPU oracle and the GPU contract
exactly the same tree.
"""
from __future__ import annotations

import heapq
from math import prod
from typing import Dict, List, Sequence, Tuple

import numpy as np


def _size(ix, sizes):
    return prod(sizes[i] for i in ix)


def greedy_path(inputs: Sequence[Sequence[int]], output: Sequence[int], sizes: Dict[int, int], seed: int = 0,
                alpha: float = 1.0, temperature: float = 0.0) -> List[Tuple[int, int]]:
    """SSA-style greedy path: repeatedly contract the connected pair minimising
    ``size(out) - alpha * (size(a) + size(b))`` (opt_einsum's greedy cost), tie-broken by a seeded
    random key; disconnected components are joined at the end by outer products, smallest first.
    Returns a list of (i, j) SSA ids: inputs are 0..n-1, the k-th contraction creates id n+k."""
    rng = np.random.default_rng(seed)
    n = len(inputs)
    tensors: Dict[int, frozenset] = {i: frozenset(ix) for i, ix in enumerate(inputs)}
    out_set = frozenset(output)
    owners: Dict[int, set] = {}
    for t, ix in tensors.items():
        for i in ix:
            owners.setdefault(i, set()).add(t)
    # how many live tensors (plus the output) still carry each index
    count = {i: len(o) + (1 if i in out_set else 0) for i, o in owners.items()}

    def result_of(a, b):
        ia, ib = tensors[a], tensors[b]
        shared = ia & ib
        drop = {i for i in shared if count[i] == 2 and i not in out_set}
        return (ia | ib) - drop

    def cost(a, b):
        res = result_of(a, b)
        c = _size(res, sizes) - alpha * (_size(tensors[a], sizes) + _size(tensors[b], sizes))
        if temperature > 0:
            c -= temperature * abs(c) * float(rng.gumbel())
        return c

    heap = []
    seen_pairs = set()

    def push_pairs(t):
        for i in tensors[t]:
            for u in owners[i]:
                if u != t and u in tensors:
                    key = (min(t, u), max(t, u))
                    if key not in seen_pairs:
                        seen_pairs.add(key)
                        heapq.heappush(heap, (cost(*key), float(rng.random()), key))

    for t in list(tensors):
        push_pairs(t)
    path = []
    next_id = n
    while heap:
        c, _, (a, b) = heapq.heappop(heap)
        if a not in tensors or b not in tensors:
            continue
        res = frozenset(result_of(a, b))
        for i in tensors[a] | tensors[b]:
            owners[i].discard(a)
            owners[i].discard(b)
        for i in (tensors[a] & tensors[b]):
            count[i] -= 1  # two carriers become one (or zero if dropped)
            if i not in res:
                count[i] -= 1
        del tensors[a], tensors[b]
        tensors[next_id] = res
        for i in res:
            owners[i].add(next_id)
        path.append((a, b))
        push_pairs(next_id)
        next_id += 1
    # outer products of what is left (disconnected components), smallest first
    rest = sorted(tensors, key=lambda t: (_size(tensors[t], sizes), t))
    while len(rest) > 1:
        a, b = rest[0], rest[1]
        tensors[next_id] = tensors[a] | tensors[b]
        path.append((a, b))
        rest = sorted([t for t in rest[2:]] + [next_id], key=lambda t: (_size(tensors[t], sizes), t))
        next_id += 1
    return path


def path_cost(inputs, output, sizes, path, sliced: Sequence[int] = ()):
    """(total multiply-adds, largest intermediate in elements, per-step list of (m, n, k)) of one
    slice of the contraction (sliced indices have extent 1)."""
    sizes = dict(sizes)
    for i in sliced:
        sizes[i] = 1
    tensors = {i: tuple(ix) for i, ix in enumerate(inputs)}
    out_set = set(output)
    count: Dict[int, int] = {}
    for ix in tensors.values():
        for i in ix:
            count[i] = count.get(i, 0) + 1
    for i in out_set:
        count[i] = count.get(i, 0) + 1
    next_id = len(inputs)
    total, largest, steps = 0, max((_size(ix, sizes) for ix in tensors.values()), default=1), []
    for a, b in path:
        ia, ib = tensors.pop(a), tensors.pop(b)
        shared = [i for i in ia if i in ib]
        contracted = [i for i in shared if count[i] == 2]
        keep_a = [i for i in ia if i not in contracted]
        keep_b = [i for i in ib if i not in contracted and i not in ia]
        for i in shared:
            count[i] -= 1
        for i in contracted:
            count[i] -= 1
        res = tuple(keep_a + keep_b)
        m = _size([i for i in ia if i not in contracted], sizes)
        nn = _size([i for i in ib if i not in contracted], sizes)
        k = _size(contracted, sizes)
        total += m * nn * k
        steps.append((m, nn, k))
        largest = max(largest, _size(res, sizes))
        tensors[next_id] = res
        next_id += 1
    return total, largest, steps


def find_slices(inputs, output, sizes, path, target_size: int, max_slices: int = 1 << 40) -> List[int]:
    """Greedy slice finder: while the largest intermediate exceeds ``target_size`` elements, slice
    the (non-output) index that shrinks the largest intermediate most and, on ties, adds the
    least total work.  Returns the sliced index labels (cotengra.SliceFinder's role)."""
    sliced: List[int] = []
    out_set = set(output)
    n_slices = 1
    while True:
        total, largest, _ = path_cost(inputs, output, sizes, path, sliced)
        if largest <= target_size or n_slices >= max_slices:
            return sliced
        # indices on the largest intermediates are the only useful candidates
        cand = _indices_on_large_intermediates(inputs, output, sizes, path, sliced, largest)
        cand = [i for i in cand if i not in out_set and i not in sliced and sizes[i] > 1]
        if not cand:
            return sliced
        best = None
        for i in cand:
            t2, l2, _ = path_cost(inputs, output, sizes, path, sliced + [i])
            key = (l2, t2 * sizes[i], i)
            if best is None or key < best[0]:
                best = (key, i)
        sliced.append(best[1])
        n_slices *= sizes[best[1]]


def _indices_on_large_intermediates(inputs, output, sizes, path, sliced, largest):
    sizes = dict(sizes)
    for i in sliced:
        sizes[i] = 1
    tensors = {i: tuple(ix) for i, ix in enumerate(inputs)}
    count: Dict[int, int] = {}
    for ix in tensors.values():
        for i in ix:
            count[i] = count.get(i, 0) + 1
    for i in output:
        count[i] = count.get(i, 0) + 1
    next_id = len(inputs)
    found = {}
    for ix in tensors.values():  # an input tensor can be the largest object too
        if _size(ix, sizes) * 2 >= largest:
            for i in ix:
                found[i] = found.get(i, 0) + 1
    for a, b in path:
        ia, ib = tensors.pop(a), tensors.pop(b)
        shared = [i for i in ia if i in ib]
        contracted = [i for i in shared if count[i] == 2]
        for i in shared:
            count[i] -= 1
        for i in contracted:
            count[i] -= 1
        res = tuple([i for i in ia if i not in contracted] + [i for i in ib if i not in contracted and i not in ia])
        if _size(res, sizes) * 2 >= largest:
            for i in res:
                found[i] = found.get(i, 0) + 1
        tensors[next_id] = res
        next_id += 1
    return sorted(found, key=lambda i: (-found[i], i))


# ------------------------------------------------------------------------------------------
# random-restart greedy ("hyper-greedy") with slicing in the objective
# ------------------------------------------------------------------------------------------
# achieved rates of the engine on one B200: algorithmic flop/s of the grouped GEMM, bytes per element, HBM passes of
# the matricize step per operand element (single precision: read once, hi/lo planes written and read back; double: read + write)
_DTYPE_MODEL = {
    "complex64": (8, 250e12, 8, 4.0), "float32": (2, 200e12, 4, 4.0),
    "complex128": (8, 44e12, 16, 2.0), "float64": (2, 35e12, 8, 2.0),
}


def step_time_model(m: int, n: int, k: int, itemsize: int = 8, tensor_flops: float = None, hbm_bytes: float = 4.5e12,
                    launch_s: float = 3e-6, dtype: str = None) -> float:
    """Seconds one pairwise step costs on a B200 under a two-term roofline, by the route the engine takes for it
    (``engine/plan.py``): a gate application (one operand of <= 16 rows, k <= 64) streams both operands and the result
    through HBM once; anything else pays the tensor work at the achieved rate of the grouped GEMM of ``dtype`` plus the HBM
    traffic of matricizing both operands and writing the result.  Plus a launch.  ``dtype`` None: complex64 rates with
    ``itemsize`` bytes per element (16 -> complex128).  Calibrated on Sycamore-53 m=14 trees (model 2.3 / 5.9 ms, measured
    1.9 / 5.4 ms); only used to RANK trees."""
    if dtype is None:
        dtype = "complex128" if itemsize == 16 else "complex64"
    per, rate, size, passes = _DTYPE_MODEL[str(dtype)]
    if tensor_flops is not None:
        rate = tensor_flops
    if min(m, n) <= 16 and k <= 64:
        return size * (m * k + n * k + m * n) / hbm_bytes + launch_s
    return per * m * n * k / rate + size * (passes * (m * k + n * k) + 2 * m * n) / hbm_bytes + launch_s


def search_path(inputs: Sequence[Sequence[int]], output: Sequence[int], sizes: Dict[int, int], trials: int = 128, seed: int = 0,
                target_size: int = None, minimize: str = "flops", alphas: Tuple[float, float] = (0.25, 8.0),
                temperatures: Tuple[float, float] = (0.0, 0.15), itemsize: int = 8, refine: int = 3,
                candidates: Sequence[Tuple[float, float, int]] = (), dtype: str = None):
    """Best of ``trials`` greedy trees: every trial draws the greedy cost's ``alpha`` log-uniformly from ``alphas`` and a
    Boltzmann ``temperature`` from ``temperatures`` (every fourth trial runs at temperature 0) — the role of cotengra's
    ``HyperOptimizer`` over its greedy driver in the reference's benchmarks (``benchmark/bench_sycamore.py:75-82``,
    ``bench_random_tn.py:80-110``).  Deterministic for a given ``seed``.

    ``minimize``: ``"flops"`` (total multiply-adds over all slices), ``"size"`` (largest intermediate, then flops) or
    ``"time"`` (``step_time_model`` summed over the steps of all slices, at the rates of ``dtype`` - or of complex64 /
    complex128 by ``itemsize`` when ``dtype`` is None).  ``target_size``: largest intermediate allowed, in
    elements; trees above it are sliced with ``find_slices`` and the slicing overhead counts (the ``refine`` best trees by a
    cheap estimate — work x width / target — are sliced for real, the others are not).

    ``candidates``: ``(alpha, temperature, trial_seed)`` triples evaluated next to the random trials — trees recorded by
    an earlier, longer search (``info`` carries the triple that rebuilds a tree), the role of the path cache the reference's
    benchmarks keep on disk (``ReusableHyperOptimizer(directory=...)``, ``benchmark/bench_sycamore.py:75-82``).

    Returns ``(path, sliced, info)``; ``info``: ``alpha``, ``temperature``, ``trial_seed``, ``madds`` (one slice),
    ``n_slices``, ``total_madds``, ``largest`` (elements, after slicing), ``time_model_s``, ``trials``, and ``baseline`` = the same
    figures for the plain ``greedy_path(seed=seed)``, which is trial 0 (the result is never worse than it)."""
    if minimize not in ("flops", "size", "time"):
        raise ValueError("minimize must be 'flops', 'size' or 'time'")
    rng = np.random.default_rng(seed)
    lo_a, hi_a = float(alphas[0]), float(alphas[1])
    cands = []

    def evaluate(path, sliced=()):
        madds, largest, steps = path_cost(inputs, output, sizes, path, sliced)
        n_sl = prod(sizes[i] for i in sliced) if sliced else 1
        t = n_sl * sum(step_time_model(m, n, k, itemsize, dtype=dtype) for m, n, k in steps)
        return madds, largest, n_sl, t

    def score(madds, largest, n_sl, t):
        if minimize == "flops":
            return (madds * n_sl, largest)
        if minimize == "size":
            return (largest, madds * n_sl)
        return (t, madds * n_sl)

    recorded = [(float(a), float(t), int(sd)) for a, t, sd in candidates]
    n_random = max(int(trials), 1)
    for trial in range(n_random + len(recorded)):
        if trial == 0:
            alpha, temp, tseed = 1.0, 0.0, seed          # the plain greedy tree
        elif trial >= n_random:
            alpha, temp, tseed = recorded[trial - n_random]
        else:
            alpha = float(np.exp(rng.uniform(np.log(lo_a), np.log(hi_a))))
            temp = 0.0 if trial % 4 == 1 else float(rng.uniform(temperatures[0], temperatures[1]))
            tseed = int(rng.integers(0, 2 ** 31 - 1))
        path = greedy_path(inputs, output, sizes, seed=tseed, alpha=alpha, temperature=temp)
        madds, largest, n_sl, t = evaluate(path)
        over = max(1.0, largest / target_size) if target_size else 1.0
        # cheap pre-ranking of trees that still need slicing: ideal slicing has overhead 1, wide trees slice badly in
        # practice - rank as if the overhead grew with the square root of the excess width; the finalists are sliced for real
        est = score(madds * over ** 0.5, largest, 1, t * over ** 0.5)
        cands.append(dict(est=est, path=path, alpha=alpha, temperature=temp, trial_seed=tseed, madds=madds, largest=largest, over=over))
    baseline = cands[0]
    order = sorted(range(len(cands)), key=lambda i: (cands[i]["est"], i))
    finalists = order[:max(int(refine), 1)] if target_size else order[:1]
    if 0 not in finalists:
        finalists.append(0)
    best, base_rec = None, None
    for i in finalists:
        c = cands[i]
        sliced = []
        if target_size and c["largest"] > target_size:
            if c["over"] > 2 ** 16:          # hopeless: slicing would need more than 16 index cuts
                continue
            sliced = find_slices(inputs, output, sizes, c["path"], target_size)
        madds, largest, n_sl, t = evaluate(c["path"], sliced)
        if target_size and largest > target_size:
            continue
        rec = dict(alpha=c["alpha"], temperature=c["temperature"], trial_seed=c["trial_seed"], madds=madds, n_slices=n_sl,
                   total_madds=madds * n_sl, largest=largest, time_model_s=t)
        key = (score(madds, largest, n_sl, t), i)
        if i == 0:
            base_rec = dict(rec)
        if best is None or key < best[0]:
            best = (key, c["path"], sliced, rec)
    if best is None:
        raise ValueError("no tree fits target_size=%r" % (target_size,))
    info = best[3]
    info["trials"] = len(cands)
    info["baseline"] = base_rec          # None when the plain greedy tree does not fit target_size
    return best[1], best[2], info
