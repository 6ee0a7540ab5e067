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
                candidates: Sequence[Tuple[float, float, int]] = (), dtype: str = None, reconfigure: bool = True,
                subtree_size: int = 8):
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

    ``reconfigure``: the finalists (the ``refine`` best trees and the plain greedy one) go through ``reconfigure_path`` (optimal
    re-contraction of subtrees of ``subtree_size`` nodes, same objective) before they are sliced and compared; ``info["reconfigured"]``
    says whether the returned tree did — it is then ``reconfigure_path(greedy_path(alpha, temperature, trial_seed), ...)``.

    Returns ``(path, sliced, info)``; ``info``: ``alpha``, ``temperature``, ``trial_seed``, ``madds`` (one slice),
    ``n_slices``, ``total_madds``, ``largest`` (elements, after slicing), ``time_model_s``, ``reconfigured``, ``trials``, and
    ``baseline`` = the same figures for the plain ``greedy_path(seed=seed)``, which is trial 0 (the result is never worse than it)."""
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
    for extra in [0] + list(range(n_random, len(cands))):       # the plain greedy tree and every recorded tree are finalists too
        if extra not in finalists:
            finalists.append(extra)
    best, base_rec = None, None

    def finish(i, tree, reconfigured):
        nonlocal best, base_rec
        c = cands[i]
        madds, largest, _, _ = evaluate(tree)
        sliced = []
        if target_size and largest > target_size:
            if largest / target_size > 2 ** 16:          # hopeless: slicing would need more than 16 index cuts
                return
            sliced = find_slices(inputs, output, sizes, tree, target_size)
        madds, largest, n_sl, t = evaluate(tree, sliced)
        if target_size and largest > target_size:
            return
        rec = dict(alpha=c["alpha"], temperature=c["temperature"], trial_seed=c["trial_seed"], madds=madds, n_slices=n_sl,
                   total_madds=madds * n_sl, largest=largest, time_model_s=t, reconfigured=reconfigured)
        key = (score(madds, largest, n_sl, t), i, reconfigured)
        if i == 0 and not reconfigured:
            base_rec = dict(rec)
        if best is None or key < best[0]:
            best = (key, tree, sliced, rec)

    for i in finalists:
        finish(i, cands[i]["path"], False)
        if reconfigure and len(cands[i]["path"]) >= 3:
            limit = max(cands[i]["largest"], target_size) if target_size else None
            tree = reconfigure_path(inputs, output, sizes, cands[i]["path"], subtree_size=subtree_size, minimize="time" if minimize == "time" else "flops",
                                    max_size=limit, itemsize=itemsize, dtype=dtype)
            if tree != cands[i]["path"]:
                finish(i, tree, True)
    if best is None:
        raise ValueError("no tree fits target_size=%r" % (target_size,))
    info = best[3]
    info["trials"] = len(cands)
    info["baseline"] = base_rec          # None when the plain greedy tree does not fit target_size
    return best[1], best[2], info


# ------------------------------------------------------------------------------------------
# subtree reconfiguration: optimal re-contraction of small subtrees of a given tree
# ------------------------------------------------------------------------------------------
def _tree_of(n_inputs: int, path):
    """children[node] = (left, right) for the internal nodes of an SSA path (inputs are 0..n-1; the k-th pair creates node n+k)."""
    children = {}
    for k, (a, b) in enumerate(path):
        children[n_inputs + k] = (a, b)
    return children


def reconfigure_path(inputs: Sequence[Sequence[int]], output: Sequence[int], sizes: Dict[int, int], path, subtree_size: int = 8,
                     minimize: str = "flops", max_size: int = None, passes: int = 2, itemsize: int = 8, dtype: str = None):
    """Local search on a contraction tree (cotengra's ``subtree_reconfigure``): for every internal node, take the subtree made of
    up to ``subtree_size`` of its descendants (the frontier grown breadth-first, largest intermediate first), find the OPTIMAL
    order of contracting that frontier by dynamic programming over its subsets, and splice it in when it is cheaper.  The rest
    of the tree is untouched, so the value of the network is unchanged.  ``minimize``: ``"flops"`` or ``"time"``
    (``step_time_model``); ``max_size``: no intermediate of a re-contracted subtree may exceed this many elements (default: the
    tree's current largest intermediate).  Returns a new SSA path (connected networks; disconnected components are left as
    they are)."""
    if minimize not in ("flops", "time"):
        raise ValueError("minimize must be 'flops' or 'time'")
    n = len(inputs)
    if len(path) != n - 1 or n < 3:
        return list(path)
    labels = sorted(sizes)
    bit = {lab: i for i, lab in enumerate(labels)}
    lg = [sizes[lab] for lab in labels]
    out_mask = 0
    for lab in output:
        out_mask |= 1 << bit[lab]

    def mask_of(ix):
        m = 0
        for lab in ix:
            m |= 1 << bit[lab]
        return m

    def size_of(mask):
        s = 1
        while mask:
            low = mask & -mask
            s *= lg[low.bit_length() - 1]
            mask ^= low
        return s

    children = _tree_of(n, path)
    root = n + len(path) - 1
    leaf_mask = [mask_of(ix) for ix in inputs]
    # hyper-edges (an index on more than two tensors) are not handled by the pairwise bookkeeping below
    cnt = {}
    for ix in inputs:
        for lab in ix:
            cnt[lab] = cnt.get(lab, 0) + 1
    if any(c + (1 if lab in output else 0) > 2 for lab, c in cnt.items()):
        return list(path)

    def node_masks():
        """index mask of every node's tensor: indices of its leaves that also live outside the node's leaf set (or are open)"""
        inside = {}
        stack = [(root, False)]
        while stack:                                   # explicit post-order: a chain-shaped tree is n levels deep
            v, done = stack.pop()
            if v < n:
                inside[v] = leaf_mask[v]
            elif done:
                l, r = children[v]
                ml, mr = inside[l], inside[r]
                # every index is on exactly two tensors (or one + output): shared between the two sides -> contracted here
                inside[v] = (ml | mr) & ~(ml & mr & ~out_mask)
            else:
                stack.append((v, True))
                stack.append((children[v][1], False))
                stack.append((children[v][0], False))
        return inside

    def cost_of(ma, mb):
        """(flops-or-time, result mask) of contracting tensors with index masks ma, mb"""
        shared = ma & mb & ~out_mask
        res = (ma | mb) & ~shared
        if minimize == "flops":
            return size_of(ma | mb), res
        k = size_of(shared)
        m_ = size_of(ma & ~shared)
        n_ = size_of(mb & ~shared)
        return step_time_model(m_, n_, k, itemsize, dtype=dtype), res

    def tree_cost():
        masks = node_masks()
        total, largest = 0, 0
        for v, (l, r) in children.items():
            c, res = cost_of(masks[l], masks[r])
            total += c
            largest = max(largest, size_of(res))
        return total, largest, masks

    total0, largest0, masks = tree_cost()
    limit = max_size if max_size is not None else largest0
    next_id = [root + 1]

    for _ in range(max(int(passes), 1)):
        improved = False
        order = sorted(children, reverse=True)            # top-down: big, late steps first
        for v in order:
            if v not in children:
                continue
            # frontier: start from v's children, expand the largest internal node until subtree_size leaves
            frontier = list(children[v])
            internal = [v]
            while len(frontier) < subtree_size:
                cand = [u for u in frontier if u in children]
                if not cand:
                    break
                u = max(cand, key=lambda w: size_of(masks[w]))
                frontier.remove(u)
                frontier.extend(children[u])
                internal.append(u)
            S = len(frontier)
            if S < 3:
                continue
            fm = [masks[u] for u in frontier]
            old = 0
            for u in internal:
                old += cost_of(masks[children[u][0]], masks[children[u][1]])[0]
            # indices that leave the subtree: on v's own tensor
            ext = masks[v]
            # DP over subsets of the frontier
            full = (1 << S) - 1
            sub_mask = [0] * (1 << S)          # index mask of the tensor a subset contracts to
            union = [0] * (1 << S)
            for U in range(1, 1 << S):
                low = U & -U
                i = low.bit_length() - 1
                union[U] = union[U ^ low] | fm[i]
            for U in range(1, 1 << S):
                rest = union[full ^ U] if U != full else 0
                sub_mask[U] = union[U] & (rest | ext | out_mask)
            best = {}
            split = {}
            for i in range(S):
                best[1 << i] = 0
            for U in range(1, 1 << S):
                if U & (U - 1) == 0:
                    continue
                b, sp = None, None
                # enumerate splits A | B = U with the lowest set bit of U in A
                low = U & -U
                rest_bits = U ^ low
                A_sub = rest_bits
                while True:
                    A = low | A_sub
                    B = U ^ A
                    if B:
                        ca, cb = best.get(A), best.get(B)
                        if ca is not None and cb is not None:
                            ma, mb = sub_mask[A], sub_mask[B]
                            if minimize == "flops":
                                c = size_of(ma | mb)
                            else:
                                shared = ma & mb & ~sub_mask[U]
                                c = step_time_model(size_of(ma & ~shared), size_of(mb & ~shared), size_of(shared), itemsize, dtype=dtype)
                            if size_of(sub_mask[U]) <= limit or U == full:
                                tot = ca + cb + c
                                if b is None or tot < b:
                                    b, sp = tot, (A, B)
                    if A_sub == 0:
                        break
                    A_sub = (A_sub - 1) & rest_bits
                if b is not None:
                    best[U], split[U] = b, sp
            if full not in best or not best[full] < old * (1 - 1e-9):
                continue
            # splice: rebuild the internal nodes of the subtree following the optimal splits; v keeps its id (its parent points to it)
            for u in internal:
                del children[u]

            def build(U, top=False):
                if U & (U - 1) == 0:
                    return frontier[U.bit_length() - 1]
                A, B = split[U]
                a_, b_ = build(A), build(B)
                if top:
                    node = v
                else:
                    node = next_id[0]
                    next_id[0] += 1
                children[node] = (a_, b_)
                return node
            build(full, top=True)
            _, _, masks = tree_cost()
            improved = True
        if not improved:
            break

    # emit an SSA path in post-order
    new_path, ssa = [], {}
    counter = [n]

    def emit(v):
        if v < n:
            return v
        stack = [(v, False)]
        while stack:
            node, done = stack.pop()
            if node < n or node in ssa:
                continue
            l, r = children[node]
            if done:
                a_ = l if l < n else ssa[l]
                b_ = r if r < n else ssa[r]
                new_path.append((a_, b_))
                ssa[node] = counter[0]
                counter[0] += 1
            else:
                stack.append((node, True))
                stack.append((r, False))
                stack.append((l, False))
        return ssa[v]
    emit(root)
    return new_path
