"""Tensor-network front-end, host side: generators, simplification, greedy path, cost model and
slice finder (no GPU: nothing here contracts on the device)."""
import numpy as np
import pytest

from rosnet_b200 import tn as T
from rosnet_b200.tn.network import simplify


def dense_value(net):
    """Reference value with numpy.einsum over integer labels (small networks only)."""
    args = []
    for arr, ix in zip(net.arrays, net.inputs):
        args += [arr, list(ix)]
    return np.einsum(*args, list(net.output), optimize="greedy")


def test_rand_equation_structure():
    net = T.rand_equation(30, 3, n_out=2, d_min=2, d_max=4, seed=1)
    assert net.n_tensors == 30 and len(net.output) == 2
    flat = [i for ix in net.inputs for i in ix]
    for i in net.sizes:
        assert flat.count(i) == (1 if i in net.output else 2)
        assert 2 <= net.sizes[i] <= 4
    assert all(len(ix) >= 1 and len(set(ix)) == len(ix) for ix in net.inputs)
    again = T.rand_equation(30, 3, n_out=2, d_min=2, d_max=4, seed=1)
    assert again.inputs == net.inputs and again.sizes == net.sizes


def test_path_cost_matches_bruteforce_einsum_path():
    net = T.rand_equation(12, 3, n_out=1, d_min=2, d_max=3, seed=3)
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    assert len(path) == net.n_tensors - 1
    madds, largest, steps = T.path_cost(net.inputs, net.output, net.sizes, path)
    assert madds == sum(m * n * k for m, n, k in steps) and largest >= 1
    # the path is deterministic
    assert path == T.greedy_path(net.inputs, net.output, net.sizes, seed=0)


def test_slice_finder_reaches_target():
    net = T.rand_equation(40, 3, d_min=2, d_max=2, seed=5)
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    target = max(largest // 8, 2)
    sliced = T.find_slices(net.inputs, net.output, net.sizes, path, target)
    _, l2, _ = T.path_cost(net.inputs, net.output, net.sizes, path, sliced)
    assert l2 <= target and len(set(sliced)) == len(sliced) > 0


def test_qft_network_value_against_statevector():
    n = 5
    net = T.qft_amplitude_network(n, dtype="complex128")
    assert net.n_tensors == 2 * n + (n - 1) + (n - 1) * n // 2
    # statevector simulation of the same gate list (H(i); CZ(j,i) for j>i)
    psi = np.zeros(2**n, dtype=complex)
    psi[0] = 1
    psi = psi.reshape((2,) * n)
    H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    for i in range(n - 1):
        psi = np.moveaxis(np.tensordot(H, psi, (1, i)), 0, i)
        for j in range(i + 1, n):
            sl = [slice(None)] * n
            sl[i], sl[j] = 1, 1
            psi[tuple(sl)] *= -1
    want = psi[(0,) * n]
    assert abs(dense_value(net) - want) < 1e-12
    assert abs(dense_value(simplify(net)) - want) < 1e-12
    assert simplify(net).n_tensors < net.n_tensors


def test_sycamore_layout_and_network():
    from rosnet_b200.tn.network import _sycamore_layout

    nq, couplers = _sycamore_layout()
    assert nq == 53
    assert sum(len(v) for v in couplers.values()) == 86        # couplers of the 53-qubit device graph
    for cls in couplers.values():                               # each class is a matching
        qs = [q for pair in cls for q in pair]
        assert len(qs) == len(set(qs))
    net = T.sycamore_amplitude_network(cycles=2, seed=0)
    flat = [i for ix in net.inputs for i in ix]
    assert all(flat.count(i) == 2 for i in net.sizes) and net.output == ()
    small = simplify(net)
    assert small.n_tensors < net.n_tensors and max(len(ix) for ix in small.inputs) <= 4


def test_headline_workloads_are_analysed():
    """Width/cost of the BASELINE TN configs under the build's greedy path (documented in DESIGN.md)."""
    from math import log2

    net = T.rand_equation(200, 3, 0, 16, 16, seed=0)
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    assert log2(largest) > 100          # config 2 at bond 16 cannot be contracted unsliced
    q = simplify(T.qft_amplitude_network(30))
    path = T.greedy_path(q.inputs, q.output, q.sizes, seed=0)
    madds, largest, _ = T.path_cost(q.inputs, q.output, q.sizes, path)
    assert log2(largest) <= 30 and log2(madds) < 40


def _contract_numpy(net, path, sliced=()):
    """Value of the network along ``path`` (summed over the slices of ``sliced``) with numpy.tensordot per pairwise step."""
    from itertools import product

    total = 0
    for combo in product(*[range(net.sizes[i]) for i in sliced]):
        fixed = dict(zip(sliced, combo))
        tensors = {}
        for t, (arr, ix) in enumerate(zip(net.arrays, net.inputs)):
            arr = np.asarray(arr).astype(np.complex128)
            keep = []
            for lab in ix:
                if lab in fixed:
                    arr = np.take(arr, fixed[lab], axis=len(keep))
                else:
                    keep.append(lab)
            tensors[t] = (arr, tuple(keep))
        nxt = len(tensors)
        for a, b in path:
            (A, ia), (B, ib) = tensors.pop(a), tensors.pop(b)
            sh = [i for i in ia if i in ib]
            C = np.tensordot(A, B, ([ia.index(i) for i in sh], [ib.index(i) for i in sh]))
            tensors[nxt] = (C, tuple(i for i in ia if i not in sh) + tuple(i for i in ib if i not in sh))
            nxt += 1
        assert len(tensors) == 1
        total = total + list(tensors.values())[0][0]
    return total


def test_search_path_is_deterministic_and_never_worse_than_plain_greedy():
    net = simplify(T.sycamore_amplitude_network(6, seed=0, dtype="complex64"))
    path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=24, seed=3)
    again = T.search_path(net.inputs, net.output, net.sizes, trials=24, seed=3)
    assert (path, sliced) == again[:2] and info == again[2]
    assert len(path) == net.n_tensors - 1 and sliced == []
    plain = T.greedy_path(net.inputs, net.output, net.sizes, seed=3)
    madds_plain, largest_plain, _ = T.path_cost(net.inputs, net.output, net.sizes, plain)
    assert info["baseline"]["madds"] == madds_plain and info["baseline"]["largest"] == largest_plain
    assert info["total_madds"] <= madds_plain
    madds, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    assert (madds, largest, 1) == (info["madds"], info["largest"], info["n_slices"]) and info["trials"] == 24
    # the tree found is replayable from its recorded parameters
    replay = T.greedy_path(net.inputs, net.output, net.sizes, seed=info["trial_seed"], alpha=info["alpha"], temperature=info["temperature"])
    if info["reconfigured"]:
        replay = T.reconfigure_path(net.inputs, net.output, net.sizes, replay)
    assert path == replay
    # without the local search the result is one of the greedy trees themselves
    p2, _, i2 = T.search_path(net.inputs, net.output, net.sizes, trials=24, seed=3, reconfigure=False)
    assert not i2["reconfigured"] and info["total_madds"] <= i2["total_madds"]
    assert p2 == T.greedy_path(net.inputs, net.output, net.sizes, seed=i2["trial_seed"], alpha=i2["alpha"], temperature=i2["temperature"])


def test_search_path_value_is_path_independent():
    net = simplify(T.sycamore_amplitude_network(4, seed=1, dtype="complex128"))
    plain = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    want = _contract_numpy(net, plain)
    for minimize in ("flops", "size", "time"):
        path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=16, seed=0, minimize=minimize)
        got = _contract_numpy(net, path, sliced)
        assert abs(got - want) <= 1e-12 * abs(want), (minimize, got, want)


def test_search_path_slices_down_to_the_target_and_counts_the_overhead():
    net = simplify(T.sycamore_amplitude_network(6, seed=0, dtype="complex128"))
    free, _, info0 = T.search_path(net.inputs, net.output, net.sizes, trials=12, seed=0)
    target = max(info0["largest"] // 8, 2)
    path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=12, seed=0, target_size=target)
    assert info["largest"] <= target and sliced
    n_slices = int(np.prod([net.sizes[i] for i in sliced]))
    madds, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path, sliced)
    assert info["n_slices"] == n_slices and info["total_madds"] == madds * n_slices and largest == info["largest"]
    assert info["total_madds"] >= info0["total_madds"]         # slicing never makes the whole job cheaper than the best free tree
    want = _contract_numpy(net, free)
    got = _contract_numpy(net, path, sliced)
    assert abs(got - want) <= 1e-12 * abs(want)
    with pytest.raises(ValueError):
        T.search_path(net.inputs, net.output, net.sizes, trials=2, minimize="speed")


def test_searched_tree_of_the_headline_circuit_is_a_chain_of_gate_applications():
    """Sycamore-53 m=14: the searched tree needs ~2^35 flop unsliced (the headline tree: 8 slices of 2^43.9) and almost all of
    its work is steps with one operand of <= 16 rows and k <= 64 - the shapes of the apply route (engine/plan.py: _nd_apply)."""
    from math import log2

    net = simplify(T.sycamore_amplitude_network(14, seed=0, dtype="complex64"))
    path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=128, seed=0, target_size=2 ** 28, reconfigure=False)
    assert sliced == [] and log2(info["largest"]) <= 27 and log2(8 * info["total_madds"]) < 37
    _, _, steps = T.path_cost(net.inputs, net.output, net.sizes, path)
    work = sum(m * n * k for m, n, k in steps)
    gate_like = sum(m * n * k for m, n, k in steps if min(m, n) <= 16 and k <= 64)
    assert gate_like >= 0.99 * work
    # the local search on top (reconfigure_path) trades the pure chain for fewer flops still: 2^35.3 -> below 2^34.5
    path2, sliced2, info2 = T.search_path(net.inputs, net.output, net.sizes, trials=128, seed=0, target_size=2 ** 28)
    assert sliced2 == [] and info2["reconfigured"] and log2(8 * info2["total_madds"]) < 34.5 and info2["largest"] <= info["largest"]


def test_search_path_takes_recorded_trees():
    """``candidates``: (alpha, temperature, trial_seed) triples of earlier searches compete with the fresh trials; ``info`` of a
    result is such a triple (the bench keeps the trees of long offline searches this way, as the reference's benchmarks keep
    cotengra's path cache on disk)."""
    net = simplify(T.sycamore_amplitude_network(6, seed=0, dtype="complex64"))
    long_path, _, long_info = T.search_path(net.inputs, net.output, net.sizes, trials=48, seed=5, minimize="time")
    triple = (long_info["alpha"], long_info["temperature"], long_info["trial_seed"])
    path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=2, seed=0, minimize="time", candidates=[triple])
    assert info["time_model_s"] <= long_info["time_model_s"] * (1 + 1e-12) and info["trials"] == 3
    short = T.search_path(net.inputs, net.output, net.sizes, trials=2, seed=0, minimize="time")[2]
    if short["time_model_s"] > long_info["time_model_s"]:
        assert path == long_path and (info["alpha"], info["temperature"], info["trial_seed"]) == triple


def test_reconfigure_path_keeps_the_value_and_never_costs_more():
    """Optimal re-contraction of subtrees (tn/path.py: reconfigure_path): same network value, total cost never above the
    tree it started from, intermediates within the limit, deterministic, output of every objective a valid SSA path."""
    from math import prod

    for cycles, seed in ((4, 0), (5, 3), (6, 1)):
        net = simplify(T.sycamore_amplitude_network(cycles, seed=seed, dtype="complex128"))
        tree = T.greedy_path(net.inputs, net.output, net.sizes, seed=seed)
        m0, w0, _ = T.path_cost(net.inputs, net.output, net.sizes, tree)
        want = _contract_numpy(net, tree)
        for minimize in ("flops", "time"):
            for S in (4, 8):
                new = T.reconfigure_path(net.inputs, net.output, net.sizes, tree, subtree_size=S, minimize=minimize)
                assert new == T.reconfigure_path(net.inputs, net.output, net.sizes, tree, subtree_size=S, minimize=minimize)
                assert len(new) == len(tree) and sorted(x for pr in new for x in pr) == list(range(2 * len(tree)))    # every id consumed once
                m1, w1, steps = T.path_cost(net.inputs, net.output, net.sizes, new)
                assert w1 <= w0
                if minimize == "flops":
                    assert m1 <= m0
                else:
                    t0 = sum(T.step_time_model(*s_) for s_ in T.path_cost(net.inputs, net.output, net.sizes, tree)[2])
                    assert sum(T.step_time_model(*s_) for s_ in steps) <= t0 * (1 + 1e-9)
                got = _contract_numpy(net, new)
                assert abs(got - want) <= 1e-12 * abs(want), (cycles, minimize, S)
    # open indices survive: a network with output labels
    net = T.rand_equation(12, 3, n_out=2, d_min=2, d_max=3, seed=4)
    net.random_arrays("float64", seed=2)
    tree = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    new = T.reconfigure_path(net.inputs, net.output, net.sizes, tree, subtree_size=6)
    assert T.path_cost(net.inputs, net.output, net.sizes, new)[0] <= T.path_cost(net.inputs, net.output, net.sizes, tree)[0]
    for seed in range(6):                       # random networks with open indices, several extents: value against numpy.einsum
        rn_ = T.rand_equation(10, 3, n_out=seed % 3, d_min=2, d_max=4, seed=20 + seed)
        rn_.random_arrays("float64", seed=seed)
        tr = T.greedy_path(rn_.inputs, rn_.output, rn_.sizes, seed=seed)
        nw = T.reconfigure_path(rn_.inputs, rn_.output, rn_.sizes, tr, subtree_size=5 + seed % 4, minimize=("flops", "time")[seed % 2])
        want_ = dense_value(rn_)
        # walk the new tree with numpy and put the open indices in output order
        tensors = {i: (np.asarray(a), tuple(ix)) for i, (a, ix) in enumerate(zip(rn_.arrays, rn_.inputs))}
        nxt = len(tensors)
        for a_, b_ in nw:
            (A, ia), (B, ib) = tensors.pop(a_), tensors.pop(b_)
            sh = [i for i in ia if i in ib and i not in rn_.output]
            Cc = np.tensordot(A, B, ([ia.index(i) for i in sh], [ib.index(i) for i in sh]))
            tensors[nxt] = (Cc, tuple(i for i in ia if i not in sh) + tuple(i for i in ib if i not in sh))
            nxt += 1
        (res, labs), = tensors.values()
        got_ = np.transpose(res, [labs.index(i) for i in rn_.output]) if labs else res
        assert np.allclose(got_, want_, rtol=1e-10, atol=1e-12), seed
    with pytest.raises(ValueError):
        T.reconfigure_path(net.inputs, net.output, net.sizes, tree, minimize="width")


def test_search_path_edge_cases():
    """One- and two-tensor networks, disconnected components, infeasible memory limits."""
    cases = [
        ([(0, 1)], (0, 1), {0: 2, 1: 3}),
        ([(0, 1), (1, 2)], (0, 2), {0: 2, 1: 3, 2: 4}),
        ([(0, 1), (1, 0), (2, 3), (3, 2)], (), {0: 2, 1: 3, 2: 2, 3: 5}),                    # two components
        ([(0,), (0, 1), (1, 2), (2,), (3,), (3,)], (), {0: 2, 1: 3, 2: 2, 3: 7}),
    ]
    for inputs, output, sizes in cases:
        for minimize in ("flops", "time", "size"):
            path, sliced, info = T.search_path(inputs, output, sizes, trials=6, seed=0, minimize=minimize)
            assert len(path) == len(inputs) - 1 and sliced == [] and info["n_slices"] == 1
            assert sorted(x for pr in path for x in pr) == list(range(2 * len(path)))
        tree = T.greedy_path(inputs, output, sizes)
        new = T.reconfigure_path(inputs, output, sizes, tree)
        assert len(new) == len(tree) and T.path_cost(inputs, output, sizes, new)[0] <= T.path_cost(inputs, output, sizes, tree)[0]
    # open indices cannot be sliced: a result larger than the limit has no feasible tree
    with pytest.raises(ValueError):
        T.search_path(cases[1][0], cases[1][1], cases[1][2], trials=4, target_size=2)
    # closed networks slice down to scalars
    path, sliced, info = T.search_path(cases[3][0], cases[3][1], cases[3][2], trials=4, target_size=2)
    assert info["largest"] <= 2 and info["n_slices"] == int(np.prod([cases[3][2][i] for i in sliced]))
