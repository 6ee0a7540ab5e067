"""Tensor-network contraction through the blocked path on the GPU vs numpy.einsum."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dense_value(net):
    args = []
    for arr, ix in zip(net.arrays, net.inputs):
        args += [arr, list(ix)]
    return np.einsum(*args, list(net.output), optimize="greedy")


def cpu_contract(net, path):
    """Pairwise numpy.tensordot along `path` (networks with more index labels than einsum has letters)."""
    tensors = {i: (np.asarray(a), tuple(ix)) for i, (a, ix) in enumerate(zip(net.arrays, net.inputs))}
    nxt = len(tensors)
    for a, b in path:
        (A, ia), (B, ib) = tensors.pop(a), tensors.pop(b)
        shared = [i for i in ia if i in ib]
        C = np.tensordot(A, B, ([ia.index(i) for i in shared], [ib.index(i) for i in shared]))
        tensors[nxt] = (C, tuple(i for i in ia if i not in shared) + tuple(i for i in ib if i not in shared))
        nxt += 1
    (res, labels), = tensors.values()
    return np.transpose(res, [labels.index(i) for i in net.output]) if labels else res


def _rel(x, ref):
    ref = np.asarray(ref)
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(ref.ravel()) or 1.0))


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("complex128", 1e-12), ("complex64", 1e-5), ("float32", 1e-5)])
def test_random_network_matches_einsum(dtype, tol, contract_route):
    from rosnet_b200 import tn as T

    net = T.rand_equation(14, 3, n_out=2, d_min=2, d_max=4, seed=7)
    net.random_arrays(dtype, seed=1)
    wide = [a.astype("complex128" if np.dtype(dtype).kind == "c" else "float64") for a in net.arrays]
    want = dense_value(T.TensorNetwork(net.inputs, net.output, net.sizes, wide))
    res, stats = T.contract_network(net)
    assert stats["steps"] == net.n_tensors - 1 and stats["flops"] > 0
    got = np.asarray(res)
    assert got.shape == want.shape
    assert _rel(got, want) <= tol


def test_slicing_loop_and_blocks_agree():
    from rosnet_b200 import tn as T

    net = T.rand_equation(16, 3, n_out=1, d_min=2, d_max=3, seed=11)
    net.random_arrays("complex128", seed=2)
    want = dense_value(net)
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    sliced = T.find_slices(net.inputs, net.output, net.sizes, path, max(largest // 6, 2))
    assert sliced
    full, _ = T.contract_network(net, path)
    loop, st = T.contract_network(net, path, sliced=sliced, slice_mode="loop")
    blocks, _ = T.contract_network(net, path, sliced=sliced, slice_mode="blocks")
    assert st["n_slices"] == st["total_slices"] > 1
    for got in (full, loop, blocks):
        assert _rel(got, want) <= 1e-12
    # round-robin halves (what two ranks would do) add up to the whole
    n = st["total_slices"]
    h0, _ = T.contract_network(net, path, sliced=sliced, slice_ids=range(0, n, 2))
    h1, _ = T.contract_network(net, path, sliced=sliced, slice_ids=range(1, n, 2))
    assert _rel(h0 + h1, want) <= 1e-12


def test_einsum_on_blockarrays():
    import rosnet_b200 as rn

    rng = np.random.default_rng(3)
    a, b, c = rng.standard_normal((6, 8)), rng.standard_normal((8, 4, 10)), rng.standard_normal((10, 6))
    A = rn.array(a, blockshape=(3, 4), inner="cuda")
    B = rn.array(b, blockshape=(4, 2, 5), inner="cuda")
    C = rn.array(c, blockshape=(5, 3))
    got = np.einsum("ij,jkl->ikl", A, B)          # NumPy protocol -> rosnet_b200.dispatch.einsum
    assert isinstance(got, rn.BlockArray) and got.shape == (6, 4, 10)
    assert _rel(got, np.einsum("ij,jkl->ikl", a, b)) < 1e-13
    got = rn.einsum("ij,jkl,lm->kmi", A, B, C)
    assert _rel(got, np.einsum("ij,jkl,lm->kmi", a, b, c)) < 1e-13
    got = rn.einsum("ij->ji", A)
    assert np.array_equal(np.asarray(got), a.T)
    with pytest.raises(NotImplementedError):
        rn.einsum("...i->i", A)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("complex64", 1e-5)])
def test_einsum_general_patterns(dtype, tol):
    """Batch indices, hyper-edges, diagonals / traces and summed-out indices (np.einsum patterns the reference's COMPSs
    einsum hands to one np.einsum task per block, rosnet/array/compss/__init__.py:554-568)."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(12)

    def rand(shape):
        x = rng.standard_normal(shape)
        if np.dtype(dtype).kind == "c":
            x = x + 1j * rng.standard_normal(shape)
        return x.astype(dtype)

    wide = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
    a, b, c = rand((6, 8, 10)), rand((6, 10, 4)), rand((8, 8, 6))
    v, w, u = rand((12,)), rand((12,)), rand((12,))
    sq = rand((8, 8))
    A = rn.array(a, blockshape=(3, 4, 5), inner="cuda")
    B = rn.array(b, blockshape=(2, 5, 2), inner="cuda")
    C = rn.array(c, blockshape=(4, 4, 3), inner="cuda")
    V, W, U = (rn.array(x, blockshape=(4,), inner="cuda") for x in (v, w, u))
    SQ = rn.array(sq, blockshape=(4, 4), inner="cuda")
    cases = [
        ("bij,bjk->bik", (A, B), (a, b)),            # batched matmul: b is kept
        ("bij,bjk->ikb", (A, B), (a, b)),
        ("bij,bjk->ik", (A, B), (a, b)),             # b shared and summed: contracted together with j
        ("bij,bjk,iib->k", (A, B, C), (a, b, c)),    # hyper-edge b (three carriers), diagonal of c, all summed
        ("bij,bij->bij", (A, A), (a, a)),            # pure batch (Hadamard)
        ("i,i,i->", (V, W, U), (v, w, u)),           # hyper-edge summed
        ("i,i,i->i", (V, W, U), (v, w, u)),
        ("ii->i", (SQ,), (sq,)),
        ("ii", (SQ,), (sq,)),                        # trace (implicit output)
        ("iij->j", (C,), (c,)),
        ("bij->j", (A,), (a,)),                      # two indices summed out of one operand
        ("ij,k->ijk", (SQ, V), (sq, v)),             # outer product
    ]
    for pattern, blocked, dense in cases:
        got = np.asarray(rn.einsum(pattern, *blocked))
        want = np.einsum(pattern, *[x.astype(wide) for x in dense])
        assert got.shape == want.shape, pattern
        assert _rel(got, want) <= tol, (pattern, _rel(got, want))
    got = np.einsum("bij,bjk->bik", A, B)            # through the NumPy protocol
    assert isinstance(got, rn.BlockArray) and got.grid[0] == 6 and got.blockshape[0] == 1


def test_qft_amplitude_and_sycamore_small():
    from rosnet_b200 import tn as T
    from rosnet_b200.tn.network import simplify

    net = simplify(T.qft_amplitude_network(8, dtype="complex128"))
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    want = complex(cpu_contract(net, path))
    got, _ = T.contract_network(net, path)
    assert abs(complex(np.asarray(got)) - complex(want)) <= 1e-12 * max(abs(want), 1e-3) + 1e-14
    syc = simplify(T.sycamore_amplitude_network(cycles=3, seed=1, dtype="complex64"))
    wide = T.TensorNetwork(syc.inputs, syc.output, syc.sizes, [a.astype(np.complex128) for a in syc.arrays])
    path = T.greedy_path(syc.inputs, syc.output, syc.sizes, seed=0, alpha=0.5)
    want = complex(cpu_contract(wide, path))
    got, stats = T.contract_network(syc, path)
    assert abs(complex(np.asarray(got)) - want) <= 1e-5 * abs(want) + 1e-9


def test_cuda_graph_replay_rereads_inputs():
    """graph=True: first call eager, second captured, later calls replayed.  Every call gets NEW input
    values (same shapes): a replay that kept the old inputs or the old result would be caught."""
    from rosnet_b200 import tn as T
    from rosnet_b200.tn import contract as C

    net = T.rand_equation(12, 3, n_out=1, d_min=2, d_max=3, seed=5)
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0)
    _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    sliced = T.find_slices(net.inputs, net.output, net.sizes, path, max(largest // 4, 2))
    for mode in ("blocks", "loop"):
        replays = 0
        for trial in range(4):
            net.random_arrays("complex128", seed=100 + trial)
            want = dense_value(net)
            got, _ = T.contract_network(net, path, sliced=sliced, slice_mode=mode, graph=True)
            assert _rel(got, want) <= 1e-12, (mode, trial)
            gp = C._graphed_pass(net, path, sliced, mode)
            replays += gp.graph is not None
        assert replays >= 2 and not gp.failed, (mode, replays, gp.failed)


def test_sycamore_m14_full_size_slicing_invariance():
    """BASELINE configs[4] at full size (Sycamore-53, m=14, complex64).  No closed form exists for the
    amplitude, so the check is a size-independent property: the sum over slices must not depend on how
    the network is sliced (8 slices of width 2^28 vs the finer slicing of width 2^27), and a slice
    subset computed alone must equal its share of the total.  Also exercises the CUDA-graph replay
    (first slice eager, second captured, the rest replayed)."""
    import bench
    from rosnet_b200 import tn as T

    spec = bench.TN_WORKLOADS["bench_sycamore"]
    net, path, sliced = bench.build_tn(spec)
    assert spec["dtype"] == "complex64" and len(sliced) >= 3
    total, st = T.contract_network(net, path, sliced=sliced, slice_mode="loop")
    assert st["n_slices"] == st["total_slices"] == 8
    finer = T.find_slices(net.inputs, net.output, net.sizes, path, 2 ** 27)
    assert len(finer) > len(sliced)
    total2, st2 = T.contract_network(net, path, sliced=finer, slice_mode="loop")
    a, b = complex(np.asarray(total)), complex(np.asarray(total2))
    print(f"sycamore m=14 amplitude {a:.6e} vs finer slicing {b:.6e}: rel. diff {abs(a - b) / abs(a):.2e} ({st2['n_slices']} slices)")
    assert abs(a) > 0 and abs(a - b) <= 1e-5 * abs(a), (a, b)
    # Porter-Thomas scale: |amplitude|^2 of a 53-qubit random circuit is of order 2^-53
    assert 2.0 ** -70 < abs(a) ** 2 < 2.0 ** -40
    half0, _ = T.contract_network(net, path, sliced=sliced, slice_ids=[0, 2, 4, 6])
    half1, _ = T.contract_network(net, path, sliced=sliced, slice_ids=[1, 3, 5, 7])
    assert abs(complex(np.asarray(half0)) + complex(np.asarray(half1)) - a) <= 1e-5 * abs(a)


def test_qft30_full_size_matches_the_cpu_path():
    """BASELINE configs[2] at FULL size: the 30-qubit QFT amplitude network, complex128, contracted as the bench does
    (slicing as blocks, greedy path) against pairwise numpy.tensordot on the host - the reference's arithmetic."""
    from rosnet_b200 import tn as T
    from rosnet_b200.tn.network import simplify

    net = simplify(T.qft_amplitude_network(30, dtype="complex128"))
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0, alpha=1.0)
    _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    sliced = T.find_slices(net.inputs, net.output, net.sizes, path, max(largest // 8, 2))
    want = complex(cpu_contract(net, path))
    got_blocks, st = T.contract_network(net, path, sliced=sliced, slice_mode="blocks")
    got_loop, _ = T.contract_network(net, path, sliced=sliced, slice_mode="loop")
    assert st["flops"] > 1e11
    for got in (got_blocks, got_loop):
        assert abs(complex(np.asarray(got)) - want) <= 1e-12 * abs(want), (complex(np.asarray(got)), want)
    assert abs(want - 2.0 ** -14.5) <= 1e-12      # 29 Hadamards act on |0..0>: <0..0| QFT |0..0> = 2^(-29/2)


def test_sycamore_m14_one_whole_slice_against_the_cpu_path():
    """BASELINE configs[4] at FULL size: ONE slice of the Sycamore-53 m=14 network (complex64, 1.66e13 flop) walked step by
    step on the GPU (rosnet_b200.tensordot) and on the host (numpy.tensordot in complex64, the reference's arithmetic,
    ~30 s): every intermediate tensor of up to 2^21 elements is compared, and the slice's amplitude."""
    import rosnet_b200 as rn
    from rosnet_b200 import tn as T
    from rosnet_b200.tn.contract import _slice_inputs
    from rosnet_b200.tn.network import simplify

    net = simplify(T.sycamore_amplitude_network(cycles=14, seed=0, dtype="complex64"))
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0, alpha=0.5)
    _, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    sliced = T.find_slices(net.inputs, net.output, net.sizes, path, 2 ** 28)
    assert largest > 2 ** 28 and sliced
    extents = [net.sizes[i] for i in sliced]
    sid = 5 % int(np.prod(extents))
    sub_inputs, sub_arrays = _slice_inputs(net, list(sliced), extents, sid)
    host = {i: (np.asarray(a), tuple(ix)) for i, (a, ix) in enumerate(zip(sub_arrays, sub_inputs))}
    dev = {i: (rn.array(a, inner="cuda") if a.ndim else rn.BlockArray(np.asarray(a)), tuple(ix)) for i, (a, ix) in enumerate(zip(sub_arrays, sub_inputs))}
    nxt, compared, worst = len(host), 0, 0.0
    for a, b in path:
        (A, ia), (B, ib) = host.pop(a), host.pop(b)
        (DA, _), (DB, _) = dev.pop(a), dev.pop(b)
        shared = [i for i in ia if i in ib]
        axes = ([ia.index(i) for i in shared], [ib.index(i) for i in shared])
        C = np.tensordot(A, B, axes)
        DC = rn.tensordot(DA, DB, axes)
        lab = tuple(i for i in ia if i not in shared) + tuple(i for i in ib if i not in shared)
        assert DC.shape == C.shape
        if C.size <= 2 ** 21:
            err = _rel(np.asarray(DC), C)
            worst = max(worst, err)
            compared += 1
            assert err <= 1e-5, (nxt, C.shape, err)
        host[nxt], dev[nxt] = (C, lab), (DC, lab)
        nxt += 1
    (res, _), = host.values()
    (dres, _), = dev.values()
    want, got = complex(res), complex(np.asarray(dres))
    assert compared >= 200
    assert abs(got - want) <= 1e-5 * abs(want) + 1e-12, (got, want)
    # the same slice through the public driver (CUDA-graph path off and on)
    drv, _ = T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=[sid], graph=False)
    assert abs(complex(np.asarray(drv)) - want) <= 1e-5 * abs(want) + 1e-12


def test_sycamore_m14_searched_trees_against_complex128_on_the_host():
    """BASELINE configs[4]'s circuit along SEARCHED trees (tn.search_path): the balanced tree of tensor-core sized
    steps recorded in bench.py's workload spec and a chain of gate applications (the apply route of rnb_contract_nd), unsliced,
    in complex64 on the GPU - against the balanced tree walked with numpy.tensordot in complex128 on the host (~1 s), and
    against each other.  The amplitude is a sum with heavy cancellation (|amp| ~ 1e-8 from O(1) tensors): <= 1e-5."""
    from rosnet_b200 import tn as T
    from rosnet_b200.engine import lib
    from rosnet_b200.tn.network import simplify

    net = simplify(T.sycamore_amplitude_network(14, seed=0, dtype="complex64"))
    balanced = T.greedy_path(net.inputs, net.output, net.sizes, seed=1783767436, alpha=0.61461059008156, temperature=0.02284424442732209)
    chain = T.greedy_path(net.inputs, net.output, net.sizes, seed=915743486, alpha=6.137498636489522, temperature=0.03991954084384389)
    mb, wb, _ = T.path_cost(net.inputs, net.output, net.sizes, balanced)
    mc, wc, steps_c = T.path_cost(net.inputs, net.output, net.sizes, chain)
    assert wb == 2 ** 24 and wc == 2 ** 26 and 8 * mb < 2 ** 38 and 8 * mc < 2 ** 36       # unsliced: 2^37.7 and 2^35.3 flop
    import dataclasses

    wide = dataclasses.replace(net, arrays=[np.asarray(a).astype(np.complex128) for a in net.arrays])
    ref = complex(np.asarray(cpu_contract(wide, balanced)).ravel()[0])
    calls = {"nd": 0}
    o_nd = lib.contract_nd
    lib.contract_nd = lambda d, s: (calls.__setitem__("nd", calls["nd"] + 1), o_nd(d, s))[1]
    try:
        got_b = complex(np.asarray(T.contract_network(net, balanced)[0]).ravel()[0])
        n_before = calls["nd"]
        got_c = complex(np.asarray(T.contract_network(net, chain)[0]).ravel()[0])
        apply_launches = calls["nd"] - n_before
    finally:
        lib.contract_nd = o_nd
    big_gate_steps = sum(1 for m, n, k in steps_c if min(m, n) <= 16 and k <= 64 and (m * n > 2 ** 20 or m * n * k > 2 ** 22))
    assert apply_launches >= big_gate_steps > 100          # the chain's large steps take the apply route, one launch each
    assert abs(got_b - ref) <= 1e-5 * abs(ref), (got_b, ref)
    assert abs(got_c - ref) <= 1e-5 * abs(ref), (got_c, ref)
    # graph replay of the chain gives the same number
    again = complex(np.asarray(T.contract_network(net, chain, graph=True)[0]).ravel()[0])
    again = complex(np.asarray(T.contract_network(net, chain, graph=True)[0]).ravel()[0])
    assert abs(again - got_c) <= 1e-6 * abs(ref)


def test_contract_network_with_a_searched_tree():
    """``path="search"``: the tree (and its slicing) comes from tn.search_path; same value as along the plain greedy tree."""
    from rosnet_b200 import tn as T
    from rosnet_b200.tn import contract as C
    from rosnet_b200.tn.network import simplify

    net = simplify(T.sycamore_amplitude_network(6, seed=2, dtype="complex128"))
    want = complex(np.asarray(T.contract_network(net)[0]).ravel()[0])
    got, stats = T.contract_network(net, path="search")
    assert abs(complex(np.asarray(got).ravel()[0]) - want) <= 1e-12 * abs(want)
    plain = T.path_cost(net.inputs, net.output, net.sizes, T.greedy_path(net.inputs, net.output, net.sizes, seed=0))[0]
    assert stats["flops"] <= 8 * plain * 4      # the time model may trade a few flops for fewer passes over HBM, never orders of magnitude
    # a tight memory limit makes the search slice the tree; the sum over its slices is the same number
    old = C.SEARCH_TARGET_BYTES
    C.SEARCH_TARGET_BYTES = 16 * 2 ** 7
    try:
        got2, stats2 = T.contract_network(net, path="search")
    finally:
        C.SEARCH_TARGET_BYTES = old
    assert stats2["n_slices"] > 1 and stats2["largest"] <= 2 ** 7
    assert abs(complex(np.asarray(got2).ravel()[0]) - want) <= 1e-12 * abs(want)
    with pytest.raises(ValueError):
        T.contract_network(net, path="search", sliced=[net.inputs[0][0]])
    with pytest.raises(ValueError):
        T.contract_network(net, path="optimal")
