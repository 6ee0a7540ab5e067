"""Host planning of the blocked contraction, checked on the CPU by interpreting the plan with
NumPy (tests/plan_interp.py) against the recorded reference outputs and dense numpy.tensordot."""
from math import prod

import numpy as np
import pytest

from cases import CASES
from oracle import blocked_tensordot as orc
from rosnet_b200.engine import lib
from rosnet_b200.engine.plan import plan_tensordot
import plan_interp as pi

TOL = {"float64": 1e-12, "complex128": 1e-12, "float32": 1e-5, "complex64": 1e-5}


def run_case(a, ab, b, bb, axes):
    dtype = np.result_type(a, b)
    a, b = a.astype(dtype), b.astype(dtype)
    fa, ga = pi.block_major(a, ab)
    fb, gb = pi.block_major(b, bb)
    plan = plan_tensordot(ga, ab, gb, bb, dtype, axes)
    out = pi.run_plan(plan, fa, ab, fb, bb)
    return plan, pi.from_block_major(out, plan.out_grid, plan.out_blockshape)


@pytest.mark.parametrize("index", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_plan_matches_dense_and_oracle_layout(golden, index):
    case = CASES[index]
    n = case["name"]
    a, b = golden[f"{n}/a"], golden[f"{n}/b"]
    plan, got = run_case(a, case["ab"], b, case["bb"], case["axes"])
    dense = golden[f"{n}/dense"]
    assert got.shape == dense.shape
    assert orc.relative_frobenius(got, dense) <= TOL[str(dense.dtype)]
    # block layout and index ordering exactly as the reference reports them
    assert plan.out_grid == tuple(golden[f"{n}/grid"])
    assert plan.out_blockshape == tuple(golden[f"{n}/blockshape"])
    want = orc.blocked_tensordot(orc.split_blocks(a, case["ab"]), orc.split_blocks(b, case["bb"]), case["axes"], order="C")
    assert plan.out_grid == want.shape
    # per-block agreement with the C-order oracle (same block, same position)
    blk = prod(plan.out_blockshape)
    for i, idx in enumerate(np.ndindex(*want.shape)):
        sl = tuple(slice(j * s, (j + 1) * s) for j, s in zip(idx, plan.out_blockshape))
        assert orc.relative_frobenius(got[sl], np.asarray(want[idx])) <= TOL[str(dense.dtype)]


def test_c1_plan_is_one_launch_no_permute():
    p = plan_tensordot((2,) * 4, (32,) * 4, (2,) * 4, (32,) * 4, "float64", [(2, 3), (0, 1)])
    assert p.a.direct and p.a.major == lib.RNB_MAJOR_K
    assert p.b.direct and p.b.major == lib.RNB_MAJOR_MN
    assert (p.m, p.n, p.k, p.n_out, p.n_pairs) == (1024, 1024, 1024, 16, 4)
    assert p.flops == 2 * 4096**3
    p2 = plan_tensordot((2,) * 4, (32,) * 4, (2,) * 4, (32,) * 4, "float64", [(0, 2), (1, 3)])
    assert not p2.a.direct and not p2.b.direct and p2.a.perm == (1, 3, 0, 2) and p2.b.perm == (0, 2, 1, 3)


def test_pair_order_is_chosen_to_avoid_permutes():
    # axes listed in an order that would need a permute of a; re-ordering the pairs avoids it
    p = plan_tensordot((1,) * 3, (8, 4, 8), (1,) * 3, (4, 8, 16), "float64", [(2, 1), (1, 0)])
    assert p.ax_a == (1, 2) and p.ax_b == (0, 1)
    assert p.a.direct and p.b.direct


def test_padding_when_k_not_multiple_of_4():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((6, 3))
    b = rng.standard_normal((3, 10))
    plan, got = run_case(a, (3, 3), b, (3, 5), [(1,), (0,)])
    assert not plan.a.direct and plan.a.zero_fill and plan.a.k == 4 and plan.b.k == 4
    assert orc.relative_frobenius(got, a @ b) < 1e-14


def test_outer_product_and_scalar():
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal((4, 2)), rng.standard_normal((6,))
    plan, got = run_case(a, (2, 2), b, (3,), [(), ()])
    assert plan.n_pairs == 1 and plan.k == 4
    assert orc.relative_frobenius(got, np.tensordot(a, b, 0)) < 1e-14
    plan, got = run_case(a, (2, 1), a, (2, 1), [(0, 1), (0, 1)])
    assert plan.out_grid == () and plan.out_blockshape == () and plan.n_out == 1
    assert abs(got - (a * a).sum()) < 1e-12


def test_errors():
    with pytest.raises(ValueError):
        plan_tensordot((2, 2), (4, 4), (2, 2), (5, 4), "float64", [(1,), (0,)])  # block extent mismatch
    with pytest.raises(ValueError):
        plan_tensordot((2, 2), (4, 4), (3, 2), (4, 4), "float64", [(1,), (0,)])  # grid extent mismatch
    with pytest.raises(NotImplementedError):
        plan_tensordot((1,), (4,), (1,), (4,), "int64", [(0,), (0,)])


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_property(seed):
    """Random ranks, grids, block extents and axis orders: plan == dense numpy.tensordot."""
    rng = np.random.default_rng(100 + seed)
    nd_a, nd_b = rng.integers(1, 5), rng.integers(1, 5)
    nc = int(rng.integers(0, min(nd_a, nd_b) + 1))
    ax_a = list(rng.permutation(nd_a)[:nc])
    ax_b = list(rng.permutation(nd_b)[:nc])
    ga = rng.integers(1, 3, nd_a); ba = rng.integers(1, 5, nd_a)
    gb = rng.integers(1, 3, nd_b); bb = rng.integers(1, 5, nd_b)
    for i, j in zip(ax_a, ax_b):
        gb[j], bb[j] = ga[i], ba[i]
    dtype = [np.float64, np.complex128, np.float32, np.complex64][seed % 4]
    a = rng.standard_normal(tuple(ga * ba)).astype(dtype)
    b = rng.standard_normal(tuple(gb * bb)).astype(dtype)
    plan, got = run_case(a, tuple(int(x) for x in ba), b, tuple(int(x) for x in bb), [ax_a, ax_b])
    want = np.tensordot(a, b, [ax_a, ax_b])
    assert got.shape == want.shape
    assert orc.relative_frobenius(got, want) <= TOL[str(np.dtype(dtype))]


def test_permute_desc_merges_and_caches():
    """Host side of rnb_permute: extent-1 axes dropped, axes adjacent in BOTH layouts merged, descriptor cached."""
    from rosnet_b200.engine import lib

    # (2, 3, 4) C-order -> same layout: everything merges into one axis of 24
    d = lib.permute_desc(8, [2, 3, 4], [12, 4, 1], [12, 4, 1])
    assert d.rank == 1 and d.extent[0] == 24 and d.src_stride[0] == 1 and d.dst_stride[0] == 1
    # transpose of the first two axes: the last axis stays the inner run, nothing else merges
    d2 = lib.permute_desc(8, [2, 3, 4], [12, 4, 1], [4, 8, 1])
    assert d2.rank == 3 and sorted(d2.extent[i] for i in range(3)) == [2, 3, 4]
    # extent-1 axes vanish; 30 binary axes moved as two groups merge down to two axes
    ext = [2] * 30
    src = [2 ** (29 - i) for i in range(30)]
    dst = [2 ** (14 - i) for i in range(15)] + [2 ** (29 - i) for i in range(15)]     # swap the two halves
    d3 = lib.permute_desc(8, ext + [1], src + [7], dst + [9])
    assert d3.rank == 2 and {d3.extent[0], d3.extent[1]} == {2 ** 15}
    assert lib.permute_desc(8, ext + [1], src + [7], dst + [9]) is d3                  # cached by shape
    with pytest.raises(lib.RosnetB200Error):
        lib.permute_desc(8, [2] * 40, [3 ** i for i in range(40)], [5 ** i for i in range(40)])


def test_plan_fast_cache_returns_the_same_plan():
    from rosnet_b200.engine.plan import plan_tensordot

    args = ((1, 1, 1), (4, 6, 8), (1, 1), (8, 6), np.dtype("complex64"), ([2, 1], [0, 1]))
    p1 = plan_tensordot(*args)
    p2 = plan_tensordot(*args)
    assert p1 is p2 and p1.m == 4 and p1.n == 1 and p1.k == 48


def test_slice_stager_layout_is_aligned():
    from rosnet_b200.tn.contract import _SliceStager

    arrays = [np.zeros((2, 2), np.complex64), np.zeros((), np.complex64), np.zeros((3, 5, 7), np.complex64)]
    offsets, total = _SliceStager.layout(arrays)
    assert offsets == [0, 256, 512] and total == 512 + 1024
    assert all(o % _SliceStager.ALIGN == 0 for o in offsets)


def _nd_offsets(ext, stride):
    """row-major enumeration (last axis fastest) of the offsets an (extent, stride) list addresses"""
    off = np.zeros((), dtype=np.int64)
    for e, s in zip(ext, stride):
        off = np.add.outer(off, np.arange(e, dtype=np.int64) * s)
    return np.asarray(off).reshape(-1)


@pytest.mark.parametrize("dtype", ["float64", "complex64"])
def test_nd_lists_of_small_contractions_address_the_blocks_in_place(dtype):
    """plan.nd (the axis lists rnb_contract_nd consumes) interpreted with NumPy equals dense tensordot: tensor-network
    style operands (many short axes, contracted axes scattered, different relative order in a and b, two blocks along a
    contracted axis)."""
    rng = np.random.default_rng(17)
    a = rng.standard_normal((2, 3, 4, 2, 2, 2)).astype(dtype)
    b = rng.standard_normal((2, 2, 5, 4, 2)).astype(dtype)
    if np.dtype(dtype).kind == "c":
        a = a + 1j * rng.standard_normal(a.shape).astype("float32")
        b = b + 1j * rng.standard_normal(b.shape).astype("float32")
    ab, bb = (2, 3, 2, 2, 2, 1), (2, 1, 5, 2, 2)
    axes = [(4, 2, 0, 5), (0, 3, 4, 1)]
    fa_, ga = pi.block_major(a, ab)
    fb_, gb = pi.block_major(b, bb)
    plan = plan_tensordot(ga, ab, gb, bb, dtype, axes)
    assert plan.nd is not None and not (plan.a.direct and plan.b.direct)
    ext_m, sa_m, ext_n, sb_n, ext_k, sa_k, sb_k = plan.nd
    om, on = _nd_offsets(ext_m, sa_m), _nd_offsets(ext_n, sb_n)
    ka, kb = _nd_offsets(ext_k, sa_k), _nd_offsets(ext_k, sb_k)
    assert om.size == plan.m and on.size == plan.n
    A = np.asarray(fa_).reshape(prod(ga), -1)
    B = np.asarray(fb_).reshape(prod(gb), -1)
    out = np.zeros((plan.n_out, plan.m, plan.n), dtype=np.result_type(dtype, np.float64))
    for g in range(plan.n_out):
        for pr in range(plan.n_pairs):
            Ab, Bb = A[plan.pair_a[g, pr]], B[plan.pair_b[g, pr]]
            out[g] += Ab[om[:, None] + ka[None, :]].astype(out.dtype) @ Bb[on[:, None] + kb[None, :]].astype(out.dtype).T
    got = pi.from_block_major(out.reshape(-1), plan.out_grid, plan.out_blockshape)
    want = np.tensordot(a.astype(out.dtype), b.astype(out.dtype), axes)
    assert orc.relative_frobenius(got, want) <= 1e-12
    # large contractions keep the GEMM route
    big = plan_tensordot((1, 1), (1024, 1024), (1, 1), (1024, 1024), "float64", [(0,), (0,)])
    assert big.nd is None


def test_apply_plans_gate_on_a_large_tensor():
    """plan.nd_apply (one tiny operand, csrc/skinny.cu: apply_nd_kernel): which contractions take the route - the host mirror
    of `apply_eligible` - and that the axis lists of such a plan, interpreted with NumPy, equal dense tensordot (the kernel
    walks the SAME lists: free axes of the large operand in output order, k offsets of both operands)."""
    from rosnet_b200.engine.plan import plan_tensordot

    nq = 21
    # a two-qubit gate [out0, out1, in0, in1] applied to qubits (5, 13) of a 21-qubit state: m = 4, k = 4, n = 2^19
    p = plan_tensordot((1,) * 4, (2,) * 4, (1,) * nq, (2,) * nq, "complex64", [(2, 3), (5, 13)])
    assert p.nd is not None and p.nd_apply and (p.m, p.n) == (4, 2 ** 19)
    ext_m, sa_m, ext_n, sb_n, ext_k, sa_k, sb_k = p.nd
    assert prod(ext_n) == 2 ** 19 and len(ext_n) == 3          # the state's free axes merge into three runs around the two gate qubits
    # the same with the operands swapped: the tiny operand is b
    q = plan_tensordot((1,) * nq, (2,) * nq, (1,) * 4, (2,) * 4, "complex64", [(5, 13), (2, 3)])
    assert q.nd_apply and (q.m, q.n) == (2 ** 19, 4)
    # small contractions stay on the tabulated one-thread-per-output kernel, GEMM-sized ones on the tensor cores
    small = plan_tensordot((1,) * 4, (2,) * 4, (1,) * 10, (2,) * 10, "complex64", [(2, 3), (5, 7)])
    assert small.nd is not None and not small.nd_apply
    for shape_s, axes_s in (((32, 64), (1,)),          # 32 rows: not tiny
                            ((4, 128), (1,))):         # k = 128 > 64
        big = plan_tensordot((1, 1), shape_s, (1, 1), (shape_s[1], 2 ** 16), "complex64", [axes_s, (0,)])
        assert big.nd is None and not big.nd_apply
    # a contracted block axis: the tiles of all pairs must fit the kernel's shared-memory tile (n_pairs * pad(ns) * k <= 1024)
    ok = plan_tensordot((1, 16), (4, 16), (16, 1), (16, 2 ** 14), "float64", [(1,), (0,)])
    assert ok.nd_apply and ok.n_pairs == 16
    too_many = plan_tensordot((1, 32), (4, 16), (32, 1), (16, 2 ** 14), "float64", [(1,), (0,)])
    assert too_many.nd is None

    # NumPy interpretation of an apply plan, ragged extents, contracted axes scattered and listed in different orders
    rng = np.random.default_rng(23)
    a = rng.standard_normal((3, 6, 5)) + 1j * rng.standard_normal((3, 6, 5))          # gate: free (3,), contracted (6, 5)
    b = rng.standard_normal((7, 5, 9, 6, 11, 13, 8)) + 1j * rng.standard_normal((7, 5, 9, 6, 11, 13, 8))
    axes = [(1, 2), (3, 1)]
    plan = plan_tensordot((1,) * 3, a.shape, (1,) * 7, b.shape, "complex128", axes)
    assert plan.nd_apply and plan.m == 3 and plan.n == 7 * 9 * 11 * 13 * 8
    ext_m, sa_m, ext_n, sb_n, ext_k, sa_k, sb_k = plan.nd
    om, on = _nd_offsets(ext_m, sa_m), _nd_offsets(ext_n, sb_n)
    ka, kb = _nd_offsets(ext_k, sa_k), _nd_offsets(ext_k, sb_k)
    got = (a.reshape(-1)[om[:, None] + ka[None, :]] @ b.reshape(-1)[on[:, None] + kb[None, :]].T).reshape(3, 7, 9, 11, 13, 8)
    assert orc.relative_frobenius(got, np.tensordot(a, b, axes)) <= 1e-13
