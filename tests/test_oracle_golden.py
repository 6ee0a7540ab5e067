"""Pin the CPU oracle against golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py; reference rosnet/array/block/__init__.py:281-336)."""
import numpy as np
import pytest

from cases import CASES, case_inputs
from oracle import blocked_tensordot as orc

TOL = {"float64": 1e-12, "complex128": 1e-12, "float32": 1e-5, "complex64": 1e-5}


@pytest.mark.parametrize("index", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_inputs_regenerate(golden, index):
    case = CASES[index]
    a, b = case_inputs(case, index)
    assert np.array_equal(a, golden[f"{case['name']}/a"])
    assert np.array_equal(b, golden[f"{case['name']}/b"])


@pytest.mark.parametrize("index", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_oracle_K_order_is_bit_exact_with_reference(golden, index):
    """order='K' restates np.nested_iters' default exactly, defects included."""
    case = CASES[index]
    n = case["name"]
    a, b = golden[f"{n}/a"], golden[f"{n}/b"]
    A = orc.split_blocks(a, case["ab"])
    B = orc.split_blocks(b, case["bb"])
    C = orc.blocked_tensordot(A, B, case["axes"], order="K")
    dense = orc.to_numpy(C)
    ref = golden[f"{n}/ref"]
    assert dense.shape == ref.shape
    assert np.array_equal(dense, ref), "restatement differs from the recorded reference output"
    assert tuple(C.shape) == tuple(golden[f"{n}/grid"])
    if C.ndim:
        shape, grid, bs, dtype = orc.grid_layout(C)
        assert bs == tuple(golden[f"{n}/blockshape"])
        assert shape == tuple(golden[f"{n}/shape"])
        assert str(dtype) == str(golden[f"{n}/dtype"])


@pytest.mark.parametrize("index", range(len(CASES)), ids=[c["name"] for c in CASES])
def test_oracle_C_order_matches_dense(golden, index):
    """order='C' (the pairing the product implements) equals dense numpy.tensordot."""
    case = CASES[index]
    n = case["name"]
    a, b = golden[f"{n}/a"], golden[f"{n}/b"]
    C = orc.blocked_tensordot(orc.split_blocks(a, case["ab"]), orc.split_blocks(b, case["bb"]), case["axes"], order="C")
    dense = golden[f"{n}/dense"]
    got = orc.to_numpy(C)
    assert got.shape == dense.shape
    assert got.dtype == dense.dtype
    assert orc.relative_frobenius(got, dense) <= TOL[str(dense.dtype)]


def test_reference_defect_instances_are_what_we_think(golden):
    """Where the recorded reference disagrees with dense numpy, it is the A1 mis-pairing:
    >= 2 contracted grid axes with extent > 1 listed in different relative order."""
    bad = []
    for case in CASES:
        n = case["name"]
        err = orc.relative_frobenius(golden[f"{n}/ref"], golden[f"{n}/dense"])
        if err > 1e-3:
            bad.append(n)
            ax_a, ax_b = case["axes"]
            rank_a = np.argsort(np.argsort(ax_a))
            rank_b = np.argsort(np.argsort(ax_b))
            assert not np.array_equal(rank_a, rank_b), n
    assert set(bad) == {"rank3_c64", "rank3_f32", "reversed_axes", "a1_mispaired"}


def test_commutative_matches_sequential_to_rounding(golden):
    case = CASES[2]
    n = case["name"]
    A = orc.split_blocks(golden[f"{n}/a"], case["ab"])
    B = orc.split_blocks(golden[f"{n}/b"], case["bb"])
    s = orc.to_numpy(orc.blocked_tensordot(A, B, case["axes"], method="sequential"))
    c = orc.to_numpy(orc.blocked_tensordot(A, B, case["axes"], method="commutative"))
    assert orc.relative_frobenius(c, s) < 1e-15


def test_flops_model_golden(golden):
    for m, n, k, f in golden["flops_matmul"]:
        assert orc.flops_tensordot((m, k), (k, n), [(1,), (0,)]) == f == m * n * k


def test_mem_model():
    assert orc.mem_tensordot((4, 8), "float64", (8, 2), "complex128", [(1,), (0,)]) == 4 * 8 * 8 + 8 * 2 * 16 + 4 * 2 * 16


def test_transpose_reshape_intent():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 6, 8))
    g = orc.split_blocks(x, (2, 3, 4))
    t = orc.blocked_transpose(g, (2, 0, 1))
    assert np.array_equal(orc.to_numpy(t), x.transpose(2, 0, 1))
    with pytest.raises(ValueError):
        orc.blocked_transpose(g, (0, 0, 1))
    r = orc.blocked_reshape(g, (6, 4))
    assert r.shape == g.shape and r[0, 0, 0].shape == (6, 4)
    assert np.array_equal(r[1, 0, 1], np.reshape(g[1, 0, 1], (6, 4), order="F"))


def test_join_idx_and_result_shape():
    assert orc.join_idx((7, 9), (1, 2), (2, 0)) == (2, 7, 1, 9)
    assert orc.result_shape((2, 3, 4), (4, 5, 3), [(1, 2), (2, 0)]) == (2, 5)
