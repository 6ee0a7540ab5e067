"""Parity of the launch shapes tensor-network paths produce: scalar-like
results over a huge contracted extent (SIMT "skinny" route), tiny problems ("small" route),
few-tile / long-k GEMMs (split-K with a fixed-order second pass) and wide-N GEMMs (column-panel
rasterisation).  Truth = numpy in double precision; tolerances as BASELINE.json states them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {"float64": 1e-12, "complex128": 1e-12, "float32": 1e-5, "complex64": 1e-5}
DTYPES = ["float64", "complex128", "float32", "complex64"]


def _rel(x, ref):
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(ref.ravel()) or 1.0))


def _rand(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def _wide(dtype):
    return "complex128" if np.dtype(dtype).kind == "c" else "float64"


def _launches():
    from rosnet_b200.engine import lib

    return dict(lib.launch_counter)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,n", [(1, 1), (2, 3), (4, 4), (3, 1)])
@pytest.mark.parametrize("layout", ["k_major", "mn_major"])
def test_skinny_long_k(dtype, m, n, layout):
    """m, n <= 4 and k = 3 x 33331 per output (ragged, three contracted blocks): the reduction
    route; float32/complex64 accumulate in double there, so they land far inside 1e-5."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(21)
    kb, nk = 33331, 3
    K = kb * nk
    if layout == "k_major":
        a, b = _rand(rng, (m, K), dtype), _rand(rng, (n, K), dtype)
        A, B = rn.array(a, blockshape=(m, kb), inner="cuda"), rn.array(b, blockshape=(n, kb), inner="cuda")
        axes = [(1,), (1,)]
    else:
        a, b = _rand(rng, (K, m), dtype), _rand(rng, (K, n), dtype)
        A, B = rn.array(a, blockshape=(kb, m), inner="cuda"), rn.array(b, blockshape=(kb, n), inner="cuda")
        axes = [(0,), (0,)]
    C = np.tensordot(A, B, axes)
    want = np.tensordot(a.astype(_wide(dtype)), b.astype(_wide(dtype)), axes)
    assert C.shape == want.shape and C.dtype == np.dtype(dtype)
    assert _rel(np.asarray(C).astype(_wide(dtype)), want) <= min(TOL[dtype], 2e-7)


@pytest.mark.parametrize("dtype", DTYPES)
def test_full_contraction_of_large_tensors(dtype):
    """The last step of a closed tensor network: <a|b> over every axis, 2^21 elements, in blocks;
    axes listed in different orders on the two sides (one operand needs its permute first)."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(22)
    shape = (128, 64, 256)
    a, b = _rand(rng, shape, dtype), _rand(rng, (256, 128, 64), dtype)
    A = rn.array(a, blockshape=(64, 64, 128), inner="cuda")
    B = rn.array(b, blockshape=(128, 64, 64), inner="cuda")
    C = np.tensordot(A, B, [(0, 1, 2), (1, 2, 0)])
    want = np.tensordot(a.astype(_wide(dtype)), b.astype(_wide(dtype)), [(0, 1, 2), (1, 2, 0)])
    assert C.shape == () and C.grid == ()
    got = complex(np.asarray(C)) if np.dtype(dtype).kind == "c" else float(np.asarray(C))
    assert abs(got - want) <= min(TOL[dtype], 2e-7) * np.sqrt(a.size) * 2


@pytest.mark.parametrize("dtype", DTYPES)
def test_split_k_few_tiles_long_k(dtype, monkeypatch):
    """256 x 192 output (2-4 CTA tiles) over k = 4 x 8200: the (pair, k) range of every tile is split
    over the SMs, splits cross contracted-block boundaries, partials are added in a fixed order.
    The result must equal the unsplit kernel's (RNB_SPLITK=0) to rounding and be deterministic."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(23)
    a, b = _rand(rng, (256, 32800), dtype), _rand(rng, (32800, 192), dtype)
    A = rn.array(a, blockshape=(128, 8200), inner="cuda")
    B = rn.array(b, blockshape=(8200, 96), inner="cuda")
    want = a.astype(_wide(dtype)) @ b.astype(_wide(dtype))
    got1 = np.asarray(np.tensordot(A, B, [(1,), (0,)]))
    got2 = np.asarray(np.tensordot(A, B, [(1,), (0,)]))
    assert np.array_equal(got1, got2)                       # fixed-order reduction: bit-reproducible
    assert _rel(got1.astype(_wide(dtype)), want) <= TOL[dtype]
    monkeypatch.setenv("RNB_SPLITK", "0")
    unsplit = np.asarray(np.tensordot(A, B, [(1,), (0,)]))
    assert _rel(unsplit.astype(_wide(dtype)), want) <= TOL[dtype]
    assert not np.array_equal(unsplit, got1)                # really a different summation order
    monkeypatch.setenv("RNB_SPLITK", "7")                   # forced odd split count
    forced = np.asarray(np.tensordot(A, B, [(1,), (0,)]))
    assert _rel(forced.astype(_wide(dtype)), want) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["float64", "float32", "complex64"])
def test_split_k_single_block_ragged_tile_accumulate(dtype):
    """One output block, ragged tile edges, split-K, through the Sequence overload whose
    commutative method accumulates into an existing result (accumulate=1 in the second pass)."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(24)
    As = [_rand(rng, (130, 20000), dtype) for _ in range(2)]
    Bs = [_rand(rng, (20000, 70), dtype) for _ in range(2)]
    want = sum(x.astype(_wide(dtype)) @ y.astype(_wide(dtype)) for x, y in zip(As, Bs))
    for method in ("sequential", "commutative"):
        got = rn.tensordot(As, Bs, [(1,), (0,)], method=method)
        assert _rel(np.asarray(got).astype(_wide(dtype)), want) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["float32", "complex64"])
def test_wide_n_panel_rasterisation(dtype):
    """More than 12 n-tiles: tiles are walked in column panels of 8 (last panel narrower); every
    output element must still come from the right (tile_m, tile_n)."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(25)
    n = (256 * 13 + 40) // (2 if dtype == "complex64" else 1)     # 14 n-tiles of 256 floats, last ragged
    a, b = _rand(rng, (300, 96), dtype), _rand(rng, (96, n), dtype)
    A = rn.array(a, blockshape=(300, 96), inner="cuda")
    B = rn.array(b, blockshape=(96, n), inner="cuda")
    got = np.asarray(np.tensordot(A, B, [(1,), (0,)]))
    want = a.astype(_wide(dtype)) @ b.astype(_wide(dtype))
    assert _rel(got.astype(_wide(dtype)), want) <= TOL[dtype]
    # two output blocks, three contracted blocks, panels inside each block
    A2 = rn.array(np.concatenate([a, a[::-1]], 0), blockshape=(300, 32), inner="cuda")
    B2 = rn.array(b, blockshape=(32, n), inner="cuda")
    got2 = np.asarray(np.tensordot(A2, B2, [(1,), (0,)]))
    assert _rel(got2[:300].astype(_wide(dtype)), want) <= TOL[dtype]
    assert _rel(got2[300:].astype(_wide(dtype)), want[::-1]) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_small_route_is_one_launch(dtype):
    """Tiny problems (the first hundreds of steps of a TN path): one SIMT launch, no split pre-pass."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(26)
    a, b = _rand(rng, (2, 2, 2, 2), dtype), _rand(rng, (2, 2, 2), dtype)
    A, B = rn.array(a, blockshape=(2, 2, 2, 2), inner="cuda"), rn.array(b, blockshape=(2, 2, 2), inner="cuda")
    before = _launches()
    C = np.tensordot(A, B, [(1, 3), (2, 0)])
    after = _launches()
    want = np.tensordot(a.astype(_wide(dtype)), b.astype(_wide(dtype)), [(1, 3), (2, 0)])
    assert _rel(np.asarray(C).astype(_wide(dtype)), want) <= min(TOL[dtype], 2e-7)
    assert after["contract"] - before["contract"] == 1


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", ["skinny", "small", "splitk", "plain"])
def test_accumulate_into_existing_result(dtype, case):
    """accumulate=1 (C += sum, the MaybeArray / commutative semantics of
    rosnet/array/compss/task/tensordot.py:45-49) on every route of rnb_contract_grouped."""
    import rosnet_b200 as rn
    from rosnet_b200.engine import plan as _plan, runtime

    rng = np.random.default_rng(27)
    m, n, k, kb = {"skinny": (2, 2, 40000, 20000), "small": (8, 6, 64, 32), "splitk": (130, 70, 16384, 8192),
                   "plain": (256, 512, 256, 128)}[case]
    a, b = _rand(rng, (m, k), dtype), _rand(rng, (n, k), dtype)
    A, B = rn.array(a, blockshape=(m, kb), inner="cuda"), rn.array(b, blockshape=(n, kb), inner="cuda")
    axes = [(1,), (1,)]
    C = np.tensordot(A, B, axes)
    first = np.asarray(C).copy()
    p = _plan.plan_tensordot(A.grid, A.blockshape, B.grid, B.blockshape, np.dtype(dtype), axes)
    runtime.execute_tensordot(p, A._storage, A.blockshape, B._storage, B.blockshape, out_storage=C._storage, accumulate=True)
    twice = np.asarray(C)
    want = np.tensordot(a.astype(_wide(dtype)), b.astype(_wide(dtype)), axes)
    assert _rel(first.astype(_wide(dtype)), want) <= TOL[dtype]
    assert _rel(twice.astype(_wide(dtype)), 2 * want) <= TOL[dtype]
