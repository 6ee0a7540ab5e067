"""Parity of the CUDA path (through np.tensordot -> __array_function__ -> C ABI) with the CPU
oracle, the recorded reference outputs and dense numpy.tensordot.

Tolerances (BASELINE.json north_star): relative Frobenius <= 1e-12 for float64/complex128,
<= 1e-5 for float32/complex64; grid, blockshape, shape, dtype and index order exactly equal."""
import numpy as np
import pytest

from cases import CASES

pytestmark = pytest.mark.gpu

TOL = {"float64": 1e-12, "complex128": 1e-12, "float32": 1e-5, "complex64": 1e-5}
F64_CASES = list(range(len(CASES)))  # all dtypes: float64/complex128 (DMMA) and float32/complex64 (3xTF32 tcgen05)


def _rel(x, ref):
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(ref.ravel()) or 1.0))


@pytest.mark.parametrize("index", F64_CASES, ids=[CASES[i]["name"] for i in F64_CASES])
@pytest.mark.parametrize("resident", ["host_blocks", "device_blocks"])
def test_golden_cases(golden, index, resident, contract_route):
    import rosnet_b200 as rn
    from oracle import blocked_tensordot as orc

    case = CASES[index]
    n = case["name"]
    a, b = golden[f"{n}/a"], golden[f"{n}/b"]
    inner = "numpy" if resident == "host_blocks" else "cuda"
    A = rn.array(a, blockshape=case["ab"], inner=inner)
    B = rn.array(b, blockshape=case["bb"], inner=inner)
    C = np.tensordot(A, B, case["axes"])
    assert isinstance(C, rn.BlockArray) and C.is_device
    dense = golden[f"{n}/dense"]
    got = np.asarray(C)
    assert got.shape == dense.shape and got.dtype == dense.dtype
    assert _rel(got, dense) <= TOL[str(dense.dtype)]
    # layout exactly as the reference reports it
    assert C.grid == tuple(golden[f"{n}/grid"])
    assert C.blockshape == tuple(golden[f"{n}/blockshape"])
    assert C.shape == tuple(golden[f"{n}/shape"])
    # block-by-block against the C-order oracle (same block in the same grid cell)
    want = orc.blocked_tensordot(orc.split_blocks(a, case["ab"]), orc.split_blocks(b, case["bb"]), case["axes"], order="C")
    H = C.to_host()
    assert H.grid == want.shape
    for idx in np.ndindex(*want.shape):
        assert _rel(H.data[idx], np.asarray(want[idx])) <= TOL[str(dense.dtype)]
    # where the recorded reference is right (no A1 mis-pairing) we match it too
    ref = golden[f"{n}/ref"]
    if _rel(ref, dense) < 1e-9:
        assert _rel(got, ref) <= TOL[str(dense.dtype)]


@pytest.mark.parametrize("dtype", ["float64", "complex128", "float32", "complex64"])
@pytest.mark.parametrize("shape", [
    # (a_shape, a_block, b_shape, b_block, axes): ragged tiles, all four operand storage orders
    ((200, 72), (100, 24), (72, 136), (24, 68), [(1,), (0,)]),        # K-major x MN-major
    ((200, 72), (100, 24), (136, 72), (68, 24), [(1,), (1,)]),        # K-major x K-major
    ((72, 200), (24, 40), (72, 136), (24, 136), [(0,), (0,)]),        # MN-major x MN-major
    ((72, 200), (24, 40), (136, 72), (136, 24), [(0,), (1,)]),        # MN-major x K-major
    ((130, 20), (130, 20), (20, 70), (20, 70), [(1,), (0,)]),         # one block, odd tiles
    ((8, 6, 10), (4, 3, 5), (10, 6, 12), (5, 3, 4), [(1, 2), (1, 0)]),  # needs a permute of b
    ((6, 3), (3, 3), (3, 10), (3, 5), [(1,), (0,)]),                  # k % 4 != 0 -> padded workspace
])
def test_shapes_and_layouts(dtype, shape, contract_route):
    import rosnet_b200 as rn

    sa, ba, sb, bb, axes = shape
    rng = np.random.default_rng(7)
    a = rng.standard_normal(sa)
    b = rng.standard_normal(sb)
    if dtype.startswith("complex"):
        a = a + 1j * rng.standard_normal(sa)
        b = b + 1j * rng.standard_normal(sb)
    a, b = a.astype(dtype), b.astype(dtype)
    C = np.tensordot(rn.array(a, blockshape=ba, inner="cuda"), rn.array(b, blockshape=bb, inner="cuda"), axes)
    assert C.dtype == np.dtype(dtype)
    wide = "complex128" if dtype.startswith("complex") else "float64"
    want = np.tensordot(a.astype(wide), b.astype(wide), axes)   # truth in double precision
    assert _rel(np.asarray(C).astype(wide), want) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["float32", "complex64"])
def test_3xtf32_accuracy_at_depth(dtype):
    """k = 2 x 2048 contracted elements per output: the compensated product must stay at the
    reference's own single-precision noise floor (~1e-7), far inside the 1e-5 bar; a plain TF32
    product (RNB_PREC_TF32X1) must NOT (it is what the compensation is for)."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(11)
    m, n, k = 384, 320, 4096
    a = rng.standard_normal((m, k))
    b = rng.standard_normal((k, n))
    if dtype == "complex64":
        a = a + 1j * rng.standard_normal((m, k))
        b = b + 1j * rng.standard_normal((k, n))
    a, b = a.astype(dtype), b.astype(dtype)
    wide = "complex128" if dtype == "complex64" else "float64"
    truth = a.astype(wide) @ b.astype(wide)
    A = rn.array(a, blockshape=(192, 2048), inner="cuda")
    B = rn.array(b, blockshape=(2048, 160), inner="cuda")
    err3 = _rel(np.asarray(np.tensordot(A, B, [(1,), (0,)])).astype(wide), truth)
    with rn.tuning.configure(precision="tf32x1"):
        err1 = _rel(np.asarray(np.tensordot(A, B, [(1,), (0,)])).astype(wide), truth)
    with rn.tuning.configure(precision="tf32_bf16x2"):
        err2 = _rel(np.asarray(np.tensordot(A, B, [(1,), (0,)])).astype(wide), truth)
    ref_err = _rel((a @ b).astype(wide), truth)   # the reference's own arithmetic (numpy sgemm/cgemm)
    print(f"{dtype}: 3xTF32 {err3:.2e}  TF32+2xBF16 {err2:.2e}  1xTF32 {err1:.2e}  numpy {ref_err:.2e}")
    assert err3 <= 2e-6
    assert err2 <= 5e-6     # opt-in mode: cross terms in BF16, still inside the 1e-5 bar
    assert err1 > 1e-5


def test_c1_full_size_both_axes_variants():
    """BASELINE configs[0] at full size: (64,)^4 float64, blockshape (32,)^4, 2 contracted axes."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(0)
    a = rng.standard_normal((64,) * 4)
    b = np.random.default_rng(1).standard_normal((64,) * 4)
    A = rn.array(a, blockshape=(32,) * 4, inner="cuda")
    B = rn.array(b, blockshape=(32,) * 4, inner="cuda")
    for axes in ([(2, 3), (0, 1)], [(0, 2), (1, 3)]):
        C = np.tensordot(A, B, axes)
        assert C.grid == (2,) * 4 and C.blockshape == (32,) * 4 and C.dtype == np.float64
        want = np.tensordot(a, b, axes)
        assert _rel(C, want) <= 1e-12


def test_linearity_and_scalar_and_accumulate_properties(contract_route):
    import rosnet_b200 as rn

    rng = np.random.default_rng(3)
    a1, a2 = rng.standard_normal((64, 48)), rng.standard_normal((64, 48))
    b = rng.standard_normal((48, 40))
    f = lambda x: np.asarray(np.tensordot(rn.array(x, blockshape=(32, 16), inner="cuda"), rn.array(b, blockshape=(16, 20), inner="cuda"), [(1,), (0,)]))
    lhs = f(a1 + 2.0 * a2)
    rhs = f(a1) + 2.0 * f(a2)
    assert _rel(lhs, rhs) < 1e-13
    # full contraction -> 0-d result, grid ()
    s = np.tensordot(rn.array(a1, blockshape=(32, 16)), rn.array(a2, blockshape=(32, 16)), [(0, 1), (0, 1)])
    assert s.grid == () and s.shape == ()
    assert abs(float(np.asarray(s)) - float((a1 * a2).sum())) <= 1e-12 * abs(float((a1 * a2).sum())) + 1e-12


def test_mixed_dtype_promotes_like_numpy():
    import rosnet_b200 as rn

    rng = np.random.default_rng(4)
    a = rng.standard_normal((16, 8))
    b = rng.standard_normal((8, 12)) + 1j * rng.standard_normal((8, 12))
    C = np.tensordot(rn.array(a, blockshape=(8, 4)), rn.array(b, blockshape=(4, 6), inner="cuda"), [(1,), (0,)])
    assert C.dtype == np.complex128
    assert _rel(C, a @ b) < 1e-13


def test_sequence_overload_and_methods(contract_route):
    import rosnet_b200 as rn

    rng = np.random.default_rng(5)
    As = [rng.standard_normal((8, 4)) for _ in range(3)]
    Bs = [rng.standard_normal((4, 6)) for _ in range(3)]
    want = sum(x @ y for x, y in zip(As, Bs))
    for method in ("sequential", "commutative", "commutative-but-first"):
        got = rn.tensordot(As, Bs, [(1,), (0,)], method=method)
        assert _rel(got, want) < 1e-13
    with pytest.raises(ValueError):
        rn.tensordot(As, Bs, [(1,), (0,)], method="nope")


def test_transpose_reshape_tonumpy_on_device():
    import rosnet_b200 as rn
    from oracle import blocked_tensordot as orc

    rng = np.random.default_rng(6)
    x = rng.standard_normal((8, 12, 10)) + 1j * rng.standard_normal((8, 12, 10))
    A = rn.array(x, blockshape=(4, 3, 5), inner="cuda")
    assert np.array_equal(np.asarray(A), x)                      # blockify + unblockify are exact
    T = np.transpose(A, (2, 0, 1))                               # in place, returns A (reference semantics)
    assert T is A and A.grid == (2, 2, 4) and A.blockshape == (5, 4, 3)
    assert np.array_equal(np.asarray(A), x.transpose(2, 0, 1))
    with pytest.raises(ValueError):
        np.transpose(A, (0, 0, 1))
    B = rn.array(x, blockshape=(4, 3, 5))
    R = rn.reshape(B, (6, 2, 5), inplace=False)                  # per-block, order="F" default
    want = orc.blocked_reshape(orc.split_blocks(x, (4, 3, 5)), (6, 2, 5), order="F")
    H = R.to_host()
    for idx in np.ndindex(*want.shape):
        assert np.array_equal(H.data[idx], want[idx])
    Rc = rn.reshape(B, (6, 2, 5), order="C", inplace=False).to_host()
    assert np.array_equal(Rc.data[1, 2, 1], x[4:8, 6:9, 5:10].reshape(6, 2, 5))


def test_pipelined_host_operand_path():
    """Host-block operands above the size threshold take the streamed path (3 CUDA streams);
    result, layout and the asynchronously filled host mirror must match the simple path."""
    import rosnet_b200 as rn
    from rosnet_b200.engine import pipeline

    rng = np.random.default_rng(8)
    a = rng.standard_normal((4 * 64, 2 * 96))
    b = rng.standard_normal((2 * 96, 6 * 80))
    A = rn.array(a, blockshape=(64, 96))       # pageable host blocks (non-contiguous views copied)
    B = rn.array(b, blockshape=(96, 80))
    old = pipeline.MIN_BYTES
    pipeline.MIN_BYTES = 0
    try:
        C = np.tensordot(A, B, [(1,), (0,)])
        assert C._host_mirror is not None
        H = C.to_host()
    finally:
        pipeline.MIN_BYTES = old
    want = a @ b
    assert H.grid == (4, 6) and H.blockshape == (64, 80)
    assert _rel(np.asarray(H), want) <= 1e-12          # assembled from the streamed host mirror
    assert _rel(np.asarray(C), want) <= 1e-12          # and from the device copy
    # a layout change drops the mirror
    np.transpose(C, (1, 0))
    assert C._host_mirror is None and _rel(np.asarray(C), want.T) <= 1e-12
    # complex64 through the same path
    a = (rng.standard_normal((128, 64)) + 1j * rng.standard_normal((128, 64))).astype(np.complex64)
    b = (rng.standard_normal((96, 64)) + 1j * rng.standard_normal((96, 64))).astype(np.complex64)
    pipeline.MIN_BYTES = 0
    try:
        C = np.tensordot(rn.array(a, blockshape=(32, 64)), rn.array(b, blockshape=(32, 64)), [(1,), (1,)])
        H = C.to_host()
    finally:
        pipeline.MIN_BYTES = old
    assert _rel(np.asarray(H).astype(np.complex128), a.astype(np.complex128) @ b.astype(np.complex128).T) <= 1e-5


def test_complex128_3m_mode(monkeypatch):
    """RNB_CPLX_3M: three real DMMA products per complex tile product; same tolerance."""
    import rosnet_b200 as rn

    monkeypatch.setenv("RNB_SIMT", "0")   # 3M lives in the DMMA kernel
    rng = np.random.default_rng(13)
    for (sa, ba, sb, bb, axes) in [((200, 72), (100, 24), (72, 136), (24, 68), [(1,), (0,)]),
                                   ((72, 200), (24, 40), (136, 72), (136, 24), [(0,), (1,)]),
                                   ((256, 512), (128, 256), (512, 96), (256, 32), [(1,), (0,)])]:
        a = rng.standard_normal(sa) + 1j * rng.standard_normal(sa)
        b = rng.standard_normal(sb) + 1j * rng.standard_normal(sb)
        A, B = rn.array(a, blockshape=ba, inner="cuda"), rn.array(b, blockshape=bb, inner="cuda")
        with rn.tuning.configure(complex_mode="3m"):
            C3 = np.asarray(np.tensordot(A, B, axes))
        with rn.tuning.configure(complex_mode="4m"):
            C4 = np.asarray(np.tensordot(A, B, axes))
        Cauto = np.asarray(np.tensordot(A, B, axes))
        want = np.tensordot(a, b, axes)
        assert _rel(C3, want) <= 1e-12 and _rel(C4, want) <= 1e-12
        assert not np.array_equal(C3, C4)   # really a different arithmetic path
        # RNB_CPLX_AUTO: complex128 takes 3M from a summed contracted extent of 64 up (csrc/gemm_f64.cu: f64_mode)
        k_total = int(np.prod([sa[i] for i in axes[0]]))
        assert np.array_equal(Cauto, C3 if k_total >= 64 else C4)
    # below the threshold AUTO stays 4M
    a = rng.standard_normal((160, 48)) + 1j * rng.standard_normal((160, 48))
    b = rng.standard_normal((48, 130)) + 1j * rng.standard_normal((48, 130))
    A, B = rn.array(a, blockshape=(160, 48), inner="cuda"), rn.array(b, blockshape=(48, 130), inner="cuda")
    with rn.tuning.configure(complex_mode="4m"):
        C4 = np.asarray(np.tensordot(A, B, [(1,), (0,)]))
    assert np.array_equal(np.asarray(np.tensordot(A, B, [(1,), (0,)])), C4)


@pytest.mark.parametrize("dtype,mode", [("float32", "auto"), ("complex64", "4m"), ("complex64", "3m")])
def test_fused_matricize_split_is_bit_identical_to_the_two_pass_route(dtype, mode, monkeypatch):
    """float32 / complex64 operands that are not in GEMM layout: rnb_permute_split (matricize fused with the hi/lo TF32
    split, every mode: same-offset planes, complex b of 4M, 3M planes) against rnb_permute + the split pre-pass -
    the same planes, so the same bits - on multi-block operands whose contracted axes are scattered."""
    import rosnet_b200 as rn
    from rosnet_b200.engine import lib

    monkeypatch.setenv("RNB_SIMT", "0")
    rng = np.random.default_rng(21)
    a = rng.standard_normal((8, 64, 4, 32, 4))
    b = rng.standard_normal((32, 4, 48, 8, 4, 2))
    if dtype == "complex64":
        a = a + 1j * rng.standard_normal(a.shape)
        b = b + 1j * rng.standard_normal(b.shape)
    a, b = a.astype(dtype), b.astype(dtype)
    A = rn.array(a, blockshape=(4, 64, 4, 16, 4), inner="cuda")
    B = rn.array(b, blockshape=(16, 4, 48, 4, 4, 2), inner="cuda")
    axes = [(0, 3, 2), (3, 0, 4)]
    with rn.tuning.configure(complex_mode=mode):
        before = dict(lib.launch_counter)
        fused = np.asarray(np.tensordot(A, B, axes))
        n_fused = lib.launch_counter["permute"] - before["permute"]
        monkeypatch.setenv("RNB_FUSE_SPLIT", "0")
        two_pass = np.asarray(np.tensordot(A, B, axes))
    assert n_fused == 2                      # one fused launch per operand, no separate matricize
    assert np.array_equal(fused, two_pass)
    wide = np.complex128 if dtype == "complex64" else np.float64
    assert _rel(fused, np.tensordot(a.astype(wide), b.astype(wide), axes)) <= 1e-5
