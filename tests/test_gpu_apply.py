"""The "apply" route of rnb_contract_nd (csrc/skinny.cu: apply_nd_kernel): contractions with one tiny operand (a gate of
<= 16 rows, k <= 64) and one large operand left in its n-d layout, through the public route
(np.tensordot on BlockArrays -> plan.nd_apply -> rnb_contract_nd), against numpy.tensordot on the same inputs.
Replaces, for these shapes, the same reference lines as every other contraction route
(rosnet/array/block/__init__.py:281-336).  Tolerances: <= 1e-12 float64 / complex128, <= 1e-5 float32 / complex64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {"float64": 1e-12, "complex128": 1e-12, "float32": 1e-5, "complex64": 1e-5}
DTYPES = ["float64", "complex128", "float32", "complex64"]


def _rel(x, ref):
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(ref.ravel()) or 1.0))


def _rand(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if dtype.startswith("complex"):
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def _spy(monkeypatch):
    """Counts what the runtime launches: (rnb_contract_nd calls, rnb_contract_grouped calls, permute launches)."""
    from rosnet_b200.engine import lib

    seen = {"nd": 0, "grouped": 0, "permute": 0}
    o_nd, o_gr, o_pm, o_ps = lib.contract_nd, lib.contract_grouped, lib.permute_launch, lib.permute_split
    monkeypatch.setattr(lib, "contract_nd", lambda d, s: (seen.__setitem__("nd", seen["nd"] + 1), o_nd(d, s))[1])
    monkeypatch.setattr(lib, "contract_grouped", lambda d, s: (seen.__setitem__("grouped", seen["grouped"] + 1), o_gr(d, s))[1])
    monkeypatch.setattr(lib, "permute_launch", lambda *a: (seen.__setitem__("permute", seen["permute"] + 1), o_pm(*a))[1])
    monkeypatch.setattr(lib, "permute_split", lambda *a: (seen.__setitem__("permute", seen["permute"] + 1), o_ps(*a))[1])
    return seen


# gate on qubits (p, q) of a 20-qubit state: every position of the contracted axes inside the big operand, both operand orders
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("qubits", [(0, 1), (18, 19), (0, 19), (7, 3), (19, 10)])
@pytest.mark.parametrize("gate_first", [True, False])
def test_two_qubit_gate_on_a_state_tensor(dtype, qubits, gate_first, monkeypatch):
    import rosnet_b200 as rn

    rng = np.random.default_rng(hash((dtype, qubits, gate_first)) % 2 ** 31)
    nq = 20
    state = _rand(rng, (2,) * nq, dtype)
    gate = _rand(rng, (2, 2, 2, 2), dtype)            # [out0, out1, in0, in1]
    S = rn.array(state, blockshape=state.shape, inner="cuda")
    G = rn.array(gate, blockshape=gate.shape, inner="cuda")
    seen = _spy(monkeypatch)
    if gate_first:
        got = np.tensordot(G, S, axes=[(2, 3), qubits])
        ref = np.tensordot(gate, state, axes=[(2, 3), qubits])
    else:
        got = np.tensordot(S, G, axes=[qubits, (2, 3)])
        ref = np.tensordot(state, gate, axes=[qubits, (2, 3)])
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert seen == {"nd": 1, "grouped": 0, "permute": 0}, seen       # one launch, nothing matricized
    assert _rel(got, ref) <= TOL[dtype]


# ragged extents (nothing a power of two), small extents 1..16, k up to 64, contracted axes scattered and in swapped order
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [
    # big shape, big contracted axes, small free extents
    ((3, 5, 7, 11, 13, 6, 9), (1, 4), (5,)),
    ((3, 5, 7, 11, 13, 6, 9), (6, 0), (16,)),
    ((33, 4, 201, 8, 7), (1, 3), (3, 2)),
    ((33, 4, 201, 8, 7), (3, 1, 4), (1,)),
    ((1200, 64, 5), (1,), (9,)),
    ((5, 700, 2, 300), (2,), (2, 2, 3)),
])
@pytest.mark.parametrize("small_first", [True, False])
def test_ragged_shapes_any_axis_positions(dtype, case, small_first, monkeypatch):
    import rosnet_b200 as rn

    big_shape, k_axes, small_free = case
    rng = np.random.default_rng(hash((dtype, case, small_first)) % 2 ** 31)
    big = _rand(rng, big_shape, dtype)
    k_ext = tuple(big_shape[x] for x in k_axes)
    # the small operand lists its contracted axes in a different relative order than the big one
    order = list(rng.permutation(len(k_axes)))
    small = _rand(rng, tuple(small_free) + tuple(k_ext[o] for o in order), dtype)
    small_k = tuple(len(small_free) + order.index(i) for i in range(len(k_axes)))
    B = rn.array(big, blockshape=big.shape, inner="cuda")
    Sm = rn.array(small, blockshape=small.shape, inner="cuda")
    seen = _spy(monkeypatch)
    if small_first:
        got, ref = np.tensordot(Sm, B, axes=[small_k, k_axes]), np.tensordot(small, big, axes=[small_k, k_axes])
    else:
        got, ref = np.tensordot(B, Sm, axes=[k_axes, small_k]), np.tensordot(big, small, axes=[k_axes, small_k])
    assert got.shape == ref.shape
    assert seen["nd"] == 1 and seen["grouped"] == 0 and seen["permute"] == 0, seen
    assert _rel(got, ref) <= TOL[dtype]


# blocked operands: several output blocks (blockIdx.y) and a contracted block axis summed inside the kernel (n_pairs > 1)
@pytest.mark.parametrize("dtype", DTYPES)
def test_blocked_operands_block_sum_inside_the_kernel(dtype, monkeypatch):
    import rosnet_b200 as rn
    from oracle import blocked_tensordot as orc

    rng = np.random.default_rng(11)
    big = _rand(rng, (6, 8, 4096, 10), dtype)          # blocks (3, 4, 1024, 5): grid (2, 2, 4, 2)
    small = _rand(rng, (8, 10, 4), dtype)              # blocks (4, 5, 4): grid (2, 2, 1)
    B = rn.array(big, blockshape=(3, 4, 1024, 5), inner="cuda")
    Sm = rn.array(small, blockshape=(4, 5, 4), inner="cuda")
    seen = _spy(monkeypatch)
    axes = [(1, 3), (0, 1)]
    got = np.tensordot(B, Sm, axes=axes)
    wide = "complex128" if dtype.startswith("complex") else "float64"
    want = orc.blocked_tensordot(orc.split_blocks(big.astype(wide), (3, 4, 1024, 5)), orc.split_blocks(small.astype(wide), (4, 5, 4)), axes, order="C")
    assert got.grid == want.shape and got.blockshape == want.flat[0].shape
    assert seen["nd"] == 1 and seen["grouped"] == 0 and seen["permute"] == 0, seen
    assert _rel(got, orc.to_numpy(want)) <= TOL[dtype]


def test_maybearray_accumulates_through_the_apply_route(monkeypatch):
    import rosnet_b200 as rn

    rng = np.random.default_rng(12)
    state = _rand(rng, (2,) * 19, "complex128")
    gates = [_rand(rng, (2, 2, 2, 2), "complex128") for _ in range(3)]
    S = rn.array(state, blockshape=state.shape, inner="cuda")
    acc = rn.MaybeArray()
    ref = 0
    seen = _spy(monkeypatch)
    for g in gates:
        rn.commutative(acc, rn.array(g, blockshape=g.shape, inner="cuda"), S, [(2, 3), (5, 11)])     # accumulate=1 after the first
        ref = ref + np.tensordot(g, state, axes=[(2, 3), (5, 11)])
    assert seen == {"nd": 3, "grouped": 0, "permute": 0}, seen
    assert _rel(np.asarray(acc), ref) <= 1e-12


def test_apply_route_can_be_switched_off(monkeypatch):
    """RNB_SIMT=0 keeps everything on the tensor-core route, RNB_APPLY=0 only what the apply kernel would take: same numbers."""
    import rosnet_b200 as rn

    rng = np.random.default_rng(13)
    state, gate = _rand(rng, (2,) * 21, "complex64"), _rand(rng, (2, 2, 2, 2), "complex64")
    S = rn.array(state, blockshape=state.shape, inner="cuda")
    G = rn.array(gate, blockshape=gate.shape, inner="cuda")
    seen0 = _spy(monkeypatch)
    a = np.asarray(np.tensordot(G, S, axes=[(2, 3), (4, 9)]))
    assert seen0 == {"nd": 1, "grouped": 0, "permute": 0}
    monkeypatch.undo()
    monkeypatch.setenv("RNB_SIMT", "0")
    b = np.asarray(np.tensordot(G, S, axes=[(2, 3), (4, 9)]))
    assert _rel(a, b.astype(np.complex128)) <= 2e-6
    monkeypatch.delenv("RNB_SIMT")
    monkeypatch.setenv("RNB_APPLY", "0")          # only the apply kernel off: the contraction goes to the tensor cores as well
    seen = _spy(monkeypatch)
    c = np.asarray(np.tensordot(G, S, axes=[(2, 3), (4, 9)]))
    assert seen["nd"] == 0 and seen["grouped"] == 1
    assert _rel(c, a.astype(np.complex128)) <= 2e-6


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_host_blocks_streamed_through_the_apply_route(dtype, monkeypatch):
    """Host-block operands in GEMM layout with several output blocks take the pipelined host path (engine/pipeline.py: chunks of
    output blocks, copies overlapped with compute); each chunk's launch is still the apply kernel when one operand is tiny."""
    import rosnet_b200 as rn
    from rosnet_b200.engine import pipeline

    rng = np.random.default_rng(21)
    big = _rand(rng, (8, 2 ** 18, 8), dtype)           # blocks (2, 2^18, 8): [rows][k = 8], 4 blocks
    small = _rand(rng, (8, 6), dtype)
    B = rn.array(big, blockshape=(2, 2 ** 18, 8))       # host blocks
    Sm = rn.array(small, blockshape=(8, 6))
    monkeypatch.setattr(pipeline, "MIN_BYTES", 1 << 20)
    seen = _spy(monkeypatch)
    got = np.tensordot(B, Sm, axes=[(2,), (0,)])
    ref = np.tensordot(big, small, axes=[(2,), (0,)])
    assert got.shape == ref.shape and got.grid == (4, 1, 1)
    assert seen["nd"] >= 1 and seen["grouped"] == 0 and seen["permute"] == 0, seen
    assert _rel(got, ref) <= TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("small_first", [True, False])
def test_apply_kernel_writes_nothing_outside_its_result(dtype, small_first):
    """Guard bands (compute-sanitizer is not available on the GPU pool): the result block sits between two canary-filled
    blocks of one allocation; sizes that are no multiple of the 256 (x J) indices a CTA unit covers, so the clamped tail of
    the last unit runs.  The bands must come back untouched, the block itself equal to numpy.tensordot."""
    import torch

    from rosnet_b200.array.device import DeviceStorage
    from rosnet_b200.engine import runtime
    from rosnet_b200.engine.plan import plan_tensordot

    rng = np.random.default_rng(31)
    big_shape = (257, 5, 1031, 3)                 # free extent 257 * 1031 = 264967 (prime factors only), k = 15
    big = _rand(rng, big_shape, dtype)
    small = _rand(rng, (7, 3, 5), dtype)          # 7 rows, contracted (3, 5) listed in the other order
    if small_first:
        shapes, axes, ref = (small, big), [(1, 2), (3, 1)], np.tensordot(small, big, axes=[(1, 2), (3, 1)])
    else:
        shapes, axes, ref = (big, small), [(3, 1), (1, 2)], np.tensordot(big, small, axes=[(3, 1), (1, 2)])
    a, b = shapes
    plan = plan_tensordot((1,) * a.ndim, a.shape, (1,) * b.ndim, b.shape, dtype, axes)
    assert plan.nd_apply and plan.n_out == 1
    tdt = getattr(torch, dtype)
    sa, sb = DeviceStorage(a.nbytes), DeviceStorage(b.nbytes)
    sa.tensor[:a.nbytes].view(tdt).copy_(torch.from_numpy(np.ascontiguousarray(a).reshape(-1)))
    sb.tensor[:b.nbytes].view(tdt).copy_(torch.from_numpy(np.ascontiguousarray(b).reshape(-1)))
    blk_bytes = plan.out_block_elems * np.dtype(dtype).itemsize
    out = DeviceStorage(3 * blk_bytes)
    out.tensor.fill_(0xA5)
    runtime.execute_tensordot(plan, sa, a.shape, sb, b.shape, out_storage=out, out_first_block=1)
    torch.cuda.synchronize()
    raw = out.tensor.cpu().numpy()
    assert (raw[:blk_bytes] == 0xA5).all() and (raw[2 * blk_bytes: 3 * blk_bytes] == 0xA5).all()
    got = raw[blk_bytes: 2 * blk_bytes].view(dtype).reshape(ref.shape)
    assert _rel(got, ref) <= TOL[dtype]
    # accumulate = 1 into the same block: twice the result, bands still untouched
    runtime.execute_tensordot(plan, sa, a.shape, sb, b.shape, out_storage=out, out_first_block=1, accumulate=True)
    torch.cuda.synchronize()
    raw = out.tensor.cpu().numpy()
    assert (raw[:blk_bytes] == 0xA5).all() and (raw[2 * blk_bytes: 3 * blk_bytes] == 0xA5).all()
    assert _rel(raw[blk_bytes: 2 * blk_bytes].view(dtype).reshape(ref.shape), 2 * ref) <= TOL[dtype]
