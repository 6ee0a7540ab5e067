"""Device-backed MaybeArray (``res += tensordot`` reaching the kernel's accumulate=1), the sequence overload on device
blocks without a host round trip, and mixed BlockArray / ndarray operands
(reference: rosnet/array/maybe.py:5-38, rosnet/array/compss/task/tensordot.py:45-49, rosnet/array/compss/__init__.py:412-451)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(x, ref):
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(np.asarray(ref).ravel()) or 1.0))


def _rand(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("complex128", 1e-12), ("complex64", 1e-5), ("float32", 1e-5)])
def test_maybearray_accumulates_on_the_device(dtype, tol, contract_route):
    import rosnet_b200 as rn
    from rosnet_b200.engine import lib

    rng = np.random.default_rng(3)
    pairs = [(_rand(rng, (96, 64), dtype), _rand(rng, (64, 80), dtype)) for _ in range(3)]
    res = rn.MaybeArray()
    assert not res.isinit
    before = dict(lib.launch_counter)
    for a, b in pairs:
        A = rn.array(a, blockshape=(48, 32), inner="cuda")
        B = rn.array(b, blockshape=(32, 40), inner="cuda")
        rn.commutative(res, A, B, [(1,), (0,)])          # res += tensordot(a, b): one grouped launch each, accumulate=1 after the first
    assert res.isinit and isinstance(res.value, rn.BlockArray) and res.value.is_device
    assert lib.launch_counter["block_sum"] == before["block_sum"]      # no separate add pass
    wide = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
    want = sum(a.astype(wide) @ b.astype(wide) for a, b in pairs)
    assert res.shape == (96, 80) and res.dtype == np.dtype(dtype)
    assert _rel(np.asarray(res), want) <= tol
    # res += BlockArray: device-side add
    extra = _rand(rng, (96, 80), dtype)
    res += rn.array(extra, blockshape=(48, 40), inner="cuda")
    assert lib.launch_counter["block_sum"] == before["block_sum"] + 1
    assert _rel(np.asarray(res), want + extra.astype(wide)) <= tol
    # an empty accumulator adopts a BlockArray handed to it
    m = rn.MaybeArray()
    m += rn.array(extra, blockshape=(48, 40), inner="cuda")
    assert isinstance(m.value, rn.BlockArray) and np.array_equal(np.asarray(m), extra)
    with pytest.raises(ValueError):
        rn.commutative(res, rn.array(pairs[0][0], blockshape=(96, 64), inner="cuda"), rn.array(pairs[0][1], blockshape=(64, 80), inner="cuda"), [(1,), (0,)])


def test_sequence_tensordot_on_device_blocks_stays_on_the_device(monkeypatch):
    import rosnet_b200 as rn
    from rosnet_b200.array import device as dev

    rng = np.random.default_rng(4)
    a, b = rng.standard_normal((4 * 32, 48)), rng.standard_normal((4 * 32, 56))
    A = rn.array(a, blockshape=(32, 48), inner="cuda")
    B = rn.array(b, blockshape=(32, 56), inner="cuda")
    la, lb = list(A.data.flat), [B.data.flat[i] for i in (2, 0, 3, 1)]     # consecutive run / scattered blocks
    reads = []
    orig = dev.DeviceBlock.__array__
    monkeypatch.setattr(dev.DeviceBlock, "__array__", lambda self, *a_, **k: (reads.append(1), orig(self, *a_, **k))[1])
    got = rn.tensordot(la, lb, [(0,), (0,)], method="commutative")
    assert not reads, "device blocks were read back to the host"
    monkeypatch.undo()
    want = sum(a[i * 32:(i + 1) * 32].T @ b[j * 32:(j + 1) * 32] for i, j in zip(range(4), (2, 0, 3, 1)))
    assert _rel(np.asarray(got), want) <= 1e-12


def test_mixed_blockarray_ndarray_operands():
    import rosnet_b200 as rn

    rng = np.random.default_rng(5)
    a, b = rng.standard_normal((64, 96, 8)), rng.standard_normal((8, 96, 40))
    A = rn.array(a, blockshape=(32, 48, 4), inner="cuda")
    C = np.tensordot(A, b, [(1, 2), (1, 0)])              # the ndarray is cut to match a's blocks along the contracted axes
    assert isinstance(C, rn.BlockArray) and C.grid == (2, 1) and C.blockshape == (32, 40)
    assert _rel(np.asarray(C), np.tensordot(a, b, [(1, 2), (1, 0)])) <= 1e-12
    C = np.tensordot(b, A, [(1, 0), (1, 2)])
    assert C.grid == (1, 2) and _rel(np.asarray(C), np.tensordot(b, a, [(1, 0), (1, 2)])) <= 1e-12
    with pytest.raises(ValueError):
        np.tensordot(A, b, [(0,), (0,)])
