"""Fused GEMM + exchange epilogue (rnb_contract_desc.c_peers) on ONE GPU: the "owners" are two
buffers on the same device, so the arithmetic of the peer path — bulk reductions
(cp.reduce.async.bulk add.f64) for interior warp tiles, system-scope scalar reductions for edge
tiles — is checked without a second GPU.  The multi-process form runs in bench.py
(--workload bench_schmidt_matrix --exchange fused), which checks parity itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(x, ref):
    return float(np.linalg.norm((np.asarray(x) - ref).ravel()) / (np.linalg.norm(ref.ravel()) or 1.0))


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("mode", ["4m", "3m"])
@pytest.mark.parametrize("bulk", ["1", "0"])
@pytest.mark.parametrize("n_blk", [192, 200])    # 192: every warp tile interior or empty; 200: a partly filled warp tile (scalar path)
def test_peer_reduction_matches_plain_result(dtype, mode, bulk, n_blk, monkeypatch):
    import rosnet_b200 as rn
    from rosnet_b200.array.device import DeviceStorage
    from rosnet_b200.engine import plan as _plan, runtime

    if dtype == "float64" and mode == "3m":
        pytest.skip("3M is a complex mode")
    monkeypatch.setenv("RNB_PEER_BULK", bulk)
    monkeypatch.setenv("RNB_SIMT", "0")
    rng = np.random.default_rng(31)
    a = rng.standard_normal((3 * 256, 2 * 128))
    b = rng.standard_normal((2 * 128, 2 * n_blk))
    if dtype == "complex128":
        a = a + 1j * rng.standard_normal(a.shape)
        b = b + 1j * rng.standard_normal(b.shape)
    A = rn.array(a, blockshape=(256, 128), inner="cuda")
    B = rn.array(b, blockshape=(128, n_blk), inner="cuda")
    axes = [(1,), (0,)]
    p = _plan.plan_tensordot(A.grid, A.blockshape, B.grid, B.blockshape, np.dtype(dtype), axes)
    assert p.n_out == 6
    chunk = 3                                     # blocks per owner
    nb = p.out_block_elems * np.dtype(dtype).itemsize
    owners = [DeviceStorage(chunk * nb, zero=True) for _ in range(2)]
    dummy = DeviceStorage(nb)
    with rn.tuning.configure(complex_mode=mode):
        for _ in range(2):                        # two "ranks" adding their partial: the result is twice the contraction
            runtime.execute_tensordot(p, A._storage, A.blockshape, B._storage, B.blockshape, out_storage=dummy,
                                      peers=[o.ptr for o in owners], blocks_per_peer=chunk)
    want = 2 * (a @ b)
    for owner in range(2):
        got = rn.BlockArray._from_storage(owners[owner], (chunk,) + (1,) * 1, (256, n_blk), np.dtype(dtype))
        for local in range(chunk):
            g = owner * chunk + local
            gi, gj = divmod(g, 2)
            blk = np.asarray(got.data.flat[local])
            ref = want[gi * 256:(gi + 1) * 256, gj * n_blk:(gj + 1) * n_blk]
            assert _rel(blk, ref) <= 1e-12, (owner, local)
