"""Maximum-size cases: tensors whose byte offsets and element counts pass 2^32, checked through size-independent
properties (SURVEY 8c / task rule: "at BASELINE's full sizes — through properties the domain offers").

The operands are generated on the device and wrapped as BlockArrays (`BlockArray._from_storage`), nothing this large
crosses PCIe; the checks read back a few thousand sampled entries or compare on the device with torch (a checker outside
the product path)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_blockarray(dense, blockshape, dtype):
    """torch dense tensor -> device BlockArray (one permute launch into block-major storage)"""
    import torch

    import rosnet_b200 as rn
    from rosnet_b200.array.device import DeviceStorage
    from rosnet_b200.engine import runtime

    itemsize = np.dtype(dtype).itemsize
    grid = tuple(s // b for s, b in zip(dense.shape, blockshape))
    st = DeviceStorage(dense.numel() * itemsize)
    runtime.dense_to_blocks(dense.contiguous().data_ptr(), st, grid, blockshape, itemsize)
    torch.cuda.synchronize()
    return rn.BlockArray._from_storage(st, grid, blockshape, dtype)


def test_transpose_round_trip_of_an_8_gib_tensor_is_the_identity():
    """complex128, 2^29 elements = 8 GiB: transpose by a permutation and by its inverse gives the original bits back, and the
    first transpose equals torch's permute block by block (byte offsets up to 2^33 in the permute kernels)."""
    import torch

    import rosnet_b200 as rn

    shape, bs = (64, 128, 256, 256), (32, 64, 128, 256)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.randn(shape, dtype=torch.complex128, device="cuda", generator=g)
    A = _device_blockarray(x, bs, "complex128")
    perm = (2, 0, 3, 1)
    inv = tuple(int(i) for i in np.argsort(perm))
    T = rn.transpose(A, perm, inplace=False)
    assert T.shape == tuple(shape[p] for p in perm) and T.blockshape == tuple(bs[p] for p in perm)
    # one block of the transposed array against torch (grid cell (1, 1, 0, 1))
    nb = T.blocksize * 16
    idx = (1, 1, 0, 1)
    flat = int(np.ravel_multi_index(idx, T.grid))
    got = T._storage.tensor[flat * nb:(flat + 1) * nb].view(torch.complex128).reshape(T.blockshape)
    sl = tuple(slice(i * b, (i + 1) * b) for i, b in zip(idx, T.blockshape))
    assert torch.equal(got, x.permute(perm)[sl])
    B = rn.transpose(T, inv, inplace=False)
    assert B.grid == A.grid and B.blockshape == A.blockshape
    assert torch.equal(B._storage.tensor[:A._storage.nbytes], A._storage.tensor[:A._storage.nbytes])


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-5), ("complex64", 1e-5)])
def test_contraction_with_2_pow_32_output_elements(dtype, tol):
    """m = n = 65536, k = 64: 2^32 output elements (16 / 32 GiB), 16 output blocks of 2^28 elements: the epilogue's
    addressing past 2^32 elements; 4096 sampled entries against float64 dot products of the operand rows."""
    import torch

    import rosnet_b200 as rn

    m = n = 65536
    k = 64
    tdt = torch.float32 if dtype == "float32" else torch.complex64
    wide = torch.float64 if dtype == "float32" else torch.complex128
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    a = torch.randn((m, k), dtype=tdt, device="cuda", generator=g)
    b = torch.randn((n, k), dtype=tdt, device="cuda", generator=g)
    A = _device_blockarray(a, (m // 4, k), dtype)
    B = _device_blockarray(b, (n // 4, k), dtype)
    C = np.tensordot(A, B, [(1,), (1,)])
    assert C.shape == (m, n) and C.grid == (4, 4) and C.blockshape == (m // 4, n // 4)
    itemsize = np.dtype(dtype).itemsize
    flat = C._storage.tensor[:C.nblock * C.blocksize * itemsize].view(tdt)
    rng = np.random.default_rng(0)
    rows = torch.from_numpy(rng.integers(0, m, 4096)).cuda()
    cols = torch.from_numpy(rng.integers(0, n, 4096)).cuda()
    # the largest indices explicitly (last element of the last block)
    rows[0], cols[0] = m - 1, n - 1
    bm, bn = m // 4, n // 4
    off = ((rows // bm) * 4 + (cols // bn)) * (bm * bn) + (rows % bm) * bn + (cols % bn)     # block-major element offset
    got = flat[off].to(wide)
    want = (a[rows].to(wide) * b[cols].to(wide)).sum(dim=1)
    err = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
    assert err <= tol, err
    assert int(off.max()) >= 2**32 - 2**20
