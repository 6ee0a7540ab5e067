"""rnb_permute through the C ABI vs numpy (pure data movement: bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(src: np.ndarray, perm, dst_pad=0, src_view=None):
    """Permute `src` (C-contiguous) by `perm` on the device; returns the device result as ndarray.
    dst_pad > 0 pads the destination's last axis (tests arbitrary destination strides)."""
    import torch

    from rosnet_b200.engine import lib

    lib.load()
    es = src.dtype.itemsize
    d_src = torch.from_numpy(src.view(np.uint8).reshape(-1)).cuda()
    out_shape = tuple(src.shape[p] for p in perm)
    padded = out_shape[:-1] + (out_shape[-1] + dst_pad,)
    d_dst = torch.zeros(int(np.prod(padded)) * es, dtype=torch.uint8, device="cuda")
    src_strides = [int(np.prod(src.shape[i + 1:])) for i in range(src.ndim)]
    dst_strides_out = [int(np.prod(padded[i + 1:])) for i in range(len(padded))]
    # descriptor is indexed by SOURCE axis: axis perm[j] of src lands on axis j of dst
    dst_strides = [0] * src.ndim
    for j, p in enumerate(perm):
        dst_strides[p] = dst_strides_out[j]
    lib.permute(d_src.data_ptr(), d_dst.data_ptr(), es, list(src.shape), src_strides, dst_strides, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_dst.cpu().numpy().view(src.dtype).reshape(padded)
    return got[..., : out_shape[-1]]


def _data(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    if np.dtype(dtype).kind == "c":
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dtype).reshape(shape)
    return rng.standard_normal(n).astype(dtype).reshape(shape)


CASES = [
    ((64, 64), (1, 0)),
    ((100, 37), (1, 0)),
    ((3, 1000), (1, 0)),
    ((1000, 3), (1, 0)),
    ((7, 5, 3), (2, 0, 1)),
    ((16, 32, 32, 32), (1, 3, 0, 2)),
    ((4, 32, 32, 32, 32), (0, 2, 4, 1, 3)),        # C1 matricize: block axis leading
    ((2, 2, 16, 16, 8, 8), (0, 2, 4, 1, 3, 5)),    # blockify-like interleave
    ((2,) * 12, (11, 3, 7, 0, 9, 1, 5, 10, 2, 8, 4, 6)),
    ((2,) * 18, tuple(np.random.default_rng(5).permutation(18))),   # merged rank > 16 -> generic kernel
    ((5, 4, 3, 2, 7), (4, 3, 2, 1, 0)),
    ((257, 129), (1, 0)),
    ((6, 1, 50, 1, 9), (4, 1, 0, 3, 2)),
    ((1,), (0,)),
    ((4096,), (0,)),
    ((33, 65, 17), (0, 1, 2)),                     # identity (pure copy)
    ((33, 65, 17), (1, 0, 2)),                     # shared fastest axis
    ((3, 33331, 2), (0, 2, 1)),                    # 2-element source rows: one load unit per row (divisor 1 in the tile index maths)
    ((3, 2, 33331), (0, 2, 1)),                    # ... and one store unit per destination row
    ((2, 33331), (1, 0)),
    ((1021, 6), (1, 0)),
    # shapes the all-TMA transpose kernel takes for 8/16-byte elements (tile-multiple extents)
    ((128, 48), (1, 0)),                           # 128-long destination runs, 3 column tiles
    ((96, 32), (1, 0)),                            # run box 96 (6 run groups)
    ((3, 32, 64), (0, 2, 1)),
    ((2, 4, 32, 32), (0, 3, 1, 2)),                # destination run continues over a second axis (box 4 x 32)
    ((2, 8, 3, 16, 48), (0, 2, 4, 1, 3)),          # C1-like interleave with odd extents around the tile axes
    ((256, 272), (1, 0)),                          # several tiles per CTA, ragged tile count
    ((5, 16, 7, 16), (2, 3, 0, 1)),
    # copy-like moves (shared fastest axis, inner run >= 256 B): TMA load -> TMA store
    ((4, 64, 6, 64), (0, 2, 1, 3)),                # blockify-like, one inner tile
    ((3, 5, 1024), (1, 0, 2)),                     # inner run of several 2 KB tiles
    ((2, 3, 4, 96), (2, 0, 1, 3)),
    ((7, 2, 3, 2, 128), (3, 1, 2, 0, 4)),          # five axes, odd extents
]


@pytest.mark.parametrize("dtype", ["float32", "float64", "complex128"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_permute_bit_exact(case, dtype):
    shape, perm = CASES[case]
    src = _data(shape, dtype, case)
    got = _run(src, perm)
    assert np.array_equal(got, np.transpose(src, perm))


@pytest.mark.parametrize("case", range(len(CASES)))
def test_permute_old_tiled_kernel_still_exact(case, monkeypatch):
    """8-byte elements normally take the micro-transpose kernel; force the LUT-tiled one."""
    monkeypatch.setenv("RNB_PERMUTE_KERNEL", "tiled")
    shape, perm = CASES[case]
    src = _data(shape, "float64", case)
    assert np.array_equal(_run(src, perm), np.transpose(src, perm))


@pytest.mark.parametrize("case", [0, 4, 8, 9, 12])
def test_permute_generic_fallback_kernel(case, monkeypatch):
    monkeypatch.setenv("RNB_PERMUTE_KERNEL", "generic")
    shape, perm = CASES[case]
    src = _data(shape, "complex128", case)
    assert np.array_equal(_run(src, perm), np.transpose(src, perm))


def test_permute_rank_24_to_32_small_extents():
    """tensor-network shaped operands: many extent-2 axes, arbitrary order (merged rank up to 32)"""
    rng = np.random.default_rng(3)
    for rank, dtype in ((24, "complex64"), (22, "float64"), (20, "complex128"), (26, "float32")):
        perm = tuple(int(x) for x in rng.permutation(rank))
        src = _data((2,) * rank, dtype, rank)
        assert np.array_equal(_run(src, perm), np.transpose(src, perm))
    # contraction-like permutation: free axes keep their order, contracted axes move to the end
    rank = 24
    contracted = sorted(int(x) for x in rng.choice(rank, 9, replace=False))
    perm = tuple(i for i in range(rank) if i not in contracted) + tuple(contracted)
    src = _data((2,) * rank, "complex64", 5)
    assert np.array_equal(_run(src, perm), np.transpose(src, perm))


@pytest.mark.parametrize("shape,perm", [
    ((130, 70), (1, 0)), ((66, 34, 6), (2, 1, 0)), ((2, 50, 2, 30), (3, 2, 1, 0)), ((4, 6, 8, 10, 12), (4, 2, 0, 3, 1)),
    ((2,) * 14, tuple(reversed(range(14)))), ((3, 4, 258, 62), (0, 3, 1, 2)), ((1026, 514), (1, 0)),
])
def test_permute_micro_kernel_edges(shape, perm):
    """even extents, ragged tiles: exercises partial chunks / skipped chunks of the micro kernel"""
    for dtype in ("float64", "complex64"):
        src = _data(shape, dtype, 21)
        assert np.array_equal(_run(src, perm), np.transpose(src, perm))


@pytest.mark.parametrize("dtype", ["float64", "complex64"])
def test_permute_padded_destination(dtype):
    src = _data((48, 10, 6), dtype, 3)
    got = _run(src, (2, 0, 1), dst_pad=3)
    assert np.array_equal(got, np.transpose(src, (2, 0, 1)))


@pytest.mark.parametrize("mode", ["bulk", "cpasync"])
def test_permute_load_paths_agree(mode, monkeypatch):
    if mode == "cpasync":
        monkeypatch.setenv("RNB_PERMUTE_LOAD", "cpasync")
    src = _data((8, 32, 32, 32, 32), "float64", 9)
    got = _run(src, (0, 2, 4, 1, 3))
    assert np.array_equal(got, np.transpose(src, (0, 2, 4, 1, 3)))
    src = _data((130, 70), "float64", 10)   # edge tiles on both axes
    assert np.array_equal(_run(src, (1, 0)), src.T)


def test_permute_large_is_exact_via_checksum():
    """128 MiB tensor (BASELINE full size of one operand): checksum of checksums + spot rows."""
    src = _data((64, 64, 64, 64), "float64", 11)
    got = _run(src, (1, 3, 0, 2))
    want = np.transpose(src, (1, 3, 0, 2))
    assert got.view(np.uint64).sum() == want.view(np.uint64).sum()
    assert np.array_equal(got[5, 7], want[5, 7]) and np.array_equal(got[63, 63], want[63, 63])
    assert np.array_equal(got, want)
