"""Seeded random parity cases for the whole public route (np.tensordot / np.transpose / np.einsum on BlockArrays ->
__array_function__ -> plan -> C ABI) against the CPU oracle and NumPy on the same inputs.

The golden and hand-written cases pin the reference's behaviour; these cases walk the space between them: ranks 1-5,
0-3 contracted axes in any pairing, ragged block extents (odd, prime, 1), grids of 1-3 blocks per axis, host and device
blocks, every dtype, as routed and with the tensor-core kernels forced.  A second family scales one free and one
contracted extent up so that the tcgen05 / DMMA kernels, split-K and the 3M arrangement see irregular shapes too.
Tolerances as everywhere (BASELINE.json north_star): <= 1e-12 float64 / complex128, <= 1e-5 float32 / complex64;
grid, blockshape and index order exactly equal to the oracle's (rosnet/array/block/__init__.py:286-336)."""
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


def _random_case(rng, big=False):
    """(a_shape, a_block, b_shape, b_block, axes) with matching contracted extents and blockings."""
    if big:   # at least one free axis on each side and one contracted axis
        ra, rb = int(rng.integers(2, 5)), int(rng.integers(2, 5))
        nc = int(rng.integers(1, min(ra, rb, 3)))
    else:
        ra, rb = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        nc = int(rng.integers(0, min(ra, rb, 3) + 1))
    ax_a = [int(x) for x in rng.permutation(ra)[:nc]]
    ax_b = [int(x) for x in rng.permutation(rb)[:nc]]
    blocks = [1, 2, 3, 4, 5, 7, 8, 9]

    def dims(r):
        return [int(rng.choice(blocks)) for _ in range(r)], [int(rng.integers(1, 4)) for _ in range(r)]

    ab, ag = dims(ra)
    bb, bg = dims(rb)
    for i, j in zip(ax_a, ax_b):
        bb[j], bg[j] = ab[i], ag[i]
    if big:
        # one free axis of each operand and one contracted axis grow: m, n in the hundreds, k up to a few thousand
        free_a = [i for i in range(ra) if i not in ax_a]
        free_b = [j for j in range(rb) if j not in ax_b]
        if free_a:
            ab[free_a[0]] = int(rng.choice([96, 130, 257]))
        if free_b:
            bb[free_b[0]] = int(rng.choice([72, 136, 300]))
        if nc:
            ab[ax_a[0]] = bb[ax_b[0]] = int(rng.choice([33, 260, 1100]))
            ag[ax_a[0]] = bg[ax_b[0]] = int(rng.integers(1, 3))
        # keep the whole case small: cap the remaining extents
        for i in range(ra):
            if i not in ax_a[:1] and i not in (free_a[:1] if free_a else []):
                ab[i], ag[i] = min(ab[i], 3), min(ag[i], 2)
        for j in range(rb):
            if j not in ax_b[:1] and j not in (free_b[:1] if free_b else []):
                bb[j], bg[j] = min(bb[j], 3), min(bg[j], 2)
        for i, j in zip(ax_a, ax_b):
            bb[j], bg[j] = ab[i], ag[i]
    sa = tuple(x * g for x, g in zip(ab, ag))
    sb = tuple(x * g for x, g in zip(bb, bg))
    return sa, tuple(ab), sb, tuple(bb), [tuple(ax_a), tuple(ax_b)]


def _check_tensordot(seed, dtype, inner, big):
    import rosnet_b200 as rn
    from oracle import blocked_tensordot as orc

    rng = np.random.default_rng(seed)
    sa, ab, sb, bb, axes = _random_case(rng, big=big)
    k = int(np.prod([sa[i] for i in axes[0]], dtype=np.int64))
    out_elems = int(np.prod(sa, dtype=np.int64) // k) * int(np.prod(sb, dtype=np.int64) // k)
    if out_elems > 2**24 or out_elems * k > 2**33:
        pytest.skip("case too large for a seconds-long oracle run")
    a, b = _rand(rng, sa, dtype), _rand(rng, sb, dtype)
    C = np.tensordot(rn.array(a, blockshape=ab, inner=inner), rn.array(b, blockshape=bb, inner=inner), axes)
    wide = "complex128" if dtype.startswith("complex") else "float64"
    want_grid = orc.blocked_tensordot(orc.split_blocks(a.astype(wide), ab), orc.split_blocks(b.astype(wide), bb), axes, order="C")
    want = orc.to_numpy(want_grid)
    got = np.asarray(C)
    label = f"seed={seed} a={sa}/{ab} b={sb}/{bb} axes={axes}"
    assert C.dtype == np.dtype(dtype), label
    assert C.grid == want_grid.shape and C.blockshape == np.asarray(want_grid.flat[0]).shape, label
    assert got.shape == want.shape, label
    assert _rel(got.astype(wide), want) <= TOL[dtype], label
    # dense truth as well (the oracle's C order equals numpy.tensordot on the assembled tensors)
    assert _rel(got.astype(wide), np.tensordot(a.astype(wide), b.astype(wide), axes)) <= TOL[dtype], label


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("seed", range(24))
def test_random_small_cases(seed, dtype, contract_route):
    _check_tensordot(1000 + seed, dtype, "cuda" if seed % 2 else "numpy", big=False)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("seed", range(12))
def test_random_tensor_core_cases(seed, dtype):
    _check_tensordot(5000 + seed, dtype, "cuda", big=True)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("seed", range(10))
def test_random_transpose_then_tensordot(seed, dtype):
    """transpose (blocks and grid permuted on the device) feeding a contraction: the layout bookkeeping of two ops in a row"""
    import rosnet_b200 as rn

    rng = np.random.default_rng(9000 + seed)
    r = int(rng.integers(2, 6))
    bs = tuple(int(rng.choice([1, 2, 3, 4, 6])) for _ in range(r))
    grid = tuple(int(rng.integers(1, 4)) for _ in range(r))
    shape = tuple(x * g for x, g in zip(bs, grid))
    perm = tuple(int(x) for x in rng.permutation(r))
    a = _rand(rng, shape, dtype)
    A = rn.array(a, blockshape=bs, inner="cuda")
    At = np.transpose(A, perm)
    assert At.shape == tuple(shape[p] for p in perm) and At.blockshape == tuple(bs[p] for p in perm)
    assert np.array_equal(np.asarray(At), np.transpose(a, perm))           # pure data movement: bit-exact
    # contract the transposed tensor with a fresh one over its two leading axes (or one, for rank 2)
    nc = 2 if r > 2 else 1
    bshape = tuple(At.shape[:nc]) + (10,)
    bblock = tuple(At.blockshape[:nc]) + (5,)
    b = _rand(rng, bshape, dtype)
    C = np.tensordot(At, rn.array(b, blockshape=bblock, inner="cuda"), [tuple(range(nc)), tuple(range(nc))])
    want = np.tensordot(np.transpose(a, perm).astype("complex128"), b.astype("complex128"), [tuple(range(nc)), tuple(range(nc))])
    assert _rel(np.asarray(C).astype("complex128"), want) <= TOL[dtype]


@pytest.mark.parametrize("dtype", ["complex64", "float64"])
@pytest.mark.parametrize("seed", range(10))
def test_random_einsum_patterns(seed, dtype):
    """two-operand einsum over BlockArrays with random index sets: contracted, batch (kept on both and in the output),
    summed-out and free indices in a random output order, against numpy.einsum"""
    import rosnet_b200 as rn

    rng = np.random.default_rng(7000 + seed)
    letters = list("abcdefgh")
    n_idx = int(rng.integers(3, 7))
    idx = letters[:n_idx]
    size = {c: int(rng.choice([2, 3, 4, 6])) for c in idx}
    ia = [c for c in idx if rng.random() < 0.65]
    ib = [c for c in idx if rng.random() < 0.65]
    if not ia:
        ia = idx[:1]
    if not ib:
        ib = idx[-1:]
    ia = [ia[i] for i in rng.permutation(len(ia))]
    ib = [ib[i] for i in rng.permutation(len(ib))]
    used = [c for c in idx if c in ia or c in ib]
    out = [c for c in used if rng.random() < 0.6]
    out = [out[i] for i in rng.permutation(len(out))]
    pattern = f"{''.join(ia)},{''.join(ib)}->{''.join(out)}"
    a = _rand(rng, tuple(size[c] for c in ia), dtype)
    b = _rand(rng, tuple(size[c] for c in ib), dtype)

    def blk(c):
        return 2 if size[c] % 2 == 0 and size[c] > 2 else size[c]

    A = rn.array(a, blockshape=tuple(blk(c) for c in ia), inner="cuda")
    B = rn.array(b, blockshape=tuple(blk(c) for c in ib), inner="cuda")
    got = np.asarray(np.einsum(pattern, A, B))
    wide = "complex128" if dtype.startswith("complex") else "float64"
    want = np.einsum(pattern, a.astype(wide), b.astype(wide))
    assert got.shape == want.shape, pattern
    assert _rel(got.astype(wide), want) <= TOL[dtype], pattern


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("seed", range(10))
def test_random_high_rank_extent2_pairs(seed, dtype):
    """tensor-network shaped operands: rank 6-20, every extent 2, a random subset of axes contracted in a random pairing,
    up to two axes per operand cut into blocks of extent 1 (slicing-as-blocks); these go through the LUT-tiled permute
    kernel (with the fused TF32 split for single precision), the n-d strided small kernel and the GEMMs"""
    import rosnet_b200 as rn

    rng = np.random.default_rng(3000 + seed)
    ra, rb = int(rng.integers(6, 21)), int(rng.integers(6, 21))
    nc = int(rng.integers(1, min(ra, rb) + 1))
    while (ra - nc) + (rb - nc) > 22:       # keep the result at or below 2^22 elements
        nc += 1
        if nc > min(ra, rb):
            ra, rb, nc = 16, 16, 8
    ax_a = tuple(int(x) for x in rng.permutation(ra)[:nc])
    ax_b = tuple(int(x) for x in rng.permutation(rb)[:nc])
    ab, bb = [2] * ra, [2] * rb
    for i in rng.permutation(ra)[:int(rng.integers(0, 3))]:
        ab[int(i)] = 1
    for j in rng.permutation(rb)[:int(rng.integers(0, 3))]:
        bb[int(j)] = 1
    for i, j in zip(ax_a, ax_b):
        bb[j] = ab[i]
    a, b = _rand(rng, (2,) * ra, dtype), _rand(rng, (2,) * rb, dtype)
    C = np.tensordot(rn.array(a, blockshape=tuple(ab), inner="cuda"), rn.array(b, blockshape=tuple(bb), inner="cuda"), [ax_a, ax_b])
    wide = "complex128" if dtype.startswith("complex") else "float64"
    want = np.tensordot(a.astype(wide), b.astype(wide), [ax_a, ax_b])
    label = f"seed={seed} ra={ra} rb={rb} axes={[ax_a, ax_b]} ab={ab} bb={bb}"
    assert np.asarray(C).shape == want.shape, label
    assert _rel(np.asarray(C).astype(wide), want) <= TOL[dtype], label


@pytest.mark.parametrize("seed", range(8))
def test_random_reshape_and_sequence(seed):
    """per-block reshape in both orders against the oracle (rosnet/array/block/__init__.py:251-261) and the Sequence
    overload (`:281-283`) on random block lists, host and device blocks"""
    import rosnet_b200 as rn
    from oracle import blocked_tensordot as orc

    rng = np.random.default_rng(11000 + seed)
    r = int(rng.integers(1, 5))
    bs = tuple(int(rng.choice([2, 3, 4, 6])) for _ in range(r))
    grid = tuple(int(rng.integers(1, 4)) for _ in range(r))
    x = _rand(rng, tuple(b_ * g for b_, g in zip(bs, grid)), "complex128")
    vol = int(np.prod(bs))
    # a random factorisation of the block volume as the new block shape
    new, rest = [], vol
    for p in (2, 3, 2, 2, 3):
        if rest % p == 0 and len(new) < r - 1:
            new.append(p)
            rest //= p
    new.append(rest)
    new = tuple(new + [1] * (r - len(new)))[:r] if len(new) <= r else (vol,) + (1,) * (r - 1)
    if int(np.prod(new)) != vol:
        new = (vol,) + (1,) * (r - 1)
    for order in ("F", "C"):
        for inner in ("numpy", "cuda"):
            R = rn.reshape(rn.array(x, blockshape=bs, inner=inner), new, order=order, inplace=False).to_host()
            want = orc.blocked_reshape(orc.split_blocks(x, bs), new, order=order)
            assert R.grid == want.shape and R.blockshape == new
            for idx in np.ndindex(*want.shape):
                assert np.array_equal(R.data[idx], want[idx]), (seed, order, inner, idx)
    n = int(rng.integers(1, 6))
    As = [_rand(rng, (6, 5, 4), "float64") for _ in range(n)]
    Bs = [_rand(rng, (4, 5, 7), "float64") for _ in range(n)]
    want = sum(np.tensordot(p, q, [(1, 2), (1, 0)]) for p, q in zip(As, Bs))
    for method in ("sequential", "commutative", "commutative-but-first"):
        assert _rel(rn.tensordot(As, Bs, [(1, 2), (1, 0)], method=method), want) <= 1e-12
