"""Seeded blocked-contraction cases shared by the golden generator and the parity tests.

Each case: operand shapes, blockshapes, dtypes, seeds and the explicit ``axes`` pair.  Inputs are
``default_rng(seed).standard_normal`` (plus ``1j *`` a second draw for complex dtypes), cast to
dtype; case ``ones_ref_test`` mirrors the reference's only tensordot test
(``test/array/test_blockarray.py:115-125``) and ``uniform_f64`` mirrors ``rand``'s U[0,1)
(``rosnet/array/block/__init__.py:377``).
"""
import numpy as np

CASES = [
    # name, a_shape, a_block, b_shape, b_block, dtype_a, dtype_b, axes, kind
    dict(name="ones_ref_test", a=(8, 2, 2), ab=(4, 1, 1), b=(8, 2, 2), bb=(4, 1, 1), da="float64", db="float64", axes=[(0, 1), (0, 2)], fill="ones"),
    dict(name="rank3_f64", a=(8, 6, 4), ab=(4, 3, 2), b=(8, 6, 10), bb=(4, 3, 5), da="float64", db="float64", axes=[(0, 1), (0, 1)]),
    dict(name="c1_mini_nopermute", a=(8, 8, 8, 8), ab=(4, 4, 4, 4), b=(8, 8, 8, 8), bb=(4, 4, 4, 4), da="float64", db="float64", axes=[(2, 3), (0, 1)]),
    dict(name="c1_mini_permute", a=(8, 8, 8, 8), ab=(4, 4, 4, 4), b=(8, 8, 8, 8), bb=(4, 4, 4, 4), da="float64", db="float64", axes=[(0, 2), (1, 3)]),
    dict(name="schmidt_mini_c128", a=(12, 8), ab=(3, 4), b=(10, 8), bb=(5, 4), da="complex128", db="complex128", axes=[(1,), (1,)]),
    dict(name="rank3_c64", a=(6, 8, 4), ab=(3, 4, 2), b=(4, 8, 6), bb=(2, 4, 3), da="complex64", db="complex64", axes=[(1, 2), (1, 0)]),
    dict(name="rank3_f32", a=(6, 8, 4), ab=(3, 4, 2), b=(4, 8, 6), bb=(2, 4, 3), da="float32", db="float32", axes=[(1, 2), (1, 0)]),
    dict(name="full_contraction_scalar", a=(4, 6), ab=(2, 3), b=(4, 6), bb=(2, 3), da="float64", db="float64", axes=[(0, 1), (0, 1)]),
    dict(name="reversed_axes", a=(8, 6, 4, 10), ab=(4, 3, 2, 5), b=(4, 6, 12, 8), bb=(2, 3, 6, 4), da="float64", db="float64", axes=[(1, 0), (1, 3)]),
    dict(name="slices_as_blocks_c64", a=(2, 2, 8, 4), ab=(1, 1, 8, 4), b=(2, 4, 2, 6), bb=(1, 4, 1, 6), da="complex64", db="complex64", axes=[(0, 3), (0, 1)]),
    dict(name="mixed_f64_c128", a=(6, 4), ab=(3, 2), b=(4, 6), bb=(2, 3), da="float64", db="complex128", axes=[(1,), (0,)]),
    dict(name="uniform_f64", a=(8, 8), ab=(4, 4), b=(8, 8), bb=(4, 4), da="float64", db="float64", axes=[(1,), (0,)], fill="uniform"),
    dict(name="single_block", a=(4, 5, 6), ab=(4, 5, 6), b=(6, 5, 3), bb=(6, 5, 3), da="complex128", db="complex128", axes=[(2, 1), (0, 1)]),
    # the reference pairs the wrong blocks here (order='K' inner iteration)
    dict(name="a1_mispaired", a=(8, 6, 4, 10), ab=(4, 3, 2, 5), b=(4, 8, 12, 6), bb=(2, 4, 6, 3), da="float64", db="float64", axes=[(0, 1, 2), (1, 3, 0)], defect="A1"),
]


def make_operand(shape, dtype, seed, fill=None):
    dtype = np.dtype(dtype)
    if fill == "ones":
        return np.ones(shape, dtype=dtype)
    rng = np.random.default_rng(seed)
    if fill == "uniform":
        return rng.random(shape).astype(dtype)
    x = rng.standard_normal(shape)
    if dtype.kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def case_inputs(case, index):
    a = make_operand(case["a"], case["da"], 2 * index, case.get("fill"))
    b = make_operand(case["b"], case["db"], 2 * index + 1, case.get("fill"))
    return a, b
