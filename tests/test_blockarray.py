"""BlockArray layout contract: the reference's own constructor / tensordot-layout tests
(``test/array/test_blockarray.py:9-125``) restated against rosnet_b200, plus the host-side
behaviour around them.  No GPU needed: nothing here computes."""
from math import prod

import numpy as np
import pytest

import rosnet_b200 as rn
from rosnet_b200 import BlockArray
from mock import MockArray


class TestConstructors:
    gridshape = (2, 2, 2)
    blockshape = (4, 8, 1)
    dtype = np.float64

    def _grid_of(self, make):
        data = np.empty(self.gridshape, dtype=object)
        for idx in np.ndindex(*self.gridshape):
            data[idx] = make()
        return data

    def _check(self, arr, grid=None):
        assert arr.grid == (grid or self.gridshape)
        assert arr.blockshape == self.blockshape
        assert arr.dtype == self.dtype

    def test_nested_list_of_ndarray(self):
        self._check(BlockArray(self._grid_of(lambda: np.empty(self.blockshape, dtype=self.dtype)).tolist()))

    def test_nested_list_of_array(self):
        self._check(BlockArray(self._grid_of(lambda: MockArray(self.blockshape, dtype=self.dtype)).tolist()))

    def test_list_of_ndarray(self):
        self._check(BlockArray([np.zeros(self.blockshape, dtype=self.dtype)] * prod(self.gridshape), grid=self.gridshape))

    def test_list_of_array(self):
        self._check(BlockArray([MockArray(self.blockshape, dtype=self.dtype)] * prod(self.gridshape), grid=self.gridshape))

    def test_nested_ndarray(self):
        self._check(BlockArray(self._grid_of(lambda: np.empty(self.blockshape, dtype=self.dtype))))

    def test_nested_array(self):
        self._check(BlockArray(self._grid_of(lambda: MockArray(self.blockshape, dtype=self.dtype))))

    def test_ndarray(self):
        arr = BlockArray(np.zeros(self.blockshape, dtype=self.dtype))
        self._check(arr, grid=(1, 1, 1))
        assert arr.blockshape == arr.shape == self.blockshape

    def test_array(self):
        arr = BlockArray(MockArray(self.blockshape, dtype=self.dtype))
        self._check(arr, grid=(1, 1, 1))
        assert arr.blockshape == arr.shape == self.blockshape

    def test_invalid(self):
        with pytest.raises(ValueError):
            BlockArray(object())
        with pytest.raises(ValueError):
            BlockArray([np.zeros((2, 2))], grid=(1,))  # block ndim != grid ndim


def test_properties():
    a = BlockArray([np.zeros((4, 8, 1), dtype=np.complex64)] * 8, grid=(2, 2, 2))
    assert a.shape == (8, 16, 2) and a.ndim == 3 and a.nblock == 8
    assert a.size == 256 and a.itemsize == 8 and a.nbytes == 2048
    assert a.blocksize == 32 and a.blocknbytes == 256
    assert "BlockArray(shape=(8, 16, 2)" in str(a)
    assert not a.is_device


def test_blocks_are_adopted_by_reference():
    blk = np.zeros((2, 2))
    a = BlockArray([blk], grid=(1, 1))
    assert a.data[0, 0] is blk


def test_tensordot_layout_contract_without_compute():
    """The reference's only tensordot test asserts layout (``test_blockarray.py:115-125``); the
    plan carries exactly that information and needs no device."""
    from rosnet_b200.engine.plan import plan_tensordot

    p = plan_tensordot((2, 2, 2), (4, 1, 1), (2, 2, 2), (4, 1, 1), np.float64, [(0, 1), (0, 2)])
    assert tuple(g * b for g, b in zip(p.out_grid, p.out_blockshape)) == (2, 2)
    assert p.out_blockshape == (1, 1)
    assert p.out_grid == (2, 2)


def test_constructors_zeros_ones_full_rand_array():
    z = rn.zeros((8, 4), dtype=np.float32, blockshape=(4, 2))
    assert z.grid == (2, 2) and z.blockshape == (4, 2) and z.dtype == np.float32
    assert np.array_equal(np.asarray(z), np.zeros((8, 4), np.float32))
    o = rn.ones((4, 4), blockshape=(2, 2))
    assert o.dtype == np.dtype(int) and np.asarray(o).sum() == 16
    f = rn.full((2, 6), 2.5, blockshape=(1, 3))
    assert f.dtype == np.float64 and np.all(np.asarray(f) == 2.5)
    # the reference's rand builds grid (ndim,); the intent is shape // blockshape
    r = rn.rand((8, 8), (4, 4))
    assert r.grid == (2, 2) and r.blockshape == (4, 4) and r.dtype == np.float64
    d = np.asarray(r)
    assert d.shape == (8, 8) and (d >= 0).all() and (d < 1).all()
    x = np.arange(24.0).reshape(4, 6)
    a = rn.array(x, blockshape=(2, 3))
    assert a.grid == (2, 2) and np.array_equal(np.asarray(a), x)
    assert np.array_equal(a.data[1, 0], x[2:4, 0:3])
    with pytest.raises(ValueError):
        rn.array(x, blockshape=(3, 3))


def test_getitem_setitem_deepcopy():
    import copy

    x = np.arange(24.0).reshape(4, 6)
    a = rn.array(x, blockshape=(2, 3))
    assert a[3, 4] == x[3, 4]
    a[3, 4] = -1.0
    assert a[3, 4] == -1.0
    b = copy.deepcopy(a)
    b[0, 0] = 99.0
    assert a[0, 0] == 0.0
    with pytest.raises(IndexError):
        a[1]


def test_to_numpy_host_blocks_and_scalar():
    x = np.arange(16.0).reshape(2, 2, 4)
    a = rn.array(x, blockshape=(1, 2, 2))
    assert np.array_equal(rn.to_numpy(a), x)
    s = BlockArray(np.asarray(3.0))
    assert s.grid == () and np.asarray(s) == 3.0


def test_ufunc_is_todo_like_the_reference():
    a = rn.zeros((2, 2), dtype=np.float64, blockshape=(1, 1))
    with pytest.raises(NotImplementedError):
        a + 1
