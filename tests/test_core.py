"""core helpers and protocols (reference ``test/core/test_math.py:37-110``,
``test/core/test_interface.py:8-51``)."""
import numpy as np
import pytest

from rosnet_b200 import BlockArray, MaybeArray
from rosnet_b200.array.device import DeviceBlock
from rosnet_b200.core.interface import Array, ArrayConvertable
from rosnet_b200.core.util import isunique, join_idx, measure_shape, nest_level, normalize_axes, result_shape, space
from mock import MockArray


class TestNestLevel:
    def test_ndarray(self):
        assert nest_level(np.empty((1,))) == 0
        assert measure_shape(np.empty((1,))) == ()

    def test_array(self):
        assert nest_level(MockArray((1,), dtype=np.generic)) == 0
        assert measure_shape(MockArray((1,), dtype=np.generic)) == ()

    def test_ndarray_of_ndarray(self):
        arr = np.empty((1,), dtype=object)
        arr.flat[0] = np.empty((1,))
        assert nest_level(arr) == 1 and measure_shape(arr) == (1,)

    def test_ndarray_of_array(self):
        arr = np.empty((1,), dtype=object)
        arr.flat[0] = MockArray((1,), dtype=np.generic)
        assert nest_level(arr) == 1 and measure_shape(arr) == (1,)

    @pytest.mark.parametrize("level", [1, 2, 3, 4, 5])
    @pytest.mark.parametrize("leaf", [lambda: np.zeros((1,)), lambda: MockArray((1,))])
    def test_nested_lists(self, level, leaf):
        arr = np.empty(tuple([1] * level), dtype=object)
        arr.flat[0] = leaf()
        assert nest_level(arr.tolist()) == level
        assert measure_shape(arr.tolist()) == tuple([1] * level)


@pytest.mark.parametrize("cls", [MockArray, np.ndarray, BlockArray, DeviceBlock, MaybeArray])
def test_array_convertable(cls):
    assert issubclass(cls, ArrayConvertable)


@pytest.mark.parametrize("cls", [MockArray, np.ndarray, BlockArray, DeviceBlock])
def test_array_protocol(cls):
    assert issubclass(cls, Array)


def test_index_helpers():
    assert isunique([0, 2, 1]) and not isunique([0, 0])
    assert list(space([2, 2])) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert result_shape((2, 3, 4), (4, 5, 3), [(1, 2), (2, 0)]) == (2, 5)
    assert join_idx((7, 9), (1, 2), (2, 0)) == (2, 7, 1, 9)
    assert normalize_axes(2, 4, 4) == ([2, 3], [0, 1])
    assert normalize_axes([(-1,), (0,)], 3, 2) == ([2], [0])
    assert normalize_axes((1, 0), 3, 2) == ([1], [0])
    with pytest.raises(ValueError):
        normalize_axes([(0, 0), (0, 1)], 2, 2)
    with pytest.raises(ValueError):
        normalize_axes([(0, 1), (0,)], 2, 2)


def test_maybe_array_accumulator():
    m = MaybeArray()
    assert not m.isinit
    with pytest.raises(ValueError):
        np.asarray(m)
    x = np.arange(4.0)
    m += x            # first += adopts the operand (rosnet/array/maybe.py:27-34)
    assert m.isinit and np.array_equal(np.asarray(m), x)
    m += x
    assert np.array_equal(np.asarray(m), 2 * x)
    assert np.array_equal(x, np.arange(4.0))  # adopted by value here: operand not aliased
    import pickle

    m2 = pickle.loads(pickle.dumps(m))
    assert np.array_equal(np.asarray(m2), 2 * x)
