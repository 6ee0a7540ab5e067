"""``MaybeArray``: an accumulator that is empty until the first ``+=``.

Reference: ``rosnet/array/maybe.py:5-38`` — the target of the ``commutative`` reduction
(``rosnet/array/compss/__init__.py:441-444``): the first in-place ufunc adopts the operand, later
ones apply the ufunc.  In the B200 build the k-block sum is fused inside the grouped GEMM, so
this class is only the API-compatible face for callers that accumulate partial results
themselves; ``isinit`` uses an identity test (the reference's ``!= None`` turns elementwise once an
ndarray is stored).
"""
import numpy as np

from ..core.mixin import ArrayFunctionMixin


class MaybeArray(np.lib.mixins.NDArrayOperatorsMixin, ArrayFunctionMixin):
    def __init__(self, array=None):
        self._array = array

    @property
    def isinit(self) -> bool:
        return self._array is not None

    @property
    def shape(self):
        return self.__array__().shape

    @property
    def dtype(self):
        return self.__array__().dtype

    def __getstate__(self):
        return {"init": self.isinit, "array": self._array}

    def __setstate__(self, d):
        self._array = d["array"]

    def __array__(self, dtype=None, copy=None):
        if not self.isinit:
            raise ValueError("array not initialized")
        out = np.asarray(self._array)
        return out if dtype is None else out.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if ufunc.nin > 2 or method != "__call__":
            return NotImplemented
        out = kwargs.get("out")
        inplace = out is not None and any(o is self for o in (out if isinstance(out, tuple) else (out,)))
        others = [x for x in inputs if x is not self]
        if not self.isinit:
            other = others[0] if others else None
            if inplace:
                self._array = np.array(other, copy=True)
                return self
            return other
        args = [self._array if x is self else x for x in inputs]
        if inplace:
            kwargs = dict(kwargs)
            kwargs["out"] = (self._array,)
            ufunc(*args, **kwargs)
            return self
        return ufunc(*args, **kwargs)
