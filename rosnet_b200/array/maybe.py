"""``MaybeArray``: an accumulator that is empty until the first ``+=``.

Reference: ``rosnet/array/maybe.py:5-38`` — the target of the ``commutative`` reduction
(``rosnet/array/compss/__init__.py:441-444``, task body ``res += np.tensordot(a, b, axes)`` at
``rosnet/array/compss/task/tensordot.py:45-49``): the first in-place ufunc adopts the operand, later ones apply the
ufunc.  ``isinit`` uses an identity test (the reference's ``!= None`` turns elementwise once an ndarray is stored,
SURVEY section 2.1 row 9).

Device-backed form.  When the accumulated value is a ``BlockArray`` the sum stays in HBM:

* ``res.iadd_tensordot(a, b, axes)`` / ``commutative(res, a, b, axes)`` is the whole task body in ONE grouped launch —
  the first call contracts into a fresh result, later calls run the same launch with ``accumulate=1`` so the kernel
  epilogue adds into the existing blocks (no temporary, no second pass);
* ``res += block_array`` adds on the device (``rnb_block_sum`` with accumulate) when both sides are device resident
  with the same grid and blockshape.

Host ndarrays keep the reference's NumPy behaviour.
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
    def value(self):
        """The accumulated array as it is held (ndarray or BlockArray); ``ValueError`` before the first ``+=``."""
        if not self.isinit:
            raise ValueError("array not initialized")
        return self._array

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    def __getstate__(self):
        return {"init": self.isinit, "array": self._array}

    def __setstate__(self, d):
        self._array = d["array"]

    def __array__(self, dtype=None, copy=None):
        if not self.isinit:
            raise ValueError("array not initialized")
        out = np.asarray(self._array)
        return out if dtype is None else out.astype(dtype)

    # ---- device path ------------------------------------------------------------------
    def iadd_tensordot(self, a, b, axes=2):
        """``self += tensordot(a, b, axes)`` for BlockArray operands as ONE grouped launch (``accumulate=1`` once the
        accumulator holds a result): the reference's ``commutative`` task."""
        from ..engine import plan as _plan
        from ..engine import runtime
        from .block import BlockArray, _local_tensordot

        if not (isinstance(a, BlockArray) and isinstance(b, BlockArray)):
            raise TypeError("iadd_tensordot takes BlockArray operands")
        if not self.isinit:
            self._array = _local_tensordot(a, b, axes)
            return self
        acc = self._array
        dtype = np.result_type(a.dtype, b.dtype)
        p = _plan.plan_tensordot(a.grid, a.blockshape, b.grid, b.blockshape, dtype, axes)
        if not (isinstance(acc, BlockArray) and acc.grid == p.out_grid and acc.blockshape == p.out_blockshape and acc.dtype == dtype):
            raise ValueError("accumulator layout %r does not match the contraction result (grid %r, blockshape %r, %s)" % (acc, p.out_grid, p.out_blockshape, dtype))
        if not acc.is_device:
            acc = self._array = acc.to_device()
        runtime.execute_tensordot(p, a._device_storage(dtype), a.blockshape, b._device_storage(dtype), b.blockshape,
                                  out_storage=acc._storage, accumulate=True)
        acc._host_mirror = None
        return self

    def _iadd_blockarray(self, other):
        from ..engine import lib as _lib
        from .block import BlockArray
        from .device import _torch, current_stream_ptr

        acc = self._array
        if not (acc.grid == other.grid and acc.blockshape == other.blockshape and acc.dtype == other.dtype):
            raise ValueError("block layouts differ: %r += %r" % (acc, other))
        if not acc.is_device:
            acc = self._array = acc.to_device()
        src = other._device_storage()
        torch = _torch()
        with torch.cuda.device(acc._storage.device):
            _lib.block_sum(acc._storage.ptr, [src.ptr], acc.nblock * acc.blocksize, _lib.DTYPE_CODES[str(acc.dtype)], True,
                           current_stream_ptr(acc._storage.device))
        acc._host_mirror = None
        return self

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if ufunc.nin > 2 or method != "__call__":
            return NotImplemented
        out = kwargs.get("out")
        inplace = out is not None and any(o is self for o in (out if isinstance(out, tuple) else (out,)))
        others = [x for x in inputs if x is not self]
        from .block import BlockArray

        blocky = any(isinstance(x, BlockArray) for x in others) or isinstance(self._array, BlockArray)
        if not self.isinit:
            other = others[0] if others else None
            if inplace:
                # adopt: BlockArrays by reference (the producer hands its result over), ndarrays by copy
                self._array = other if isinstance(other, BlockArray) else np.array(other, copy=True)
                return self
            return other
        if blocky:
            other = others[0] if others else self._array
            if ufunc is np.add and inplace and isinstance(other, BlockArray) and isinstance(self._array, BlockArray):
                return self._iadd_blockarray(other)
            raise NotImplementedError("MaybeArray holding a BlockArray supports `+=` with a BlockArray of the same layout and iadd_tensordot")
        args = [self._array if x is self else x for x in inputs]
        if inplace:
            kwargs = dict(kwargs)
            kwargs["out"] = (self._array,)
            ufunc(*args, **kwargs)
            return self
        return ufunc(*args, **kwargs)


def commutative(res: MaybeArray, a, b, axes=2):
    """``res += np.tensordot(a, b, axes)`` (the task of ``rosnet/array/compss/task/tensordot.py:45-49``): one grouped
    launch accumulating in the kernel for BlockArray operands, NumPy otherwise."""
    from .block import BlockArray

    if isinstance(a, BlockArray) and isinstance(b, BlockArray):
        return res.iadd_tensordot(a, b, axes)
    res += np.tensordot(a, b, axes)
    return res
