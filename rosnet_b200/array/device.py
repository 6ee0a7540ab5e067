"""Device-resident block storage.

One contiguous HBM allocation per BlockArray: blocks laid out block-major in row-major grid
order, every block C-contiguous — the per-block layout the reference has for ndarray blocks
(``rosnet/array/block/__init__.py:30-94``), minus the per-block Python objects.  PyTorch is used
only to allocate device / pinned memory and to name streams.

``DeviceBlock`` is the view object put into ``BlockArray.data`` so the reference's "object array
of blocks" face keeps working; it fulfils the ``Array`` protocol, and ``np.asarray(block)``
is a synchronous device-to-host read (the analogue of ``compss_wait_on``,
``rosnet/array/compss/__init__.py:259-262``).
"""
from math import prod

import numpy as np

from ..engine import lib as _lib


def _torch():
    import torch

    return torch


def current_stream_ptr(device=None) -> int:
    return _torch().cuda.current_stream(device).cuda_stream


_cuda_ok = False


def require_cuda():
    global _cuda_ok
    if _cuda_ok:
        return
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.RosnetB200Error("rosnet_b200 needs a CUDA device: there is no CPU fallback for the contraction path")
    _lib.load()
    _cuda_ok = True


class DeviceStorage:
    """A flat device buffer of ``nbytes`` bytes (torch uint8 tensor, 256-byte aligned)."""

    __slots__ = ("tensor", "nbytes", "device")

    def __init__(self, nbytes: int, device=None, zero=False):
        require_cuda()
        torch = _torch()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        alloc = torch.zeros if zero else torch.empty
        self.tensor = alloc(max(int(nbytes), 16), dtype=torch.uint8, device=device)
        self.nbytes = int(nbytes)
        self.device = device

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @classmethod
    def view(cls, parent: "DeviceStorage", offset: int, nbytes: int) -> "DeviceStorage":
        """A window of ``parent`` (shares its allocation and keeps it alive)."""
        self = cls.__new__(cls)
        self.tensor = parent.tensor[int(offset): int(offset) + max(int(nbytes), 16)]
        self.nbytes = int(nbytes)
        self.device = parent.device
        return self


def pinned_empty(shape, dtype) -> np.ndarray:
    """Page-locked host ndarray (so D2H/H2D run at full PCIe speed and asynchronously)."""
    torch = _torch()
    nbytes = max(int(prod(shape)) * np.dtype(dtype).itemsize, 1)
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    arr = t.numpy()[: int(prod(shape)) * np.dtype(dtype).itemsize].view(dtype).reshape(shape)
    return arr


class DeviceBlock:
    """View of block number ``index`` inside a ``DeviceStorage`` (fulfils the Array protocol)."""

    __slots__ = ("storage", "index", "_shape", "_dtype")

    def __init__(self, storage: DeviceStorage, index: int, shape, dtype):
        self.storage = storage
        self.index = int(index)
        self._shape = shape if type(shape) is tuple else tuple(int(s) for s in shape)
        self._dtype = dtype if type(dtype) is np.dtype else np.dtype(dtype)

    shape = property(lambda self: self._shape)
    dtype = property(lambda self: self._dtype)
    ndim = property(lambda self: len(self._shape))
    size = property(lambda self: prod(self._shape))
    itemsize = property(lambda self: self._dtype.itemsize)
    nbytes = property(lambda self: prod(self._shape) * self._dtype.itemsize)

    @property
    def ptr(self) -> int:
        return self.storage.ptr + self.index * self.nbytes

    def __array__(self, dtype=None, copy=None):
        out = np.empty(self._shape, dtype=self._dtype)
        torch = _torch()
        with torch.cuda.device(self.storage.device):
            stream = current_stream_ptr(self.storage.device)
            if out.nbytes:
                _lib.memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, stream)
            _lib.stream_synchronize(stream)
        return out if dtype is None else out.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        return NotImplemented

    def __array_function__(self, func, types, args, kwargs):
        return NotImplemented

    def __repr__(self):
        return f"DeviceBlock(shape={self._shape}, dtype={self._dtype}, device={self.storage.device}, index={self.index})"
