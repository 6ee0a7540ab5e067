"""Array types of the contraction path: ``BlockArray`` (grid of equal blocks, host ndarrays or views of
one device allocation) and ``MaybeArray`` (accumulator that is empty until the first ``+=``).  The
reference's third type, ``COMPSsArray`` (a future on a COMPSs worker), has no counterpart: results here
are device buffers with work pending on a CUDA stream."""
from .block import BlockArray  # noqa: F401
from .device import DeviceBlock, DeviceStorage  # noqa: F401
from .maybe import MaybeArray  # noqa: F401
