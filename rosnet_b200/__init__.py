"""rosnet_b200 — B200-native drop-in for rosnet's blocked-tensor contraction path.

Same names as ``rosnet/__init__.py:1-16``: ``BlockArray``, the dispatch functions
(``tensordot``, ``einsum``, ``transpose``, ``reshape``, ``to_numpy``, every NumPy ufunc), the
constructors ``zeros / ones / full / rand`` (plus ``array``, which the reference's benchmarks call
but v0.3.0 only has commented out), ``linalg`` and ``tuning``.  ``install_as("rosnet")`` makes
``import rosnet`` / ``backend="rosnet"`` resolve to this package.
"""
__version__ = "0.1.0"

import sys as _sys

from .dispatch import *  # noqa: F401,F403
from .dispatch import block, count_nonzero, cumsum, einsum, empty_like, full_like, ones_like, reshape, split, stack, tensordot, to_numpy, transpose, zeros_like  # noqa: F401
from . import dispatch  # noqa: F401
from .dispatch import linalg  # noqa: F401
from .array.block import BlockArray, array, full, ones, rand, zeros  # noqa: F401
from .array.maybe import MaybeArray, commutative  # noqa: F401
from .array.sharded import ShardedBlockArray, distributed_tensordot  # noqa: F401
from .engine import scheduler  # noqa: F401
from .array.device import DeviceBlock  # noqa: F401
from . import tuning  # noqa: F401
from . import extra  # noqa: F401
from .engine.lib import RosnetB200Error  # noqa: F401
from . import tn  # noqa: F401,E402  (registers the BlockArray overload of einsum: np.einsum must work after a bare `import rosnet_b200`)


def install_as(name: str = "rosnet"):
    """Register this package under another module name so by-name backend lookups
    (opt_einsum / quimb / autoray ``backend="rosnet"``) and ``import rosnet`` land here."""
    me = _sys.modules[__name__]
    _sys.modules[name] = me
    for sub in ("dispatch", "tuning", "array", "core", "engine"):
        full_name = f"{__name__}.{sub}"
        if full_name in _sys.modules:
            _sys.modules[f"{name}.{sub}"] = _sys.modules[full_name]
    for sub in ("flops", "mem"):
        _sys.modules[f"{name}.tuning.{sub}"] = _sys.modules[f"{__name__}.tuning.{sub}"]
    return me
