"""One dispatcher per NumPy function name handled by the blocked path
(reference ``rosnet/dispatch/numpy.py:4-68``; unregistered type combinations raise
``NotImplementedError``)."""
from ._registry import Dispatcher

tensordot = Dispatcher("tensordot")
einsum = Dispatcher("einsum")
reshape = Dispatcher("reshape")
transpose = Dispatcher("transpose")
stack = Dispatcher("stack")
split = Dispatcher("split")
block = Dispatcher("block")
zeros_like = Dispatcher("zeros_like")
ones_like = Dispatcher("ones_like")
full_like = Dispatcher("full_like")
empty_like = Dispatcher("empty_like")
cumsum = Dispatcher("cumsum")
count_nonzero = Dispatcher("count_nonzero")
