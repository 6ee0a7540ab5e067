"""``np.linalg.*`` dispatch names (reference ``rosnet/dispatch/linalg.py:4-151``).

The factorizations are outside the contraction hot path (out of
scope); the names exist so ``np.linalg.<f>(BlockArray)`` reaches a dispatcher and reports
``NotImplementedError`` exactly as the reference does for unregistered types."""
from ._registry import Dispatcher

_NAMES = [
    "multi_dot", "matrix_power", "cholesky", "qr", "svd", "eig", "eigh", "eigvals", "eigvalsh",
    "norm", "cond", "det", "matrix_rank", "slogdet", "solve", "tensorsolve", "lstsq", "inv",
    "pinv", "tensorinv",
]
for _n in _NAMES:
    globals()[_n] = Dispatcher(_n)
del _n
