"""Work model behind block sizing and the bench's flop counts.

Same entry points as the reference's ``rosnet/tuning/flops.py`` (resolved by module path, the way
``autoray.do(..., like="rosnet.tuning.flops")`` does in ``test/tuning/test_flops.py:8-16``) and the
same unit: *multiply-adds*, computed from shapes only.  ``algorithmic_flops`` converts to the figure
the rooflines use (2 flop per real multiply-add, 8 per complex one)."""
import numpy as np


def _gemm_extents(a_shape, b_shape, axes):
    """(M, N, K) of the matrix product a blocked/dense tensordot is: M, N = free extents of a, b; K = contracted."""
    ax_a, ax_b = (tuple(int(x) for x in side) for side in axes)
    k = 1
    for i, j in zip(ax_a, ax_b):
        if a_shape[i] != b_shape[j]:
            raise AssertionError(f"contracted extents differ: a.shape[{i}]={a_shape[i]} vs b.shape[{j}]={b_shape[j]}")
        k *= int(a_shape[i])
    m = 1
    for pos, extent in enumerate(a_shape):
        if pos not in ax_a:
            m *= int(extent)
    n = 1
    for pos, extent in enumerate(b_shape):
        if pos not in ax_b:
            n *= int(extent)
    return m, n, k


def tensordot(a, b, axes) -> int:
    m, n, k = _gemm_extents(tuple(a.shape), tuple(b.shape), axes)
    return m * n * k


def full(shape, fill_value, dtype=None, **kwargs) -> int:
    """One write per element."""
    count = 1
    for extent in shape:
        count *= int(extent)
    return count


def reshape(a, shape, **kwargs) -> int:
    """Metadata only for a dense block (a per-block relayout on the device is accounted as a transpose)."""
    return 1


def transpose(a, axes=None, **kwargs) -> int:
    """Every element is moved once."""
    return int(a.size)


def algorithmic_flops(a, b, axes, dtype=None) -> int:
    """Flops of ``tensordot(a, b, axes)`` as the rooflines count them: 2 per real multiply-add, 8 per complex."""
    kind = np.dtype(dtype if dtype is not None else np.result_type(a.dtype, b.dtype)).kind
    return (8 if kind == "c" else 2) * tensordot(a, b, axes)
