"""``rosnet.tuning``: cost models (``flops``, ``mem``) and B200 block sizing.

The reference's ``autotune`` wrapped every task body as a pyCOMPSs task with a
``computing_units`` constraint (``rosnet/tuning/task.py:15-81``; the constraint always came out
as "1").  With the COMPSs task graph replaced by grouped CUDA launches, ``autotune`` is kept as a
transparent decorator so code that applies it keeps working."""
from . import flops, mem  # noqa: F401
from .blocking import choose_blockshape, configure, threshold_k  # noqa: F401


def autotune(*args, **task_info):
    """Transparent stand-in for the COMPSs task wrapper: returns the function unchanged.
    Usable bare (``@autotune``) or with task parameters (``@autotune(a=IN, returns=1)``)."""
    if len(args) == 1 and callable(args[0]) and not task_info:
        return args[0]

    def deco(fn):
        return fn

    return deco


def tune(fn, *args, **kwargs):
    """Resource constraints for one call.  One stream-ordered launch per operation: always one
    computing unit (the reference's formula also always evaluated to "1", ``task.py:79``)."""
    return {"computing_units": "1"}
