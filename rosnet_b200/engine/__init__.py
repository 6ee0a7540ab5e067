"""Host runtime under the BlockArray face: ctypes binding of the C-ABI CUDA library, contraction
planning (which block pairs feed which output block, operand layouts), execution on CUDA streams
and the single-box multi-GPU scheduler.  Replaces the pyCOMPSs task layer
(rosnet/tuning/task.py:15-81, rosnet/array/compss/task/tensordot.py)."""
