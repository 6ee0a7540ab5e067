"""A few gate applications on a 2^24-element complex64 state (the apply route) for ncu: python scripts/apply_one.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rosnet_b200 as rn

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.default_rng(0)
nq = 24
state = (rng.standard_normal((2,) * nq) + 1j * rng.standard_normal((2,) * nq)).astype("complex64")
gate = (rng.standard_normal((2, 2, 2, 2)) + 1j * rng.standard_normal((2, 2, 2, 2))).astype("complex64")
S = rn.array(state, blockshape=state.shape, inner="cuda")
G = rn.array(gate, blockshape=gate.shape, inner="cuda")
for qubits in [(3, 17), (22, 23), (0, 1), (10, 11)][:reps]:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = np.tensordot(G, S, axes=[(2, 3), qubits])
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        out = np.tensordot(G, S, axes=[(2, 3), qubits])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("qubits", qubits, "ms %.3f" % ms, "GB/s %.0f" % (2 * state.nbytes / ms / 1e6), flush=True)
