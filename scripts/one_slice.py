"""One tensor-network step, once (for ncu captures):  python scripts/one_slice.py bench_sycamore [n_passes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from rosnet_b200 import tn as T  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "bench_sycamore"
spec = bench.TN_WORKLOADS[name]
net, path, sliced = bench.build_tn(spec)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    if spec["mode"] == "loop" and sliced:
        r = T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=[0])[0]
    else:
        r = np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="blocks")[0])
torch.cuda.synchronize()
print(name, r)
