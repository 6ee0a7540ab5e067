"""Time-to-amplitude of Sycamore-53 m=14 with a SEARCHED contraction tree (tn.search_path) against the headline tree
(greedy alpha=0.5, 8 slices): python scripts/searched_path_probe.py [trials] [minimize]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from math import log2

import numpy as np
import torch

import bench
from rosnet_b200 import tn as T
from rosnet_b200.engine import lib

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 128
minimize = sys.argv[2] if len(sys.argv) > 2 else "flops"
wl = os.environ.get("WL", "bench_sycamore")
spec = bench.TN_WORKLOADS[wl]
net, path0, sliced0 = bench.build_tn(spec)
lib.load()
t0 = time.time()
itemsize = np.dtype(spec["dtype"]).itemsize
path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=trials, seed=0, target_size=2 ** (spec["target_log2"] or 28), minimize=minimize, itemsize=itemsize, dtype=spec["dtype"])
if len(sys.argv) > 5:     # replay a recorded tree: alpha temperature trial_seed
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=int(sys.argv[5]), alpha=float(sys.argv[3]), temperature=float(sys.argv[4]))
    sliced = []
print("search %.1f s" % (time.time() - t0), {k: v for k, v in info.items() if k != "baseline"}, "sliced", sliced, flush=True)
madds, largest, steps = T.path_cost(net.inputs, net.output, net.sizes, path, sliced)
print("log2 flops %.2f width %.1f steps %d" % (log2(8 * madds), log2(largest), len(steps)), flush=True)


def run(k, graph=True):
    outs = []
    for _ in range(k):
        if sliced:
            outs.append(T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=list(range(8)), graph=None if graph else False)[0])
        else:
            outs.append(T.contract_network(net, path, graph=graph)[0])
    return [np.asarray(o) for o in outs]


def timed(fn, k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3) / k


amp = run(3)[0]
print("amplitude", complex(np.asarray(amp).ravel()[0]), flush=True)
for g in (True, False):
    _, ms = timed(lambda: run(10, graph=g), 10)
    per = 8 if np.dtype(spec["dtype"]).kind == "c" else 2
    nsl = 8 if sliced else 1
    print("graph" if g else "eager", "ms per %s %.3f  -> %.1f TFLOP/s" % ("8 slices" if sliced else "contraction", ms, per * madds * nsl / ms / 1e9), flush=True)
fams = bench.profile_tn_step(lambda: run(1, graph=False))
if sliced:
    print("(sliced tree: %d slices in all; host check skipped)" % int(np.prod([net.sizes[i] for i in sliced])))
    sys.exit(0)
fams.get("permute_split", {}).pop("largest", None)
print("families", {k: (round(v["ms"], 3), v["calls"]) for k, v in fams.items()}, "sum %.3f" % sum(v["ms"] for v in fams.values()), flush=True)
# reference value: the same tree in complex128 on the host
tensors = {i: (np.asarray(a).astype(np.complex128 if np.dtype(spec['dtype']).kind == 'c' else np.float64), tuple(ix)) for i, (a, ix) in enumerate(zip(net.arrays, net.inputs))}
nxt = len(tensors)
t0 = time.time()
for a, b in path:
    (A, ia), (B, ib) = tensors.pop(a), tensors.pop(b)
    sh = [i for i in ia if i in ib]
    C = np.tensordot(A, B, ([ia.index(i) for i in sh], [ib.index(i) for i in sh]))
    tensors[nxt] = (C, tuple(i for i in ia if i not in sh) + tuple(i for i in ib if i not in sh))
    nxt += 1
ref = complex(list(tensors.values())[0][0])
got = complex(np.asarray(amp).ravel()[0])
print("host complex128 amplitude", ref, "in %.1f s; rel diff of the GPU complex64 result %.3e" % (time.time() - t0, abs(got - ref) / abs(ref)), flush=True)
