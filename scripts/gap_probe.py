"""Where does a Sycamore slice's time go in STEADY STATE?  (bench.py's per-family split comes from one step
run at the end of the process on a cooled-down GPU.)  Phases, back to back on a hot GPU:
  A  20 graph-replayed slices through contract_network (host in, host out)
  B  10 eager slices (graph=False)
  C  6 consecutive eager slices with every C-ABI launch bracketed by events (bench.profile_tn_step), family sums per slice
Usage: python scripts/gap_probe.py [workload]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import rosnet_b200 as rn
from rosnet_b200 import tn as T
from rosnet_b200.engine import lib

wl = sys.argv[1] if len(sys.argv) > 1 else "bench_sycamore"
spec = bench.TN_WORKLOADS[wl]
net, path, sliced = bench.build_tn(spec)
lib.load()
n_slices = int(np.prod([net.sizes[i] for i in sliced]))
loop = spec["mode"] == "loop" and bool(sliced)


def run(k, graph=True):
    ids = [i % n_slices for i in range(k)]
    if loop:
        return T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=ids, graph=None if graph else False)[0]
    outs = [T.contract_network(net, path, sliced=sliced, slice_mode="blocks", graph=graph)[0] for _ in range(k)]
    return sum(np.asarray(o) for o in outs)


def timed(fn, k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


sampler = bench.ClockSampler(0)
run(5)
sampler.start()
print("A graph   ms/step", round(timed(lambda: run(20), 20), 3), flush=True)
print("A graph   ms/step", round(timed(lambda: run(20), 20), 3), flush=True)
print("clocks A", sampler.stop(), flush=True)
sampler = bench.ClockSampler(0); sampler.start()
print("B eager   ms/step", round(timed(lambda: run(10, graph=False), 10), 3), flush=True)
print("clocks B", sampler.stop(), flush=True)
for rep in range(6):
    run(3)   # keep the chip hot between the profiled steps
    fams = bench.profile_tn_step(lambda: run(1, graph=False))
    tot = sum(f["ms"] for f in fams.values())
    print("C step %d  sum %.3f ms  " % (rep, tot) + "  ".join("%s=%.3f(%d)" % (k, v["ms"], v["calls"]) for k, v in sorted(fams.items())), flush=True)
# D: graph replay only, no upload / download between replays
from rosnet_b200.tn import contract as C
gp = C._graphed_pass(net, path, sliced, "loop" if loop else "blocks")
if gp is not None and gp.graph is not None:
    def replays():
        for _ in range(20):
            gp.replay()
    print("D replay-only ms/step", round(timed(replays, 20), 3), flush=True)
