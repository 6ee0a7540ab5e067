"""Where does a TN slice spend its time?  Wraps the two C-ABI launch calls in CUDA events and
reports, per workload: wall time of one slice, GPU time per kernel family, the slowest calls with
their shapes, and the host-side profile (cProfile) of an untimed second slice.

    python scripts/tn_profile.py bench_sycamore bench_random_tn > gpurun_out/tn_profile.txt
"""
import cProfile
import io
import os
import pstats
import sys
import time
from math import prod

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from rosnet_b200 import tn as T  # noqa: E402
from rosnet_b200.engine import lib  # noqa: E402


def main():
    names = sys.argv[1:] or ["bench_sycamore", "bench_random_tn"]
    lib.load()
    orig_permute, orig_contract = lib.permute_launch, lib.contract_grouped
    for name in names:
        spec = bench.TN_WORKLOADS[name]
        t0 = time.perf_counter()
        net, path, sliced = bench.build_tn(spec)
        print(f"== {name}: build_tn {time.perf_counter() - t0:.2f} s, sliced={len(sliced)}")
        loop = spec["mode"] == "loop" and bool(sliced)

        def one():
            if loop:
                return T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=[0])[0]
            return np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="blocks")[0])

        one()
        torch.cuda.synchronize()
        # pass 1: plain wall time
        t0 = time.perf_counter()
        one()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        # pass 2: event-wrapped calls
        calls = []

        def permute(d, src, dst, stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_permute(d, src, dst, stream)
            e1.record()
            calls.append(("permute", e0, e1, dict(bytes=2 * d.elem_bytes * prod(int(d.extent[i]) for i in range(d.rank)), rank=d.rank)))

        def contract(d, stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_contract(d, stream)
            e1.record()
            per = 8 if d.dtype in (lib.RNB_C64, lib.RNB_C128) else 2
            calls.append(("contract", e0, e1, dict(m=d.m, n=d.n, k=d.k, n_out=d.n_out, n_pairs=d.n_pairs, am=d.a_major, bm=d.b_major,
                                                   flops=per * d.m * d.n * d.k * d.n_out * d.n_pairs, dtype=d.dtype)))

        lib.permute_launch, lib.contract_grouped = permute, contract
        t0 = time.perf_counter()
        one()
        torch.cuda.synchronize()
        wall_ev = time.perf_counter() - t0
        lib.permute_launch, lib.contract_grouped = orig_permute, orig_contract
        rows = [(k, e0.elapsed_time(e1), info) for k, e0, e1, info in calls]
        for fam in ("permute", "contract"):
            sel = [r for r in rows if r[0] == fam]
            print(f"{fam}: {len(sel)} calls, GPU {sum(r[1] for r in sel):.2f} ms")
        print(f"wall per slice {wall * 1e3:.1f} ms (with events {wall_ev * 1e3:.1f} ms); GPU sum {sum(r[1] for r in rows):.1f} ms")
        print("slowest 25 calls:")
        for k, ms, info in sorted(rows, key=lambda r: -r[1])[:25]:
            if k == "contract":
                tf = info["flops"] / (ms * 1e-3) / 1e12
                print(f"  contract {ms:8.3f} ms  m={info['m']} n={info['n']} k={info['k']} n_out={info['n_out']} pairs={info['n_pairs']} "
                      f"major={info['am']}{info['bm']}  {tf:.1f} TFLOP/s")
            else:
                print(f"  permute  {ms:8.3f} ms  {info['bytes'] / 1e6:.1f} MB rank={info['rank']}  {info['bytes'] / (ms * 1e-3) / 1e9:.0f} GB/s")
        # histogram of small calls
        small = [r for r in rows if r[1] < 0.05]
        print(f"calls under 50 us: {len(small)} totalling {sum(r[1] for r in small):.2f} ms")
        # pass 3: host profile
        pr = cProfile.Profile()
        pr.enable()
        one()
        torch.cuda.synchronize()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
        print(s.getvalue()[:6000])
        sys.stdout.flush()


if __name__ == "__main__":
    main()
