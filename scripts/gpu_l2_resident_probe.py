#!/usr/bin/env python
"""How much of the tf32 GEMM's time is the cost of its DRAM re-reads?  The same grouped launch (4096 output blocks of
256 x 256, k = 4096, float32, one tile per block) with (a) every output block pairing its own a / b blocks (32 a-blocks x
128 b-blocks: ~4.3 GB of hi/lo planes streamed from HBM, re-read across waves) and (b) every output block pairing block 0 of both
operands (the planes of two 4 MB blocks: everything after the first touch is an L2 hit).  Same tensor work, same L2 -> SM traffic."""
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import rosnet_b200 as rn
    from rosnet_b200.array.device import DeviceStorage
    from rosnet_b200.engine import runtime
    from rosnet_b200.engine.plan import plan_tensordot

    ga, gb, bm, bk = 32, 128, 256, 4096
    ta = torch.randn(ga * bm * bk, dtype=torch.float32, device="cuda")
    tb = torch.randn(gb * bm * bk, dtype=torch.float32, device="cuda")

    def storage(t):
        st = DeviceStorage.__new__(DeviceStorage)
        st.tensor = t.view(torch.uint8)
        st.nbytes = st.tensor.numel()
        st.device = t.device
        return st

    sa, sb = storage(ta), storage(tb)
    plan = plan_tensordot((ga, 1), (bm, bk), (gb, 1), (bm, bk), "float32", [(1,), (1,)])
    same = dataclasses.replace(plan, pair_a=np.zeros_like(plan.pair_a), pair_b=np.zeros_like(plan.pair_b))
    out = DeviceStorage(plan.n_out * plan.out_block_elems * 4)
    for name, p in (("distinct blocks (HBM streamed)", plan), ("all outputs pair block 0 (L2 resident)", same), ("distinct blocks (HBM streamed)", plan),
                    ("all outputs pair block 0 (L2 resident)", same)):
        for _ in range(2):
            runtime.execute_tensordot(p, sa, (bm, bk), sb, (bm, bk), out_storage=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            runtime.execute_tensordot(p, sa, (bm, bk), sb, (bm, bk), out_storage=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"case": name, "ms": ms, "tflops": plan.flops / ms / 1e9, "n_out": plan.n_out, "m": plan.m, "n": plan.n, "k": plan.k}))


if __name__ == "__main__":
    main()
