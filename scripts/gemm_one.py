#!/usr/bin/env python
"""One blocked contraction C[M, N] = sum_k A[M, K] B[N, K] through the public API, repeated: the command ncu wraps to
capture the grouped-GEMM kernel on a chosen shape (and the timing line to compare against).

    python scripts/gemm_one.py --dtype complex64 --m 8192 --n 32768 --k 4096 [--reps 3] [--complex-mode 3m] [--precision ...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="complex64")
    ap.add_argument("--m", type=int, default=8192)
    ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--complex-mode", default=None)
    ap.add_argument("--precision", default=None)
    args = ap.parse_args()

    import numpy as np
    import torch

    import rosnet_b200 as rn
    from rosnet_b200.array.device import DeviceStorage

    tdt = {"complex64": torch.complex64, "float32": torch.float32, "complex128": torch.complex128, "float64": torch.float64}[args.dtype]
    ta = torch.randn(args.m, args.k, dtype=tdt, device="cuda")
    tb = torch.randn(args.n, args.k, dtype=tdt, device="cuda")

    def wrap(t, shape):
        st = DeviceStorage.__new__(DeviceStorage)
        st.tensor = (torch.view_as_real(t) if t.is_complex() else t).reshape(-1).view(torch.uint8)
        st.nbytes = st.tensor.numel()
        st.device = t.device
        return rn.BlockArray._from_storage(st, (1, 1), shape, args.dtype)

    A, B = wrap(ta, (args.m, args.k)), wrap(tb, (args.n, args.k))
    kw = {}
    if args.complex_mode:
        kw["complex_mode"] = args.complex_mode
    if args.precision:
        kw["precision"] = args.precision
    with rn.tuning.configure(**kw):
        for _ in range(args.warmup):
            C = np.tensordot(A, B, [(1,), (1,)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            C = np.tensordot(A, B, [(1,), (1,)])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    flops = (8 if args.dtype.startswith("complex") else 2) * args.m * args.n * args.k
    c = C._storage.tensor[: args.m * args.n * ta.element_size()].view(tdt).reshape(args.m, args.n)
    wide = torch.complex128 if args.dtype.startswith("complex") else torch.float64
    ref = ta[:64].to(wide) @ tb[:64].to(wide).T
    err = float(torch.linalg.norm(c[:64, :64].to(wide) - ref) / torch.linalg.norm(ref))
    print(json.dumps({"dtype": args.dtype, "m": args.m, "n": args.n, "k": args.k, "ms": ms, "tflops": flops / ms / 1e9, "corner_rel_err": err,
                      "complex_mode": args.complex_mode, "precision": args.precision}))


if __name__ == "__main__":
    main()
