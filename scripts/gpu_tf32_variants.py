#!/usr/bin/env python
"""B200 check of the single-precision grouped GEMM variants (csrc/gemm_tf32.cu): first-generation kernel (v1), v2 on
single CTAs (coalesced epilogue) and v2 on CTA pairs (tcgen05 cta_group::2).  Parity against numpy in double
precision on aligned / ragged / multi-block / split-K shapes, then throughput on the GEMM shapes of a Sycamore slice.

Every variant runs in its own subprocess under a timeout, so a trap or a hang in one does not take the others down.

    python scripts/gpu_tf32_variants.py [--quick] [--out gpurun_out/tf32_variants.jsonl]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "v1": {"RNB_TF32_KERNEL": "v1"},
    "v2_cg1": {"RNB_TF32_KERNEL": "v2", "RNB_TF32_CG": "1"},
    "v2_cg2": {"RNB_TF32_KERNEL": "v2", "RNB_TF32_CG": "2"},
    "v2_auto": {"RNB_TF32_KERNEL": "v2"},
    "v2_4m": {"RNB_COMPLEX_MODE": "4m"},
    "v2_3m": {"RNB_COMPLEX_MODE": "3m"},
}

# (M, N, K, block of A, block of B): C[M, N] = sum_k A[M, K] B[N, K]
PARITY = [
    (256, 256, 64, (256, 64), (256, 64)),
    (128, 128, 32, (128, 32), (128, 32)),
    (300, 200, 100, (300, 100), (200, 100)),
    (384, 640, 256, (384, 256), (640, 256)),
    (1024, 512, 512, (512, 256), (256, 256)),
    (1000, 520, 1044, (500, 348), (260, 348)),
    (256, 256, 32768, (256, 32768), (256, 32768)),
    (512, 300, 16384, (512, 8192), (300, 8192)),
    (4096, 2048, 1024, (4096, 1024), (2048, 1024)),
]
PERF = [
    # the GEMM shapes of one Sycamore-53 m=14 slice (complex64), largest first
    ("complex64", 8192, 32768, 4096),
    ("complex64", 65536, 2048, 4096),
    ("complex64", 512, 2048, 131072),
    ("complex64", 1024, 32768, 4096),
    ("complex64", 512, 262144, 128),
    ("complex64", 4096, 65536, 16),
    ("complex64", 4096, 4096, 4096),
    ("float32", 4096, 4096, 4096),
    ("float32", 8192, 8192, 8192),
]


def worker(args):
    import numpy as np
    import torch

    import rosnet_b200 as rn
    from rosnet_b200.engine import lib, runtime

    if args.precision == "tf32_bf16x2":
        runtime.settings["precision"] = lib.RNB_PREC_TF32_BF16X2
    rng = np.random.default_rng(0)
    out = []

    def emit(rec):
        rec["variant"] = args.worker
        rec["precision"] = args.precision
        print("JSON " + json.dumps(rec), flush=True)

    def rand(shape, dtype):
        x = rng.standard_normal(shape)
        if np.dtype(dtype).kind == "c":
            x = x + 1j * rng.standard_normal(shape)
        return x.astype(dtype)

    os.environ["RNB_SIMT"] = "0"
    lib.reload_env()
    if not args.perf_only:
        for dtype in ("complex64", "float32"):
            for (M, N, K, ba, bb) in PARITY:
                a, b = rand((M, K), dtype), rand((N, K), dtype)
                A = rn.array(a, blockshape=ba, inner="cuda")
                B = rn.array(b, blockshape=bb, inner="cuda")
                got = np.asarray(np.tensordot(A, B, [(1,), (1,)]))
                want = np.tensordot(a.astype("complex128" if np.dtype(dtype).kind == "c" else "float64"),
                                    b.astype("complex128" if np.dtype(dtype).kind == "c" else "float64"), [(1,), (1,)])
                err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
                emit({"kind": "parity", "dtype": dtype, "M": M, "N": N, "K": K, "rel_frobenius": err, "ok": bool(err <= 1e-5)})
        # accumulate = 1 through the runtime (C += A.B)
        from rosnet_b200.engine.plan import plan_tensordot

        a, b = rand((512, 256), "complex64"), rand((384, 256), "complex64")
        A = rn.array(a, blockshape=(256, 256), inner="cuda")
        B = rn.array(b, blockshape=(384, 256), inner="cuda")
        plan = plan_tensordot(A.grid, A.blockshape, B.grid, B.blockshape, "complex64", [(1,), (1,)])
        st = runtime.execute_tensordot(plan, A._storage, A.blockshape, B._storage, B.blockshape)
        runtime.execute_tensordot(plan, A._storage, A.blockshape, B._storage, B.blockshape, out_storage=st, accumulate=True)
        got = np.asarray(rn.BlockArray._from_storage(st, plan.out_grid, plan.out_blockshape, "complex64"))
        want = 2 * np.tensordot(a.astype("complex128"), b.astype("complex128"), [(1,), (1,)])
        err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        emit({"kind": "parity", "dtype": "complex64", "case": "accumulate", "rel_frobenius": err, "ok": bool(err <= 1e-5)})

    if args.parity_only:
        return
    for (dtype, M, N, K) in PERF:
        if args.quick and M * N * K > 2 ** 41:
            continue
        ta = torch.randn(M, K, dtype=torch.complex64 if dtype == "complex64" else torch.float32, device="cuda")
        tb = torch.randn(N, K, dtype=ta.dtype, device="cuda")
        from rosnet_b200.array.device import DeviceStorage

        def wrap(t, shape):
            st = DeviceStorage.__new__(DeviceStorage)
            st.tensor = (torch.view_as_real(t) if t.is_complex() else t).reshape(-1).view(torch.uint8)
            st.nbytes = st.tensor.numel()
            st.device = t.device
            return rn.BlockArray._from_storage(st, (1, 1), shape, dtype)

        A, B = wrap(ta, (M, K)), wrap(tb, (N, K))
        for _ in range(2):
            C = np.tensordot(A, B, [(1,), (1,)])
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            C = np.tensordot(A, B, [(1,), (1,)])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = (8 if dtype == "complex64" else 2) * M * N * K
        # spot parity on a corner of the result
        c = C._storage.tensor[: M * N * ta.element_size()].view(ta.dtype).reshape(M, N)
        ref = (ta[:64].to(torch.complex128 if dtype == "complex64" else torch.float64) @ tb[:64].to(torch.complex128 if dtype == "complex64" else torch.float64).T)
        err = float(torch.linalg.norm(c[:64, :64].to(ref.dtype) - ref) / torch.linalg.norm(ref))
        emit({"kind": "perf", "dtype": dtype, "M": M, "N": N, "K": K, "ms": ms, "tflops": flops / ms / 1e9, "corner_rel_err": err})
        del A, B, C, ta, tb, c, ref
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--parity-only", action="store_true")
    ap.add_argument("--perf-only", action="store_true")
    ap.add_argument("--precision", default="default")
    ap.add_argument("--variants", default="v1,v2_cg1,v2_cg2,v2_auto")
    ap.add_argument("--precisions", default="default,tf32_bf16x2")
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tf32_variants.jsonl"))
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "a") as fout:
        for prec in args.precisions.split(","):
            for name in args.variants.split(","):
                env = dict(os.environ)
                env.update(VARIANTS[name])
                cmd = [sys.executable, os.path.abspath(__file__), "--worker", name, "--precision", prec]
                for flag in ("quick", "parity_only", "perf_only"):
                    if getattr(args, flag):
                        cmd.append("--" + flag.replace("_", "-"))
                t0 = time.time()
                try:
                    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=args.timeout)
                    rc, so, se = proc.returncode, proc.stdout, proc.stderr
                except subprocess.TimeoutExpired as e:
                    rc, so, se = -9, (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
                recs = [json.loads(l[5:]) for l in so.splitlines() if l.startswith("JSON ")]
                for r in recs:
                    fout.write(json.dumps(r) + "\n")
                bad = [r for r in recs if r["kind"] == "parity" and not r["ok"]]
                summary = {"kind": "summary", "variant": name, "precision": prec, "rc": rc, "seconds": round(time.time() - t0, 1),
                           "parity_cases": sum(r["kind"] == "parity" for r in recs), "parity_failed": len(bad),
                           "worst": max([r["rel_frobenius"] for r in recs if r["kind"] == "parity"], default=None),
                           "stderr_tail": se[-600:] if rc != 0 else ""}
                fout.write(json.dumps(summary) + "\n")
                fout.flush()
                print(json.dumps(summary))
                for r in recs:
                    if r["kind"] == "perf":
                        print("   %-10s %6d x %6d x %6d  %8.3f ms  %7.1f TFLOP/s  err %.1e" % (r["dtype"], r["M"], r["N"], r["K"], r["ms"], r["tflops"], r["corner_rel_err"]))
                for r in bad:
                    print("   PARITY FAIL", r)


if __name__ == "__main__":
    main()
