#!/usr/bin/env python
"""Benchmark of the blocked-tensor contraction hot path (BASELINE.json metric: contraction TFLOP/s
and fraction of the tensor-core roofline at 1/2/4/8 B200; Sycamore seconds per slice).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]

Default workload: ``bench_sycamore`` = BASELINE.json configs[4], the configuration the metric names
("Sycamore slice s"): the Sycamore-53 m=14 amplitude network, complex64, sliced until the largest
intermediate is 2^28 elements; a "step" is ONE slice contracted through the unchanged rosnet API
(``tn.contract_network`` -> ``tensordot``/``transpose`` on BlockArrays -> C ABI): 258 pairwise blocked
contractions, 1.66e13 flop.  configs[1] (``bench_random_tn``: 200 tensors, bond 16, mean degree 3)
cannot be a bench line as written: its contraction width is 2^156 (SURVEY 6.1 / DESIGN.md); it runs
here with the reference script's own default bond 2.  Other lines: ``--workload bench_contraction``
(configs[0], the reference's CPU-runnable case: the per-kernel rooflines of the FP64 GEMM and of the
permute are measured on it), ``bench_contraction_c64|_f32|_c128|_permute``, ``bench_qft``,
``bench_random_tn``, ``bench_schmidt_matrix`` (configs[3], contracted axis split across ranks).

* ``value``   : whole-job TFLOP/s with the step's input tensors resident in HBM and results left on
                the device; CUDA events + host clock (the larger), barrier + sync both sides, max over ranks.
* ``e2e``     : same metric through the public API starting from HOST ndarrays and ending with the
                result on the host, copies inside the timed region.
* ``roofline``: the dominant kernel (grouped GEMM) against the tensor peak measured in this run with
                cuBLAS (TF32/3 for 3xTF32, FP64 DGEMM for f64/c128; MEASURED_PEAKS.json holds no such
                figure); ``roofline_permute``: the permute kernels against measured HBM bandwidth.
* ``cpu_baseline``: the reference's arithmetic (numpy.tensordot per pairwise step / the oracle's
                blocked tensordot) on this box's host cores, bounded sample.
* N > 1 (torchrun): weak scaling.  TN workloads: independent slices dealt round-robin, no data-path
  collective.  bench_contraction*: every rank contracts its own slab of `a` with a replica of `b`.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape_a, block_a, shape_b, block_b, dtype, axes)
    "bench_contraction": ((64,) * 4, (32,) * 4, (64,) * 4, (32,) * 4, "float64", [(2, 3), (0, 1)]),
    "bench_contraction_permute": ((64,) * 4, (32,) * 4, (64,) * 4, (32,) * 4, "float64", [(0, 2), (1, 3)]),
    "bench_contraction_c64": ((64,) * 4, (32,) * 4, (64,) * 4, (32,) * 4, "complex64", [(2, 3), (0, 1)]),
    "bench_contraction_f32": ((64,) * 4, (32,) * 4, (64,) * 4, (32,) * 4, "float32", [(2, 3), (0, 1)]),
    "bench_contraction_c128": ((64,) * 4, (32,) * 4, (64,) * 4, (32,) * 4, "complex128", [(2, 3), (0, 1)]),
    "bench_schmidt_matrix": ((2**14, 2**14), (2**11, 2**11), (2**14, 2**14), (2**11, 2**11), "complex128", [(1,), (1,)]),
    "bench_schmidt_small": ((2**13, 2**13), (2**11, 2**11), (2**13, 2**13), (2**11, 2**11), "complex128", [(1,), (1,)]),
}


TN_WORKLOADS = {
    # BASELINE.json configs[1], [2], [4]
    # bond 2, not 16: with 300 bonds of extent 16 dealt at random, single INPUT tensors already reach 16^8 elements and
    # the contraction width is 2^156 (tests/test_tn_host.py); extent 2 is the reference script's own default (d_max=2)
    # tree_candidates: (alpha, temperature, trial_seed) of trees recorded by longer runs of tn.search_path (256 trials, minimize="time"),
    # re-evaluated next to a short fresh search by extra.searched_tree
    "bench_random_tn": dict(kind="random", n=200, reg=3, d=2, dtype="complex64", target_log2=27, alpha=1.0, mode="loop",
                            tree_candidates=((1.0228195800287083, 0.0, 89514544),)),
    "bench_qft": dict(kind="qft", n=30, dtype="complex128", target_log2=None, alpha=1.0, mode="blocks",
                      tree_candidates=((1.7307468709426035, 0.0, 117890804),)),
    "bench_sycamore": dict(kind="sycamore", cycles=14, dtype="complex64", target_log2=28, alpha=0.5, mode="loop",
                           # (alpha, temperature, trial_seed) of trees recorded by longer offline runs of tn.search_path (512-6000 trials, minimize="time"),
                           # re-evaluated next to a short fresh search by extra.searched_tree: a balanced tree of tensor-core sized steps (2^37.7 flop,
                           # width 2^24) and two chains of gate applications (2^35.3 / 2^34.7 flop, width 2^26)
                           tree_candidates=((0.61461059008156, 0.02284424442732209, 1783767436), (6.137498636489522, 0.03991954084384389, 915743486),
                                            (6.9072941321626251, 0.0779626728083221, 1253166252))),
}


def make_dense(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(shape, dtype=np.float64)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(shape, dtype=np.float64)
    return x.astype(dtype)


def algorithmic_flops(shape_a, shape_b, dtype, axes):
    k = int(np.prod([shape_a[i] for i in axes[0]]))
    m = int(np.prod(shape_a)) // k
    n = int(np.prod(shape_b)) // k
    return (8 if np.dtype(dtype).kind == "c" else 2) * m * n * k


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 7:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_time(a, ab, b, bb, axes, repeats):
    """The reference's algorithm (oracle restatement: per-pair numpy.tensordot + Python fold,
    rosnet/array/block/__init__.py:281-336) on the host cores.  Returns best seconds."""
    from oracle import blocked_tensordot as orc

    A = orc.split_blocks(a, ab)
    B = orc.split_blocks(b, bb)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.blocked_tensordot(A, B, axes, order="C")
        best = min(best, time.perf_counter() - t0)
    return best


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms use every host core the BLAS can take."""
    try:
        from threadpoolctl import threadpool_limits

        return threadpool_limits(limits=os.cpu_count())
    except Exception:
        import contextlib

        return contextlib.nullcontext()


def run_reference(args, wl):
    """--impl reference: the reference CPU path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with all_host_threads():
        return _run_reference(args, wl)


def _run_reference(args, wl):
    sa, ab, sb, bb, dtype, axes = wl
    a, b = make_dense(sa, dtype, 0), make_dense(sb, dtype, 1)
    flops = algorithmic_flops(sa, sb, dtype, axes)
    from oracle import blocked_tensordot as orc

    A, B = orc.split_blocks(a, ab), orc.split_blocks(b, bb)
    for _ in range(max(args.warmup, 1)):
        orc.blocked_tensordot(A, B, axes, order="C")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.blocked_tensordot(A, B, axes, order="C")
    dt = (time.perf_counter() - t0) / args.steps
    val = flops / dt / 1e12
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "contraction TFLOP/s", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"float64": "f64", "complex128": "c128", "float32": "f32 (3xTF32)", "complex64": "c64 (3xTF32)"}.get(dtype, dtype), "data": "synthetic",
        "config": {"workload": args.workload, "shape_a": list(sa), "shape_b": list(sb), "blockshape_a": list(ab), "blockshape_b": list(bb),
                   "axes": [list(x) for x in axes], "per_gpu_flops": flops},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                         "sample": f"full workload per step, {args.steps} steps, numpy {np.__version__} BLAS threads={cores}, os.cpu_count={os.cpu_count()}"},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))



def build_tn(spec):
    """Network + path + slices of a TN workload; deterministic (every rank builds the same)."""
    from rosnet_b200 import tn as T
    from rosnet_b200.tn.network import simplify

    if spec["kind"] == "random":
        net = T.rand_equation(spec["n"], spec["reg"], 0, spec["d"], spec["d"], seed=0)
        net.random_arrays(spec["dtype"], seed=0)
    elif spec["kind"] == "qft":
        net = simplify(T.qft_amplitude_network(spec["n"], dtype=spec["dtype"]))
    else:
        net = simplify(T.sycamore_amplitude_network(spec["cycles"], seed=0, dtype=spec["dtype"]))
    path = T.greedy_path(net.inputs, net.output, net.sizes, seed=0, alpha=spec["alpha"])
    madds, largest, _ = T.path_cost(net.inputs, net.output, net.sizes, path)
    sliced = []
    if spec["target_log2"] is not None and largest > 2 ** spec["target_log2"]:
        sliced = T.find_slices(net.inputs, net.output, net.sizes, path, 2 ** spec["target_log2"])
    elif spec["mode"] == "blocks":
        sliced = T.find_slices(net.inputs, net.output, net.sizes, path, max(largest // 8, 2))   # slicing-as-blocks: ~8 blocks
    return net, path, sliced


def cpu_tn_sample(net, path, sliced, budget_flops, per_madd):
    """The reference's arithmetic (numpy.tensordot per pairwise step) on the host for the first steps of
    one slice whose cumulative work stays under the budget; returns (flops done, seconds)."""
    arrays, inputs = [], []
    for ix, arr in zip(net.inputs, net.arrays):
        arr = np.asarray(arr)
        keep = []
        for lab in ix:
            if lab in sliced:
                arr = np.take(arr, 0, axis=len(keep))
            else:
                keep.append(lab)
        inputs.append(tuple(keep))
        arrays.append(arr)
    tensors = {i: (a, ix) for i, (a, ix) in enumerate(zip(arrays, inputs))}
    nxt, done, t0 = len(tensors), 0, time.perf_counter()
    for a, b_ in path:
        if a not in tensors or b_ not in tensors:
            break
        (A, ia), (B, ib) = tensors[a], tensors[b_]
        shared = [i for i in ia if i in ib]
        k = int(np.prod([A.shape[ia.index(i)] for i in shared])) if shared else 1
        work = per_madd * (A.size // k) * (B.size // k) * k
        if done + work > budget_flops and done > 0:
            break
        del tensors[a], tensors[b_]
        C = np.tensordot(A, B, ([ia.index(i) for i in shared], [ib.index(i) for i in shared]))
        tensors[nxt] = (C, tuple(i for i in ia if i not in shared) + tuple(i for i in ib if i not in shared))
        nxt += 1
        done += work
    return done, time.perf_counter() - t0


def measure_tensor_peak(single: bool, sustain_s: float = 4.0):
    """cuBLAS GEMM 8192^3 through torch.matmul: FP64 (DGEMM) or fp32 with TF32 tensor cores, measured the way the driver
    measured the bf16 figures of MEASURED_PEAKS.json (which holds no TF32 / FP64 number): best of 10 single launches
    (burst) AND back to back for ``sustain_s`` seconds under the power cap (sustained).  Returns a dict of RAW tensor
    TFLOP/s; the rooflines use the sustained figure and say so."""
    import torch

    n = 8192
    if single:
        torch.backends.cuda.matmul.allow_tf32 = True
    tdt = torch.float32 if single else torch.float64
    x = torch.randn(n, n, dtype=tdt, device="cuda")
    y = torch.randn(n, n, dtype=tdt, device="cuda")
    best = float("inf")
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
    burst = 2 * n**3 / (best * 1e-3) / 1e12
    reps = max(4, int(sustain_s / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(x, y)
    e1.record()
    torch.cuda.synchronize()
    sustained = 2 * n**3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del x, y
    torch.cuda.empty_cache()
    what = "fp32 with TF32 tensor cores" if single else "float64 (DGEMM)"
    return {"burst": burst, "sustained": sustained, "seconds_sustained": e0.elapsed_time(e1) * 1e-3,
            "how": f"measured in this run: torch.matmul {what} {n}^3 (cuBLAS), best of 10 launches (burst) and {reps} launches back to back (sustained); MEASURED_PEAKS.json has no such figure"}


def tensor_work_multiplier(dtype: str, m: int, n: int, k_total: int, precision: str, complex_mode: str = "auto"):
    """Tensor-core products EXECUTED per algorithmic product by csrc/gemm_tf32.cu / gemm_f64.cu (mirror of use_3m / f64_mode):
    float32 3xTF32: 3; complex64 4M: 3; complex64 3M: 3 * 3/4 = 2.25 (9 instead of 12 real TF32 products per complex
    multiply-add); the opt-in TF32 + 2 x BF16 arithmetic: 2 TF32-equivalents; float64: 1; complex128 4M: 1, 3M: 0.75."""
    if dtype == "float64":
        return 1.0
    if dtype == "complex128":   # csrc/gemm_f64.cu: f64_mode - 3M (three DMMA products per complex product instead of four) from k = 64 up
        return 0.75 if (complex_mode == "3m" or (complex_mode == "auto" and k_total >= 64)) else 1.0
    if precision == "tf32_bf16x2":
        return 2.0
    three_m = dtype == "complex64" and (complex_mode == "3m" or (complex_mode == "auto" and k_total >= 1024 and m >= 128 and n >= 128))
    return 2.25 if three_m else 3.0


def profile_tn_step(one):
    """One extra, untimed step with every C-ABI launch bracketed by CUDA events: GPU time per kernel
    family and the flops that went through the tensor-core GEMM (the roofline's dominant kernel)."""
    import torch

    from rosnet_b200.engine import lib

    calls = []
    orig_permute, orig_contract = lib.permute_launch, lib.contract_grouped

    def permute(d, src, dst, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_permute(d, src, dst, stream)
        e1.record()
        n = 1
        for i in range(d.rank):
            n *= int(d.extent[i])
        nbytes = 2 * d.elem_bytes * n
        # launches that move >= 64 MB are HBM-bound; the hundreds of tiny ones are launch-latency bound
        calls.append(("permute" if nbytes >= (64 << 20) else "permute_small", e0, e1, 0, nbytes))   # last field: bytes (permutes) / executed tensor flops (contractions)

    def contract(d, stream):
        route = lib.contract_route(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_contract(d, stream)
        e1.record()
        per = 8 if d.dtype in (lib.RNB_C64, lib.RNB_C128) else 2
        fam = ("gemm", "skinny", "small")[route]
        fl = per * d.m * d.n * d.k * d.n_out * d.n_pairs
        dn = {lib.RNB_F32: "float32", lib.RNB_F64: "float64", lib.RNB_C64: "complex64", lib.RNB_C128: "complex128"}[d.dtype]
        mult = tensor_work_multiplier(dn, d.m, d.n, d.n_pairs * d.k, "tf32_bf16x2" if d.precision == lib.RNB_PREC_TF32_BF16X2 else "default",
                                      {lib.RNB_CPLX_4M: "4m", lib.RNB_CPLX_3M: "3m"}.get(d.complex_mode, "auto")) if route == 0 else 0.0
        calls.append((fam, e0, e1, fl, fl * mult))

    orig_nd = lib.contract_nd
    orig_split = lib.permute_split

    def permute_split(d, src, planes, ws_ptr, rows, k, nblocks, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ok = orig_split(d, src, planes, ws_ptr, rows, k, nblocks, stream)
        e1.record()
        if ok:
            n = 1
            for i in range(d.rank):
                n *= int(d.extent[i])
            # read the operand once, write its hi and lo planes: 2x (same-offset planes), 4x (complex b of 4M), 3x (3M: re, im, re+im)
            written = {lib.RNB_SPLIT_PLANES: 2, lib.RNB_SPLIT_C64_B4M: 4, lib.RNB_SPLIT_C64_3M: 3}[planes.mode]
            nbytes = d.elem_bytes * n * (1 + written)
            calls.append(("permute_split" if nbytes >= (64 << 20) else "permute_split_small", e0, e1, 0, nbytes))
            if nbytes > largest_split.get("bytes", 0):     # remembered for isolated_permute_split (by value: descriptors are cached and re-pointed)
                dd, pp = type(d)(), type(planes)()
                ctypes.memmove(ctypes.byref(dd), ctypes.byref(d), ctypes.sizeof(d))
                ctypes.memmove(ctypes.byref(pp), ctypes.byref(planes), ctypes.sizeof(planes))
                largest_split.update(bytes=nbytes, desc=dd, planes=pp, rows=int(rows), k=int(k), nblocks=int(nblocks), written=written,
                                     src_bytes=d.elem_bytes * n)
        return ok

    def contract_nd(d, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_nd(d, stream)
        e1.record()
        m = int(np.prod([d.ext_m[i] for i in range(d.rank_m)])) if d.rank_m else 1
        n = int(np.prod([d.ext_n[i] for i in range(d.rank_n)])) if d.rank_n else 1
        k = int(np.prod([d.ext_k[i] for i in range(d.rank_k)])) if d.rank_k else 1
        eb = {lib.RNB_F32: 4, lib.RNB_F64: 8, lib.RNB_C64: 8, lib.RNB_C128: 16}[d.dtype]
        nbytes = eb * d.n_out * (d.n_pairs * (m * k + n * k) + m * n)      # both operands read once, the result written once
        # the apply kernel (one tiny operand, csrc/skinny.cu) on a large tensor is HBM-bound; everything else on this entry point is latency-bound
        calls.append(("apply" if nbytes >= (64 << 20) else "small_nd", e0, e1, 0, nbytes))

    orig_batch = lib.contract_nd_batch

    def contract_nd_batch(descs, stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_batch(descs, stream)
        e1.record()
        calls.append(("small_batch", e0, e1, 0, 0))
        batch_sizes.append(len(descs))

    batch_sizes = []
    largest_split = {}
    lib.permute_launch, lib.contract_grouped, lib.contract_nd, lib.permute_split, lib.contract_nd_batch = permute, contract, contract_nd, permute_split, contract_nd_batch
    try:
        one()
        torch.cuda.synchronize()
    finally:
        lib.permute_launch, lib.contract_grouped, lib.contract_nd, lib.permute_split, lib.contract_nd_batch = orig_permute, orig_contract, orig_nd, orig_split, orig_batch
    fams = {}
    for fam, e0, e1, flops, extra in calls:
        f = fams.setdefault(fam, {"calls": 0, "ms": 0.0, "flops": 0, "bytes": 0, "tensor_flops": 0.0})
        f["calls"] += 1
        f["ms"] += e0.elapsed_time(e1)
        f["flops"] += flops
        if fam.startswith("permute") or fam in ("apply", "small_nd"):
            f["bytes"] += extra
        else:
            f["tensor_flops"] += extra
    if batch_sizes:
        fams["small_batch"]["tasks"] = int(sum(batch_sizes))
    if largest_split and "permute_split" in fams:
        fams["permute_split"]["largest"] = largest_split
    return fams


def isolated_permute_split(rec, reps=10):
    """The step's largest fused matricize + split launch again, ALONE: same descriptor (extents, strides, plane mode) on fresh
    buffers, after the GPU has idled for a second, `reps` launches back to back.  Inside a slice the same launch starts on
    the heels of a power-capped GEMM and runs at that GEMM's throttled SM clock; this is the kernel's own rate."""
    import torch

    from rosnet_b200.array.device import current_stream_ptr

    d, planes = rec["desc"], rec["planes"]
    plane_bytes = rec["src_bytes"] * rec["written"] // 2
    src = torch.empty(rec["src_bytes"], dtype=torch.uint8, device="cuda")
    src.view(torch.float32).normal_()
    ws = torch.empty(max(planes.hi_offset, planes.lo_offset) + plane_bytes + 4096, dtype=torch.uint8, device="cuda")
    from rosnet_b200.engine import lib

    stream = current_stream_ptr()
    for _ in range(2):
        lib.permute_split(d, src.data_ptr(), planes, ws.data_ptr(), rec["rows"], rec["k"], rec["nblocks"], stream)
    torch.cuda.synchronize()
    time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.permute_split(d, src.data_ptr(), planes, ws.data_ptr(), rec["rows"], rec["k"], rec["nblocks"], stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    del src, ws
    return {"ms": ms, "bytes": rec["bytes"], "GBps": rec["bytes"] / (ms * 1e-3) / 1e9, "launches": reps}


def searched_tree_case(net, spec, headline_ms_per_step, headline_steps, headline_path, headline_sliced, with_host=True, reps=20):
    """Time to the WHOLE result of the workload's network along a SEARCHED contraction tree (tn.search_path, objective: the two-term
    roofline model of tn.step_time_model at the dtype's rates; 64 fresh seeded greedy trees plus the trees recorded in the workload
    spec) next to the headline tree: same network, same engine, host tensors in / host result out.  Sycamore-53 m=14: the searched
    trees need 10^10.5..10^11.4 flop unsliced where the headline tree spends 8 slices of 10^13.2.  Checked against the headline
    tree's result where that is affordable (<= 64 headline steps) and against the same tree in double precision on the host (unsliced trees)."""
    from math import log2

    import torch

    from rosnet_b200 import tn as T

    t0 = time.perf_counter()
    target = 2 ** (spec["target_log2"] or 28)
    path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=64, seed=0, target_size=target, minimize="time",
                                       candidates=spec.get("tree_candidates", ()), dtype=spec["dtype"])
    search_s = time.perf_counter() - t0
    n_sl = int(np.prod([net.sizes[i] for i in sliced])) if sliced else 1
    per_madd = 8 if np.dtype(spec["dtype"]).kind == "c" else 2
    flops = per_madd * info["total_madds"]
    timed_slices = min(n_sl, 16)           # sliced trees: a bounded sample of slices per repetition, scaled to all of them

    def many(k):
        # as the slice loop does: the host prepares and uploads the next contraction while the GPU runs this one; every result is
        # read back (device -> host) before the timed region ends
        if sliced:
            return [np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=list(range(timed_slices)))[0]) for _ in range(k)]
        outs = [T.contract_network(net, path, graph=True)[0] for _ in range(k)]
        return [np.asarray(o) for o in outs]

    if sliced:
        reps = max(2, reps // timed_slices)
    amp = many(2)[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record()
    amp = many(reps)[-1]
    e1.record()
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3) / reps * (n_sl / timed_slices)
    fams = profile_tn_step(lambda: np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=[0], graph=False)[0]) if sliced
                           else np.asarray(T.contract_network(net, path, graph=False)[0]))
    (fams.get("permute_split") or {}).pop("largest", None)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    ap = fams.get("apply")
    got = np.asarray(amp).ravel()
    out = {"what": "time to the whole result of the same network along the tree found by tn.search_path (64 fresh seeded greedy trees + the recorded ones of the "
                   "workload spec, objective: modelled time), through tn.contract_network from host tensors to the host result, CUDA graph replayed"
                   + ("; %d of the tree's %d slices timed per repetition, scaled to all of them" % (timed_slices, n_sl) if sliced else ""),
           "search": {"trials": info["trials"], "seconds": search_s, "alpha": info["alpha"], "temperature": info["temperature"], "trial_seed": info["trial_seed"]},
           "log2_flops": round(log2(max(flops, 1)), 2), "log2_largest_intermediate": round(log2(info["largest"]), 2), "n_slices": n_sl,
           "ms_per_result": ms, "tflops": flops / (ms * 1e-3) / 1e12,
           "headline_tree": {"steps": headline_steps, "ms_per_result": headline_ms_per_step * headline_steps},
           "speedup_time_to_result": headline_ms_per_step * headline_steps / ms,
           "per_family_ms": {k: round(v["ms"], 3) for k, v in fams.items()}, "per_family_calls": {k: v["calls"] for k, v in fams.items()},
           "roofline_apply": None if not (ap and ap["ms"]) else {
               "bound": "hbm", "achieved": ap["bytes"] / (ap["ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": ap["bytes"] / (ap["ms"] * 1e-3) / 1e9 / hbm,
               "launches": ap["calls"], "algorithmic_bytes": ap["bytes"],
               "kernel": "apply_nd_kernel (rnb_contract_nd, one operand tiny): the launches that move >= 64 MB; bytes = both operands read once + the result written once; "
                         "eager pass with every launch bracketed by CUDA events"}}
    if not sliced:
        out["result"] = [float(np.real(got[0])), float(np.imag(got[0]))]
        if headline_steps <= 64:
            # the headline tree's result: all of its steps (slices), on the GPU
            if headline_sliced and spec["mode"] == "loop":
                head = np.asarray(T.contract_network(net, headline_path, sliced=headline_sliced, slice_mode="loop")[0]).ravel()
            else:
                head = np.asarray(T.contract_network(net, headline_path, sliced=headline_sliced, slice_mode="blocks")[0]).ravel()
            out["result_headline_tree"] = [float(np.real(head[0])), float(np.imag(head[0]))]
            out["rel_diff_vs_headline_tree"] = float(np.linalg.norm(got - head) / (np.linalg.norm(head) or 1.0))
        if with_host:
            wide = np.complex128 if np.dtype(spec["dtype"]).kind == "c" else np.float64
            tensors = {i: (np.asarray(a).astype(wide), tuple(ix)) for i, (a, ix) in enumerate(zip(net.arrays, net.inputs))}
            nxt = len(tensors)
            t0 = time.perf_counter()
            with all_host_threads():
                for a, b_ in path:
                    (A, ia), (B, ib) = tensors.pop(a), tensors.pop(b_)
                    sh = [i for i in ia if i in ib]
                    C = np.tensordot(A, B, ([ia.index(i) for i in sh], [ib.index(i) for i in sh]))
                    tensors[nxt] = (C, tuple(i for i in ia if i not in sh) + tuple(i for i in ib if i not in sh))
                    nxt += 1
            ref = np.asarray(list(tensors.values())[0][0]).ravel()
            out["host_double"] = {"result": [float(np.real(ref[0])), float(np.imag(ref[0]))], "seconds": time.perf_counter() - t0,
                                  "rel_diff": float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) or 1.0)),
                                  "what": "the same tree with numpy.tensordot per pairwise step in double precision on the host cores"}
    return out


def side_configs(names, steps=8):
    """Short runs of other workloads in their own processes; returns {name: digest of the line}."""
    out = {}
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "LOCAL_WORLD_SIZE", "GROUP_RANK")}
    for name in names:
        try:
            proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", name, "--steps", str(steps), "--warmup", "3", "--no-extra",
                                   "--no-cpu-baseline", "--no-opt-in"] + (["--searched-tree"] if name in TN_WORKLOADS else []),
                                  env=env, capture_output=True, text=True, timeout=300)
            line = json.loads(proc.stdout.strip().splitlines()[-1])
            r = line.get("roofline") or {}
            out[name] = {"value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "dtype": line["dtype"], "steps": line["steps"],
                         "e2e_value": (line.get("e2e") or {}).get("value"), "roofline_frac": r.get("frac"), "roofline_peak": r.get("peak"),
                         "roofline_permute_frac": (line.get("roofline_permute") or {}).get("frac"), "parity_rel_frobenius": line.get("parity_rel_frobenius"),
                         "result": line.get("result"), "gpu_launches": line.get("gpu_launches"), "config": line.get("config")}
            st = (line.get("extra") or {}).get("searched_tree")
            if st:
                out[name]["searched_tree"] = {k: st.get(k) for k in ("log2_flops", "log2_largest_intermediate", "n_slices", "ms_per_result", "tflops", "headline_tree",
                                                                     "speedup_time_to_result", "rel_diff_vs_headline_tree", "search", "error") if k in st}
        except Exception as e:
            out[name] = {"error": repr(e)[:300]}
    return out


def run_tn(args):
    """TN workloads: a step is one slice (loop mode) or one whole contraction (blocks mode)."""
    spec = TN_WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from math import log2

    net, path, sliced = build_tn(spec)
    from rosnet_b200 import tn as T

    per_madd = 8 if np.dtype(spec["dtype"]).kind == "c" else 2
    single = spec["dtype"] in ("complex64", "float32")
    dtype_label = {"complex64": "c64 (3xTF32)", "complex128": "c128", "float32": "f32 (3xTF32)", "float64": "f64"}[spec["dtype"]]
    if args.precision == "tf32_bf16x2":
        dtype_label = dtype_label.replace("3xTF32", "TF32 + 2xBF16 cross terms")
    loop = spec["mode"] == "loop" and bool(sliced)
    madds, largest, steps = T.path_cost(net.inputs, net.output, net.sizes, path, sliced if loop else ())
    flops_step = per_madd * madds
    n_slices = int(np.prod([net.sizes[i] for i in sliced])) if loop else 1
    cfg = {"workload": args.workload, "tensors": net.n_tensors, "path_steps": len(path), "log2_flops_per_step": round(log2(max(flops_step, 1)), 2),
           "log2_largest_intermediate": round(log2(largest), 2), "sliced_indices": len(sliced), "total_slices": n_slices,
           "slice_mode": spec["mode"], "path": "greedy(seed=0, alpha=%s)" % spec["alpha"], "step": "one slice" if loop else "whole contraction",
           "cuda_graph": "one pass is captured once and replayed for every further step (inputs re-uploaded into a static arena)",
           "partition": "independent slices dealt round-robin over the ranks by the scheduler inside contract_network(comm=...); one all-reduce (NCCL) of the summed amplitude per call, inside the timed region",
           "l2": "every slice streams intermediates of up to %d MB through HBM (larger than the 126 MB L2)" % ((largest * np.dtype(spec["dtype"]).itemsize) >> 20)}
    if args.workload == "bench_random_tn":
        cfg["bond"] = spec["d"]
        cfg["rescoped"] = ("BASELINE.json names bond dim 16: 200 tensors with mean degree 3 and bond 16 have a contraction width of 2^156 and single input "
                           "tensors of 16^8 elements (tests/test_tn_host.py), so the line runs the reference script's own default bond 2")
    if args.impl == "reference":
        if rank != 0:
            return
        with all_host_threads():
            done, secs = 0, 0.0
            for _ in range(max(args.warmup, 0)):
                cpu_tn_sample(net, path, sliced if loop else (), 5e10, per_madd)
            for _ in range(max(args.steps, 1)):
                d, s = cpu_tn_sample(net, path, sliced if loop else (), 1.5e12, per_madd)
                done, secs = done + d, secs + s
            # ONE whole step (a whole slice / the whole contraction), once: the sample above only reaches the leading pairwise
            # steps, the step's large GEMMs (where the BLAS is at its best) come last
            full_d, full_s = cpu_tn_sample(net, path, sliced if loop else (), float("inf"), per_madd)
            cores = blas_threads()
        sample_val = done / secs / 1e12
        val = full_d / full_s / 1e12
        sample = ("value = ONE whole %s on the host (%.3g flop in %.1f s, numpy.tensordot per pairwise step = the reference's arithmetic), measured once; "
                  "bounded sample per step: the leading pairwise steps up to 1.5e12 flop, %d steps, %.4f TFLOP/s; numpy %s, BLAS threads=%d, os.cpu_count=%s" % (
                      "slice" if loop else "contraction", full_d, full_s, max(args.steps, 1), sample_val, np.__version__, cores, os.cpu_count()))
        print(json.dumps({"impl": "reference", "metric": "contraction TFLOP/s", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_s * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": dtype_label.split(" ")[0], "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample,
                                           "full_step_s": full_s, "full_step_flops": full_d, "sample_value": sample_val},
                          "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch

    import rosnet_b200 as rn
    from rosnet_b200.engine import lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load()
    from rosnet_b200.engine import scheduler

    ctx = scheduler.init() if world > 1 else None     # the single-box scheduler: slices are dealt inside contract_network

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)

    def all_ids(k):
        """k steps per rank: the job's slice list (the scheduler deals it round-robin, rank r gets ids r, r + world, ...)"""
        return [i % max(n_slices, 1) for i in range(world * k)]

    def step_ids(k):
        return all_ids(k)[rank::world]

    def run_e2e(k, graph=True, alone=False):
        """k steps per rank through the public API, host tensors in, host result out (H2D + D2H inside; at N > 1 the
        scheduler deals the slices and all-reduces the summed amplitude, so every rank returns the job's total).
        graph: replay the pass as a CUDA graph (the slice loop does so by default from the third slice on;
        single passes opt in because the same network is contracted repeatedly here)."""
        if loop and alone:      # this rank's share only, no collective (the profiling step rank 0 runs after the others left)
            with rn.tuning.configure(n_gpus=1):
                return T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=step_ids(k), graph=None if graph else False)[0]
        if loop:
            return T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=all_ids(k), graph=None if graph else False, comm=ctx)[0]
        outs = []
        with rn.tuning.configure(n_gpus=1):     # whole-contraction steps: every rank contracts its own replica (weak scaling)
            for _ in range(k):
                # asynchronous: the result is a device BlockArray with work pending on the stream, so the host prepares and
                # uploads the next contraction's inputs while the GPU still runs this one (as the slice loop does)
                outs.append(T.contract_network(net, path, sliced=sliced, slice_mode="blocks", graph=graph)[0])
        acc_ = 0
        for o in outs:                          # device -> host reads of every step's result, still inside the timed region
            acc_ = acc_ + np.asarray(o)
        return acc_

    peaks_t = None
    if not args.no_peak_probe and rank == 0:
        peaks_t = measure_tensor_peak(single)

    run_e2e(warm)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        w0 = time.perf_counter()
        out = fn()
        e1.record()
        barrier()
        ms_ = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
        t = torch.tensor([ms_], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0])

    # ---- value: input tensors already resident in HBM, results left on the device ----
    before = dict(lib.launch_counter)
    if loop:
        resident = T.stage_slices(net, sliced, step_ids(args.steps))
        tdt = {"complex64": torch.complex64, "complex128": torch.complex128, "float32": torch.float32, "float64": torch.float64}[spec["dtype"]]

        def resident_steps():
            kept_ = T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=step_ids(args.steps), resident=resident, fetch=False)[0]
            if ctx is None:
                return kept_
            # the job's amplitude: device-side sum of this rank's slice results (one block-sum launch), then ONE all-reduce
            from rosnet_b200.array.device import DeviceStorage, current_stream_ptr

            nbytes = kept_[0]._storage.nbytes
            acc_dev = DeviceStorage(nbytes)
            lib.block_sum(acc_dev.ptr, [x._storage.ptr for x in kept_], nbytes // np.dtype(spec["dtype"]).itemsize, lib.DTYPE_CODES[spec["dtype"]], False,
                          current_stream_ptr())
            ctx.comm.all_reduce_sum(acc_dev.tensor[:nbytes].view(tdt))
            return acc_dev

        kept, ms = timed(resident_steps)
        del kept, resident
    else:
        _, ms = timed(lambda: run_e2e(args.steps))
    launches = sum(lib.launch_counter.values()) - sum(before.values())
    # ---- e2e: the public call from host tensors to the host result ----
    acc, e2e_ms = timed(lambda: run_e2e(args.steps))
    clocks = sampler.stop() if rank == 0 else None
    value = world * flops_step * args.steps / (ms * 1e-3) / 1e12
    e2e_value = world * flops_step * args.steps / (e2e_ms * 1e-3) / 1e12
    # ---- the opt-in arithmetic next to the default one, same run, same path (single precision only) ----
    opt_in = None
    if single and args.precision == "default" and not args.no_opt_in:
        with rn.tuning.configure(precision="tf32_bf16x2"):
            run_e2e(warm)
            acc2, ms2 = timed(lambda: run_e2e(args.steps))
        a1, a2 = np.asarray(acc).ravel(), np.asarray(acc2).ravel()
        opt_in = {"precision": "tf32_bf16x2: hi.hi product in TF32, the two cross products as BF16 MMAs (2 instead of 3 TF32-equivalents of tensor work); NOT the default",
                  "e2e_value": world * flops_step * args.steps / (ms2 * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms2 / args.steps,
                  "rel_diff_of_result_vs_default": float(np.linalg.norm(a2 - a1) / (np.linalg.norm(a1) or 1.0))}
    # ---- extras the driver's scaling run carries at every N: the collective tiers and cross-rank parity ----
    extra = {}
    if loop and ctx is not None and not args.no_extra:
        # the all-reduced amplitude of `world` slices (one per rank) against rank 0 contracting the same slices alone
        ids = list(range(min(world, n_slices)))
        tot_n = np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=ids, comm=ctx, graph=False)[0]).ravel()
        same = torch.from_numpy(np.ascontiguousarray(tot_n).view(np.float64).copy()).cuda()
        lo, hi = same.clone(), same.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        extra["amplitude_vs_n1"] = {"slices": len(ids), "identical_on_every_rank": bool(torch.equal(lo, hi))}
        if rank == 0:
            with rn.tuning.configure(n_gpus=1):
                tot_1 = np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=ids, graph=False)[0]).ravel()
            extra["amplitude_vs_n1"].update({"rel_diff": float(np.linalg.norm(tot_n - tot_1) / (np.linalg.norm(tot_1) or 1.0)),
                                             "amplitude_n_ranks": [float(np.real(tot_n[0])), float(np.imag(tot_n[0]))],
                                             "amplitude_one_rank": [float(np.real(tot_1[0])), float(np.imag(tot_1[0]))]})
        barrier()
    if not args.no_extra:
        # BASELINE configs[3] (contracted axis sharded over the ranks) through the scheduler: NCCL and fused exchange
        try:
            extra["schmidt"] = schmidt_case(WORKLOADS["bench_schmidt_matrix"], ctx, steps=3, warmup=1)
        except Exception as e:   # the headline number must survive a failure of the side measurement
            extra["schmidt"] = {"error": repr(e)[:300]}
        barrier()
    if ctx is not None and not args.no_extra:
        # tier 1 behind the API: BASELINE configs[0] with its output blocks dealt over the ranks (strong scaling, no exchange)
        try:
            extra["output_blocks"] = output_blocks_case(WORKLOADS["bench_contraction"], ctx, steps=8, warmup=2)
        except Exception as e:
            extra["output_blocks"] = {"error": repr(e)[:300]}
        barrier()
    if rank == 0 and world == 1 and (not args.no_extra or args.searched_tree):
        # time to the whole result along a searched tree next to the headline tree (same network)
        try:
            extra["searched_tree"] = searched_tree_case(net, spec, ms / args.steps, n_slices if loop else 1, path, sliced, with_host=not args.no_cpu_baseline)
        except Exception as e:
            extra["searched_tree"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_extra and not args.no_side_configs:
        # the other BASELINE configs on this GPU, each as its own short bench.py run (configs[0], [1], [2]); their full lines
        # are what `--workload NAME` prints
        extra["configs"] = side_configs(("bench_contraction", "bench_random_tn", "bench_qft"))
    if rank != 0:
        if dist is not None:
            scheduler.shutdown()
            dist.destroy_process_group()
        return
    # the per-family split of a step in STEADY STATE: the side measurements above let the GPU cool down, and a slice right after an
    # idle second runs its GEMMs ~8 % faster than inside the timed loop (sw_power_cap) - so replay a few steps first
    run_e2e(4, alone=True)
    fams = profile_tn_step(lambda: run_e2e(1, graph=False, alone=True))
    largest_split = (fams.get("permute_split") or {}).pop("largest", None)
    gpu_ms = sum(f["ms"] for f in fams.values())
    gemm = fams.get("gemm", {"ms": 0.0, "flops": 0, "calls": 0, "tensor_flops": 0.0})
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else None
    # executed tensor-core flops per algorithmic flop of this step (3xTF32: 3; complex64 3M: 2.25; FP64: 1), flop-weighted
    mult = (gemm["tensor_flops"] / gemm["flops"]) if gemm.get("flops") else (3.0 if single else 1.0)
    peak_tflops = (peaks_t["sustained"] / mult) if peaks_t else None
    peak_note = "not measured" if not peaks_t else (
        "raw tensor peak %.1f TFLOP/s sustained (%.1f burst); %s; divided by %.3f = tensor-core flops this step EXECUTES per algorithmic flop "
        "(3xTF32 runs 3 TF32 products per product, complex64 3M runs 9 instead of 12 per complex multiply-add), so frac = executed tensor flop/s / raw sustained peak"
        % (peaks_t["sustained"], peaks_t["burst"], peaks_t["how"], mult))
    perm = fams.get("permute_split") or fams.get("permute")
    perm_fused = "permute_split" in fams
    kname = ("gemm_tf32_v2_kernel (3xTF32 tcgen05 grouped GEMM on CTA pairs; complex64 as three real GEMMs (3M) when k >= 1024; incl. the split pre-pass, "
             "the split-K second pass and the 3M combine pass)") if single else "gemm_f64_kernel (FP64 DMMA grouped GEMM)"
    traffic = {}
    try:
        import glob

        for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic.json"))):
            traffic.update(json.load(open(f)))
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": (achieved / peak_tflops) if (achieved and peak_tflops) else None,
                "traffic": traffic.get(("gemm_tf32_kernel:" if single else "gemm_f64_kernel:") + args.workload),
                "traffic_note": "dram read+write bytes of the kernel's launches in ONE step, summed from an ncu --set full capture (profiles/)", "kernel": kname,
                "frac_of_burst": (achieved * mult / peaks_t["burst"]) if (achieved and peaks_t) else None, "peak_used": "sustained",
                "raw_tensor_peak": peaks_t, "executed_tensor_flops_per_algorithmic_flop": mult,
                "algorithmic_flops_per_step_in_kernel": gemm["flops"], "launches_per_step": gemm["calls"], "kernel_ms_per_step": gemm["ms"],
                "share_of_gpu_time": gemm["ms"] / gpu_ms if gpu_ms else None, "peak_source": peak_note,
                "how": "one extra untimed step, run right after four replayed steps (GPU at its steady-state clock), with every launch bracketed by CUDA events; achieved = flops routed to the GEMM / its summed time",
                "per_family_ms": {k: round(v["ms"], 3) for k, v in fams.items()}, "per_family_calls": {k: v["calls"] for k, v in fams.items()},
                "small_steps_batched": fams.get("small_batch", {}).get("tasks")}
    roofline_permute = None
    if perm and perm["ms"]:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        roofline_permute = {"bound": "hbm", "achieved": perm["bytes"] / (perm["ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": perm["bytes"] / (perm["ms"] * 1e-3) / 1e9 / hbm, "traffic": None,
                            "kernel": ("permute_tiled_kernel with the fused split epilogue (rnb_permute_split: matricize + hi/lo TF32 planes in one pass), the step's %d launches that move >= 64 MB "
                                       "(rank-20+ tensors of extent 2); bytes = operand read once + planes written (2x / 3x / 4x the operand)" % perm["calls"]) if perm_fused else
                                      ("permute kernels, the step's %d matricize launches that move >= 64 MB (rank-20+ tensors of extent 2; the other %d launches are latency-bound)" % (
                                          perm["calls"], fams.get("permute_small", {"calls": 0})["calls"])),
                            "algorithmic_bytes_per_step": perm["bytes"],
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
        if largest_split:
            try:
                iso = isolated_permute_split(largest_split)
                roofline_permute["isolated"] = {
                    "achieved": iso["GBps"], "frac": iso["GBps"] / hbm, "ms_per_launch": iso["ms"], "bytes_per_launch": iso["bytes"], "launches": iso["launches"],
                    "note": "the step's largest fused matricize + split launch alone (same descriptor, fresh buffers, GPU idle for 1 s before, launches back to "
                            "back): the kernel's own rate; `achieved` above is the same kernel inside the step, where it starts at the SM clock the "
                            "power-capped GEMM before it left behind"}
            except Exception as e:
                roofline_permute["isolated"] = {"error": repr(e)[:200]}
    in_bytes = int(sum(np.asarray(a).nbytes for a in net.arrays))
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        with all_host_threads():
            cpu_tn_sample(net, path, sliced if loop else (), 5e10, per_madd)
            # one WHOLE step when that is ~10-30 s of CPU work (a Sycamore slice: ~13 s on 16 cores), else its leading pairwise steps
            whole = flops_step <= 3e13
            d, s = cpu_tn_sample(net, path, sliced if loop else (), float("inf") if whole else 6e12, per_madd)
            cores = blas_threads()
        cpu = {"value": d / s / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port", "seconds": s,
               "sample": "numpy.tensordot per pairwise step (the reference's arithmetic), %s: %.3g of %.3g flop in %.1f s; numpy %s, BLAS threads=%d, os.cpu_count=%s" % (
                   "ONE WHOLE step" if whole else "the leading pairwise steps of one step", d, flops_step, s, np.__version__, cores, os.cpu_count())}
    res0 = np.asarray(acc).ravel()
    line = {"metric": "contraction TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "s_per_slice": ms / args.steps / 1e3 if loop else None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype_label, "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(max(res0.size, 1) * np.dtype(spec["dtype"]).itemsize),
                    "ms_per_step": e2e_ms / args.steps,
                    "note": "rosnet_b200.tn.contract_network from host ndarrays to the host result: per slice one pinned H2D of all input tensors and one D2H of the result inside the timed region"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_permute": roofline_permute,
            "cpu_baseline": cpu, "result": [float(np.real(res0[0])), float(np.imag(res0[0]))] if res0.size else None,
            "opt_in_precision": opt_in, "extra": extra}
    print(json.dumps(line))
    if dist is not None:
        scheduler.shutdown()
        dist.destroy_process_group()


def output_blocks_case(wl, ctx, steps, warmup):
    """Tier 1 behind the API at N > 1: BASELINE configs[0] with REPLICATED device operands; under the active
    scheduler context np.tensordot deals the 16 output blocks over the ranks (`scheduler.choose_tier` -> "output", no
    exchange) and returns a ShardedBlockArray.  Strong scaling: the total work is fixed.  Parity: every owned block of every
    rank against torch.tensordot in float64 on the regenerated dense operands (cuBLAS DGEMM, outside the product path)."""
    import torch
    import torch.distributed as dist

    import rosnet_b200 as rn
    from rosnet_b200.array.device import DeviceStorage
    from rosnet_b200.engine import runtime

    sa, ab, sb, bb, dtype, axes = wl
    tdt = {"complex128": torch.complex128, "float64": torch.float64}[dtype]
    itemsize = np.dtype(dtype).itemsize

    def gen(shape, seed):
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        return torch.randn(shape, dtype=tdt, device="cuda", generator=g)

    def blocked(dense, blockshape):
        grid = tuple(s_ // b_ for s_, b_ in zip(dense.shape, blockshape))
        st = DeviceStorage(dense.numel() * itemsize)
        runtime.dense_to_blocks(dense.data_ptr(), st, grid, blockshape, itemsize)
        torch.cuda.synchronize()
        return rn.BlockArray._from_storage(st, grid, blockshape, dtype)

    U, V = gen(sa, 77), gen(sb, 78)
    A, B = blocked(U, ab), blocked(V, bb)
    for _ in range(warmup):
        C = np.tensordot(A, B, axes)
    ctx.barrier()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        C = np.tensordot(A, B, axes)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in evs]))], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    flops_total = algorithmic_flops(sa, sb, dtype, axes)
    out = {"workload": "bench_contraction", "dtype": dtype, "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "scaling": "strong",
           "result_type": type(C).__name__, "ms_per_step": ms, "tflops": flops_total / ms / 1e9,
           "partition": "output blocks dealt over the ranks by np.tensordot itself (scheduler tier 'output'), replicated operands, no exchange"}
    want = torch.tensordot(U, V, dims=([int(x) for x in axes[0]], [int(x) for x in axes[1]]))
    owned = C.owned() if hasattr(C, "owned") else C
    indices = C.owned_indices() if hasattr(C, "owned_indices") else list(np.ndindex(*C.grid))
    obs = C.blockshape
    worst = 0.0
    for local, gidx in enumerate(indices):
        sl = tuple(slice(g * b_, (g + 1) * b_) for g, b_ in zip(gidx, obs))
        nb = int(np.prod(obs)) * itemsize
        got = owned._storage.tensor[local * nb:(local + 1) * nb].view(tdt).reshape(obs)
        err = float(torch.linalg.norm(got - want[sl]) / torch.linalg.norm(want[sl]))
        if err != err:
            worst = float("nan")
            break
        worst = max(worst, err)
    w = torch.tensor([worst if worst == worst else 1e300], dtype=torch.float64, device="cuda")
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    n_chk = torch.tensor([len(indices)], dtype=torch.int64, device="cuda")
    dist.all_reduce(n_chk)
    out["parity_rel_frobenius_all_blocks"] = float(w[0])
    out["blocks_checked"] = int(n_chk[0])
    del U, V, want
    torch.cuda.empty_cache()
    return out


def schmidt_case(wl, ctx, steps, warmup, exchanges=("nccl", "fused"), check=True):
    """BASELINE configs[3]: C = tensordot(U, V, [(1,), (1,)]) with the CONTRACTED block axis sharded over the ranks,
    through the public API — every rank wraps its k-slab as a ShardedBlockArray and calls np.tensordot; the scheduler
    (rosnet_b200/array/sharded.py) contracts the slab into a partial of every output block and sums the partials by an
    NCCL reduce-scatter or inside the GEMM epilogue over NVLink peer memory.  N = 1: the plain one-GPU np.tensordot.
    Synthetic operands are generated on the device (seeded, identical on every rank; rank r keeps slab r).
    Parity: EVERY owned output block of every rank against torch.matmul in complex128 on the regenerated whole operands
    (cuBLAS ZGEMM, a checker outside the product path), plus output block 0 against numpy.tensordot on the host."""
    import torch

    import rosnet_b200 as rn
    from rosnet_b200.array.device import DeviceStorage
    from rosnet_b200.engine import lib, runtime

    sa, ab, sb, bb, dtype, axes = wl
    rank, world = (ctx.rank, ctx.world) if ctx is not None else (0, 1)
    tdt = {"complex128": torch.complex128, "float64": torch.float64, "complex64": torch.complex64, "float32": torch.float32}[dtype]
    itemsize = np.dtype(dtype).itemsize
    kax_a, kax_b = axes[0][0], axes[1][0]
    assert len(sa) == 2 and len(sb) == 2 and kax_a == 1 and kax_b == 1, "schmidt_case expects C = U . V^T"
    grid_k = sa[1] // ab[1]
    assert grid_k % world == 0, "contracted block axis must divide over the ranks"
    kloc = sa[1] // world

    def gen(shape, seed):
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        return torch.randn(shape, dtype=tdt, device="cuda", generator=g)

    U, V = gen(sa, 1234), gen(sb, 4321)            # whole operands (the checker's inputs), identical on every rank

    def blocked(dense, blockshape):
        dense = dense.contiguous()
        grid = tuple(s // b for s, b in zip(dense.shape, blockshape))
        st = DeviceStorage(dense.numel() * itemsize)
        runtime.dense_to_blocks(dense.data_ptr(), st, grid, blockshape, itemsize)
        torch.cuda.synchronize()
        return rn.BlockArray._from_storage(st, grid, blockshape, dtype)

    A_loc = blocked(U[:, rank * kloc:(rank + 1) * kloc], ab)
    B_loc = blocked(V[:, rank * kloc:(rank + 1) * kloc], bb)
    flops_total = algorithmic_flops(sa, sb, dtype, axes)
    out = {"workload": "bench_schmidt_matrix" if sa[0] == 2**14 else "bench_schmidt_small", "shape": list(sa), "blockshape": list(ab), "dtype": dtype,
           "n_gpus": world, "steps": steps, "warmup": warmup, "total_flops": flops_total,
           "partition": "one GPU, no partition" if world == 1 else "contracted block axis split %d blocks/rank; rank r ends up owning output blocks [r*chunk, (r+1)*chunk)" % (grid_k // world)}

    def timed(fn):
        """median over `steps` individually timed steps (CUDA events on the launching stream), max over ranks"""
        for _ in range(warmup):
            res = fn()
        if ctx is not None:
            ctx.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        t = torch.tensor([float(np.median([a.elapsed_time(b) for a, b in evs]))], dtype=torch.float64, device="cuda")
        if ctx is not None:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t[0])

    def check_blocks(owned_ba, indices):
        """max over the owned blocks of the relative Frobenius error against the ZGEMM checker"""
        worst = 0.0
        bm, bn = ab[0], bb[0]
        for local, (gi, gj) in enumerate(indices):
            want = U[gi * bm:(gi + 1) * bm] @ V[gj * bn:(gj + 1) * bn].T
            blk = owned_ba.data.flat[local]
            got = owned_ba._storage.tensor[local * blk.nbytes:(local + 1) * blk.nbytes].view(tdt).reshape(bm, bn)
            err = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
            if err != err:
                return float("nan")            # max() would hide a NaN block behind the other blocks
            worst = max(worst, err)
        return worst

    if world == 1:
        with rn.tuning.configure(n_gpus=1):
            C, ms = timed(lambda: np.tensordot(A_loc, B_loc, axes))
        out["one_gpu_ms"] = ms
        out["one_gpu_tflops"] = flops_total / ms / 1e9
        if check:
            out["parity_rel_frobenius_all_blocks"] = check_blocks(C, list(np.ndindex(*C.grid)))
            out["blocks_checked"] = C.nblock
            if out["parity_rel_frobenius_all_blocks"] == 0.0:
                # measured (scripts/_build-style probe, 2048-blocks, k = 16384): torch.equal(got, want) is True for every block —
                # cuBLAS ZGEMM and gemm_f64_kernel add the same DMMA products in the same order at this shape; not a skipped check
                out["parity_note"] = "0.0 = bit-identical to the cuBLAS ZGEMM checker on all blocks (no NaN; block 0 also checked against numpy)"
            blk0 = np.asarray(C.data.flat[0])
    else:
        import torch.distributed as dist

        A = rn.ShardedBlockArray.from_local_slab(A_loc, 1, ctx)
        B = rn.ShardedBlockArray.from_local_slab(B_loc, 1, ctx)
        # the local share of the work alone (no exchange): what the exchange adds is measured against it (timed before AND
        # after the exchange variants, the smaller of the two medians: the first launches of a process run slower)
        def local_gemm():
            with rn.tuning.configure(n_gpus=1):
                return timed(lambda: np.tensordot(A_loc, B_loc, axes))[1]

        gemm_ms = local_gemm()
        sent = (world - 1) / world * (sa[0] * sb[0] * itemsize)       # bytes of partials each rank hands to the other owners
        blk0 = None
        for ex in exchanges:
            old = ctx.exchange
            ctx.exchange = ex
            try:
                C, ms = timed(lambda: np.tensordot(A, B, axes))
            finally:
                ctx.exchange = old
            out[ex + "_ms"] = ms
            out[ex + "_tflops"] = flops_total / ms / 1e9
            if check:
                worst = torch.tensor([check_blocks(C.owned(), C.owned_indices())], dtype=torch.float64, device="cuda")
                dist.all_reduce(worst, op=dist.ReduceOp.MAX)
                n_chk = torch.tensor([C.count], dtype=torch.int64, device="cuda")
                dist.all_reduce(n_chk)
                out[ex + "_parity_rel_frobenius_all_blocks"] = float(worst[0])
                out["blocks_checked"] = int(n_chk[0])
            if rank == 0 and blk0 is None:
                blk0 = np.asarray(C.owned().data.flat[0])
        gemm_ms = min(gemm_ms, local_gemm())
        out["local_gemm_ms"] = gemm_ms
        for ex in exchanges:
            out[ex + "_exchange_ms"] = out[ex + "_ms"] - gemm_ms     # what the exchange ADDS to the local GEMM (fused: ~0, it overlaps the tiles)
        out["bytes_sent_per_gpu"] = int(sent)
        if "nccl" in exchanges:
            # the reduce-scatter runs after the GEMM: its time is exposed, so bytes / time is the link rate it achieved
            out["exchange_ms"] = out["nccl_exchange_ms"]
            out["nvlink_GBps"] = sent / max(out["nccl_exchange_ms"], 1e-3) / 1e6
        if "fused" in exchanges:
            # the epilogue pushes tiles while other tiles compute: the same bytes cross NVLink inside the GEMM's own time
            out["fused_nvlink_GBps_lower_bound"] = sent / out["fused_ms"] / 1e6
    if check and rank == 0:
        # the tie to the reference's arithmetic: output block 0 with numpy.tensordot on the host
        u0, v0 = U[:ab[0]].cpu().numpy(), V[:bb[0]].cpu().numpy()
        with all_host_threads():
            want0 = np.tensordot(u0, v0, axes)
        out["parity_block0_vs_numpy"] = float(np.linalg.norm(blk0 - want0) / np.linalg.norm(want0))
    del U, V
    torch.cuda.empty_cache()
    return out


def run_ksplit(args, wl):
    """--workload bench_schmidt_matrix|bench_schmidt_small: strong scaling of the k-split tier (total work fixed)."""
    import torch

    from rosnet_b200.engine import lib, scheduler

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ctx = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx = scheduler.init()
    lib.load()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    before = dict(lib.launch_counter)
    exchanges = ("nccl", "fused") if args.exchange == "both" else (args.exchange,)
    res = schmidt_case(wl, ctx, args.steps, max(args.warmup, 3), exchanges=exchanges)
    launches = sum(lib.launch_counter.values()) - sum(before.values())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        ms = res["one_gpu_ms"] if world == 1 else min(res[e + "_ms"] for e in exchanges)
        sa, ab, sb, bb, dtype, axes = wl
        line = {"metric": "contraction TFLOP/s", "value": res["total_flops"] / ms / 1e9, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": {"complex128": "c128"}.get(dtype, dtype), "data": "synthetic (torch.randn on the device, seeded)",
                "config": {"workload": args.workload, "shape_a": list(sa), "blockshape_a": list(ab), "axes": [list(x) for x in axes],
                           "partition": res["partition"], "exchange": list(exchanges), "total_flops": res["total_flops"],
                           "l2": "operands of 4 GiB each stream through HBM (larger than the 126 MB L2)"},
                "e2e": None, "gpu_launches": int(launches), "clocks": clocks, "roofline": None, "cpu_baseline": None, "schmidt": res}
        print(json.dumps(line))
    if ctx is not None:
        import torch.distributed as dist

        scheduler.shutdown()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="bench_sycamore", choices=sorted(WORKLOADS) + sorted(TN_WORKLOADS))
    ap.add_argument("--complex-mode", default="auto", choices=["auto", "4m", "3m"],
                    help="complex dtypes: 4 or 3 real products per complex product (auto: complex128 3M from k = 64; complex64 3M when k >= 1024 and m, n >= 128)")
    ap.add_argument("--exchange", default="both", choices=["nccl", "fused", "both"],
                    help="k-split workloads at N>1: NCCL reduce-scatter after the GEMM, or the GEMM epilogue reducing into the owner's buffer over peer memory")
    ap.add_argument("--precision", default="default", choices=["default", "tf32_bf16x2"],
                    help="float32/complex64: 3xTF32 (default) or the opt-in TF32 hi.hi + 2 BF16 cross products mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-opt-in", action="store_true", help="skip the extra timed region with the opt-in tf32_bf16x2 arithmetic")
    ap.add_argument("--no-peak-probe", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="N = 1: skip the short runs of the other BASELINE configs carried in extra.configs")
    ap.add_argument("--searched-tree", action="store_true", help="TN workloads: carry extra.searched_tree even with --no-extra (the side-config runs do)")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements carried in `extra` (k-split tier, cross-rank amplitude check)")
    args = ap.parse_args()
    if args.precision != "default" and args.impl == "native":
        from rosnet_b200.engine import lib as _l, runtime as _r

        _r.settings["precision"] = {"tf32_bf16x2": _l.RNB_PREC_TF32_BF16X2}[args.precision]
    if args.workload in TN_WORKLOADS:
        return run_tn(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.workload.startswith("bench_schmidt"):
        return run_ksplit(args, wl)

    import torch

    import rosnet_b200 as rn
    from rosnet_b200.array.device import pinned_empty
    from rosnet_b200.engine import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load()
    from rosnet_b200.engine import runtime as _rt

    _rt.settings["complex_mode"] = {"3m": lib.RNB_CPLX_3M, "4m": lib.RNB_CPLX_4M, "auto": lib.RNB_CPLX_AUTO}[args.complex_mode]

    sa, ab, sb, bb, dtype, axes = wl
    itemsize = np.dtype(dtype).itemsize
    flops = algorithmic_flops(sa, sb, dtype, axes)  # per rank (weak scaling: each rank one slab)

    # ---- synthetic operands: rank r's slab of `a` uses seed 2r, `b` (replicated) uses seed 1 ----
    a = make_dense(sa, dtype, 2 * rank)
    b = make_dense(sb, dtype, 1)

    def pinned_blocks(dense, blockshape):
        """BlockArray whose host blocks are views of ONE pinned buffer in block-major order."""
        grid = tuple(s // t for s, t in zip(dense.shape, blockshape))
        buf = pinned_empty((int(np.prod(grid)),) + tuple(blockshape), dense.dtype)
        cells = np.empty(grid, dtype=object)
        for i, idx in enumerate(np.ndindex(*grid)):
            sl = tuple(slice(j * t, (j + 1) * t) for j, t in zip(idx, blockshape))
            buf[i] = dense[sl]
            cells[idx] = buf[i]
        return rn.BlockArray(cells)

    A_host, B_host = pinned_blocks(a, ab), pinned_blocks(b, bb)
    A_dev, B_dev = A_host.to_device(), B_host.to_device()
    torch.cuda.synchronize()

    # ---- FP64 tensor peak measured in this run (cuBLAS), before the timed region ----
    peak_tflops, peak_note, peaks_t = None, "not measured", None
    single = dtype in ("float32", "complex64")
    if not args.no_peak_probe and rank == 0:
        peaks_t = measure_tensor_peak(single)
        from rosnet_b200.engine.plan import plan_tensordot as _pt

        _p = _pt(A_dev.grid, A_dev.blockshape, B_dev.grid, B_dev.blockshape, dtype, axes)
        mult = tensor_work_multiplier(dtype, _p.m, _p.n, _p.n_pairs * _p.k, args.precision, args.complex_mode if dtype in ("complex64", "complex128") else "auto")
        peak_tflops = peaks_t["sustained"] / mult
        peak_note = ("raw tensor peak %.1f TFLOP/s sustained (%.1f burst); %s; divided by %.3g = tensor-core flops executed per algorithmic flop; the fraction uses the SUSTAINED figure"
                     % (peaks_t["sustained"], peaks_t["burst"], peaks_t["how"], mult))

    def step_device():
        return np.tensordot(A_dev, B_dev, axes)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps with operands resident in HBM ----
    for _ in range(max(args.warmup, 3)):
        C = step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    before = dict(lib.launch_counter)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        evs[i][0].record()
        C = step_device()
        evs[i][1].record()
    t_end.record()
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    launches = sum(lib.launch_counter.values()) - sum(before.values())
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host (pinned) blocks in, host blocks out, through the public API ----
    def step_e2e():
        Ch = np.tensordot(A_host, B_host, axes).to_host()   # H2D a, b + launch + D2H c + sync
        return Ch

    for _ in range(2):
        Ch = step_e2e()
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        Ch = step_e2e()
    e_end.record()
    barrier()
    e2e_ms = max(e_start.elapsed_time(e_end), (time.perf_counter() - w0) * 1e3)

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    value = world * flops * args.steps / (total_ms * 1e-3) / 1e12
    e2e_value = world * flops * args.steps / (e2e_ms * 1e-3) / 1e12

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- parity spot check of the benchmarked result (not timed) ----
    got = np.asarray(Ch)
    want = np.tensordot(a, b, axes)
    rel = float(np.linalg.norm((got - want).ravel()) / np.linalg.norm(want.ravel()))

    # ---- roofline of the dominant kernel (grouped GEMM): per-launch events in the timed region ----
    gemm_ms = float(np.mean(step_ms))
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    traffic = {}
    try:
        import glob

        for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic.json"))):
            traffic.update(json.load(open(f)))
    except Exception:
        pass
    kname = "gemm_tf32_kernel" if single else "gemm_f64_kernel"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": (achieved / peak_tflops) if peak_tflops else None, "traffic": traffic.get(f"{kname}:{args.workload}"),
                "kernel": ("gemm_tf32_kernel (3xTF32 tcgen05 grouped GEMM, TMEM chunk accumulators; split pre-pass included in the step)" if single
                           else "gemm_f64_kernel (FP64 DMMA grouped GEMM, fused k-block sum)"),
                "algorithmic_flops_per_launch": flops, "avg_launch_ms": gemm_ms, "peak_source": peak_note, "peak_used": "sustained", "raw_tensor_peak": peaks_t}

    # ---- permute kernel against measured HBM (separate, after the timed region) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    # The launch measured is the matricize of operand `a` that `bench_contraction_permute`
    # (axes [(0,2),(1,3)]) issues: all blocks of a, block axes (1,3,0,2) -> K-major GEMM layout, one launch.

    from rosnet_b200.array.device import DeviceStorage, current_stream_ptr
    from rosnet_b200.engine.plan import matricize_desc, plan_tensordot

    perm_axes_spec = [(0, 2), (1, 3)] if len(sa) == 4 else [tuple(range(len(sa) - 1, -1, -1))[:len(axes[0])], axes[1]]
    try:
        pp = plan_tensordot(A_dev.grid, A_dev.blockshape, B_dev.grid, B_dev.blockshape, dtype, perm_axes_spec)
        op = pp.a if not pp.a.direct else (pp.b if not pp.b.direct else None)
        src_arr = A_dev if (op is pp.a) else B_dev
        n_free = len(pp.free_a) if (op is pp.a) else len(pp.free_b)
    except Exception:
        op = None
    roofline_permute = None
    if op is not None:
        p_ext, p_src, p_dst = matricize_desc(src_arr.blockshape, n_free, op.perm, op.ld, op.block_stride, op.nblocks)
        P_out = DeviceStorage(op.ws_elems * itemsize, zero=True)
        stream_ = current_stream_ptr()

        p_desc = lib.permute_desc(itemsize, p_ext, p_src, p_dst)   # built once: the loop below times the kernel, not Python

        def launch_permute():
            lib.permute_launch(p_desc, src_arr._storage.ptr, P_out.ptr, stream_)

        for _ in range(3):
            launch_permute()
        torch.cuda.synchronize()
        n_perm = 20
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(n_perm):
            launch_permute()
        pe1.record()
        torch.cuda.synchronize()
        perm_ms = pe0.elapsed_time(pe1) / n_perm
        # exactness of the benchmarked launch (block 0 of the workspace against numpy)
        blk0 = np.empty((op.rows, op.ld), dtype=dtype)
        lib.memcpy_d2h(blk0.ctypes.data, P_out.ptr, blk0.nbytes, stream_)
        lib.stream_synchronize(stream_)
        want0 = np.transpose(np.asarray(src_arr.data.flat[0]), op.perm).reshape(op.rows, -1)
        perm_ok = bool(np.array_equal(blk0[:, : want0.shape[1]], want0))
        perm_bytes = 2 * src_arr.nbytes
        roofline_permute = {"bound": "hbm", "achieved": perm_bytes / (perm_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": perm_bytes / (perm_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get(f"permute:{args.workload}"),
                            "kernel": "%s (matricize: all %d blocks of the operand into K-major GEMM layout, block axes %s, one launch)" % (
                                "permute_tma_kernel", op.nblocks, tuple(op.perm)),
                            "algorithmic_bytes_per_launch": perm_bytes, "avg_launch_ms": perm_ms, "peak_source": hbm_src, "bit_exact": perm_ok,
                            "timing": "20 back-to-back launches between two CUDA events"}

    # ---- CPU baseline: the oracle on this box's host cores, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        small = flops <= 5e11
        if small:
            with all_host_threads():
                secs = cpu_reference_time(a, ab, b, bb, axes, repeats=3)
            cpu_val, sample = flops / secs / 1e12, "full workload, best of 3 warm runs"
        else:
            # bounded sample: one output block's pair list (n_pairs block GEMMs + adds), extrapolated linearly
            from oracle import blocked_tensordot as orc

            A1, B1 = orc.split_blocks(a, ab), orc.split_blocks(b, bb)
            fa = [i for i in range(len(sa)) if i not in axes[0]]
            fb = [i for i in range(len(sb)) if i not in axes[1]]
            sub_a = A1[tuple(slice(0, 1) if i in fa else slice(None) for i in range(len(sa)))]
            sub_b = B1[tuple(slice(0, 1) if i in fb else slice(None) for i in range(len(sb)))]
            t0 = time.perf_counter()
            orc.blocked_tensordot(sub_a, sub_b, axes, order="C")
            secs = time.perf_counter() - t0
            n_out = int(np.prod([A1.shape[i] for i in fa]) * np.prod([B1.shape[i] for i in fb]))
            cpu_val, sample = (flops / n_out) / secs / 1e12, f"1 of {n_out} output blocks (its whole k-block list), extrapolated linearly"
        cpu = {"value": cpu_val, "unit": "TFLOP/s", "cores": blas_threads(), "kind": "port",
               "sample": sample + f"; numpy {np.__version__}, os.cpu_count={os.cpu_count()}"}

    line = {
        "metric": "contraction TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"float64": "f64", "complex128": "c128", "float32": "f32 (3xTF32)", "complex64": "c64 (3xTF32)"}.get(dtype, dtype).replace(
            "3xTF32", "3xTF32" if args.precision == "default" else "TF32 + 2xBF16 cross terms"), "data": "synthetic",
        "config": {"workload": args.workload, "shape_a": list(sa), "shape_b": list(sb), "blockshape_a": list(ab),
                   "blockshape_b": list(bb), "axes": [list(x) for x in axes], "per_gpu_flops": flops,
                   "complex_mode": args.complex_mode if np.dtype(dtype).kind == "c" else None,
                   "partition": "output blocks (a split along free block axis 0 across ranks, b replicated); no collective",
                   "l2": "operands + result = %d MB per step > 126 MB L2 (inputs larger than L2)" % ((a.nbytes + b.nbytes + want.nbytes) >> 20)},
        "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": int(a.nbytes + b.nbytes),
                "d2h_bytes_per_step": int(want.nbytes), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_permute": roofline_permute,
        "cpu_baseline": cpu, "parity_rel_frobenius": rel,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
