"""Per-launch table of one tensor-network step (round-2 entry points): every C-ABI call bracketed by CUDA events,
printed with its shape, time, achieved rate and the two floors that bound it (tensor work at the executed-flop
multiplier, HBM bytes of operands + planes + result).

    python scripts/slice_calls.py bench_sycamore > gpurun_out/slice_calls.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from rosnet_b200 import tn as T  # noqa: E402
from rosnet_b200.engine import lib  # noqa: E402

HBM = 6552e9
TENSOR = {True: 747e12, False: 35.5e12}  # cuBLAS burst on these boxes: TF32 (single-precision dtypes), DGEMM (float64 / complex128)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "bench_sycamore"
    lib.load()
    searched = name.endswith("+searched")       # bench_sycamore+searched: the tree of tn.search_path instead of the headline one
    name = name.replace("+searched", "")
    spec = bench.TN_WORKLOADS[name]
    net, path, sliced = bench.build_tn(spec)
    if searched:
        path, sliced, info = T.search_path(net.inputs, net.output, net.sizes, trials=2, seed=0, target_size=2 ** (spec["target_log2"] or 28), minimize="time",
                                           candidates=spec.get("tree_candidates", ()), dtype=spec["dtype"])
    loop = spec["mode"] == "loop" and bool(sliced)

    def one():
        if loop:
            return T.contract_network(net, path, sliced=sliced, slice_mode="loop", slice_ids=[0])[0]
        return np.asarray(T.contract_network(net, path, sliced=sliced, slice_mode="blocks")[0])

    one()
    one()
    torch.cuda.synchronize()
    rows = []
    o_contract, o_split, o_permute, o_batch, o_nd = lib.contract_grouped, lib.permute_split, lib.permute_launch, lib.contract_nd_batch, lib.contract_nd

    def ev():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def contract(d, stream):
        e0, e1 = ev()
        route = lib.contract_route(d)
        e0.record()
        o_contract(d, stream)
        e1.record()
        cplx = d.dtype in (lib.RNB_C64, lib.RNB_C128)
        eb = {lib.RNB_F32: 4, lib.RNB_F64: 8, lib.RNB_C64: 8, lib.RNB_C128: 16}[d.dtype]
        k_total = d.k * d.n_pairs
        fl = (8 if cplx else 2) * d.m * d.n * k_total * d.n_out
        dn = {lib.RNB_F32: "float32", lib.RNB_F64: "float64", lib.RNB_C64: "complex64", lib.RNB_C128: "complex128"}[d.dtype]
        mult = bench.tensor_work_multiplier(dn, d.m, d.n, k_total, "default")
        obytes = eb * d.m * d.n * d.n_out
        abytes = eb * d.m * d.k * d.a_nblocks
        bbytes = eb * d.n * d.k * d.b_nblocks
        rows.append(dict(kind=("gemm", "skinny", "small")[route], e=(e0, e1), m=d.m, n=d.n, k=d.k, n_out=d.n_out, pairs=d.n_pairs, flags=d.flags,
                         flops=fl, mult=mult, bytes=abytes + bbytes + obytes, maj=f"{d.a_major}{d.b_major}", single=d.dtype in (lib.RNB_F32, lib.RNB_C64)))

    def permute_split(d, src, planes, ws_ptr, r, k, nblocks, stream):
        e0, e1 = ev()
        e0.record()
        ok = o_split(d, src, planes, ws_ptr, r, k, nblocks, stream)
        e1.record()
        if ok:
            n = 1
            for i in range(d.rank):
                n *= int(d.extent[i])
            written = {lib.RNB_SPLIT_PLANES: 2, lib.RNB_SPLIT_C64_B4M: 4, lib.RNB_SPLIT_C64_3M: 3}[planes.mode]
            rows.append(dict(kind=f"permute_split(mode {planes.mode})", e=(e0, e1), bytes=d.elem_bytes * n * (1 + written), rank=d.rank))
        return ok

    def permute(d, src, dst, stream):
        e0, e1 = ev()
        e0.record()
        o_permute(d, src, dst, stream)
        e1.record()
        n = 1
        for i in range(d.rank):
            n *= int(d.extent[i])
        rows.append(dict(kind="permute", e=(e0, e1), bytes=2 * d.elem_bytes * n, rank=d.rank))

    def batch(descs, stream):
        e0, e1 = ev()
        e0.record()
        o_batch(descs, stream)
        e1.record()
        rows.append(dict(kind=f"small_batch[{len(descs)}]", e=(e0, e1), bytes=0))

    def nd(d, stream):
        e0, e1 = ev()
        e0.record()
        o_nd(d, stream)
        e1.record()
        m = int(np.prod([d.ext_m[i] for i in range(d.rank_m)])) if d.rank_m else 1
        n = int(np.prod([d.ext_n[i] for i in range(d.rank_n)])) if d.rank_n else 1
        k = int(np.prod([d.ext_k[i] for i in range(d.rank_k)])) if d.rank_k else 1
        eb = {lib.RNB_F32: 4, lib.RNB_F64: 8, lib.RNB_C64: 8, lib.RNB_C128: 16}[d.dtype]
        rows.append(dict(kind="nd m=%d n=%d k=%d" % (m, n, k), e=(e0, e1), bytes=eb * d.n_out * (d.n_pairs * (m * k + n * k) + m * n), rank=max(d.rank_m, d.rank_n)))

    lib.contract_grouped, lib.permute_split, lib.permute_launch, lib.contract_nd_batch, lib.contract_nd = contract, permute_split, permute, batch, nd
    try:
        s0, s1 = ev()
        s0.record()
        one()
        s1.record()
        torch.cuda.synchronize()
    finally:
        lib.contract_grouped, lib.permute_split, lib.permute_launch, lib.contract_nd_batch, lib.contract_nd = o_contract, o_split, o_permute, o_batch, o_nd
    total = 0.0
    print(f"# {name}: {len(rows)} C-ABI calls in one step, {s0.elapsed_time(s1):.2f} ms first-to-last (eager, event-bracketed)")
    print("#  idx  kind                      ms      shape / bytes                                 achieved           floor_tensor_ms  floor_hbm_ms  ms/floor")
    for i, r in enumerate(rows):
        ms = r["e"][0].elapsed_time(r["e"][1])
        total += ms
        if ms < 0.03:
            continue
        if "flops" in r:
            ft = r["flops"] * r["mult"] / TENSOR[r["single"]] * 1e3 if r["kind"] == "gemm" else 0.0
            fh = r["bytes"] / HBM * 1e3
            fl = max(ft, fh)
            print(f"  {i:4d}  {r['kind']:<24s} {ms:8.3f}  m={r['m']} n={r['n']} k={r['k']} n_out={r['n_out']} pairs={r['pairs']} maj={r['maj']} flags={r['flags']}"
                  f"  {r['flops'] / ms / 1e9:7.1f} TFLOP/s  {ft:8.3f} {fh:8.3f}  {ms / fl:5.2f}")
        else:
            fh = r["bytes"] / HBM * 1e3
            print(f"  {i:4d}  {r['kind']:<24s} {ms:8.3f}  {r['bytes'] / 1e6:9.1f} MB rank={r.get('rank', 0)}  {r['bytes'] / ms / 1e6:7.0f} GB/s  {'':8s} {fh:8.3f}  {ms / fh if fh else 0:5.2f}")
    print(f"# sum of bracketed calls {total:.2f} ms")


if __name__ == "__main__":
    main()
