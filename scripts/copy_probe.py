"""How fast is a plain device copy at the sizes the permute kernel is benchmarked on? (GPU box)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rosnet_b200.engine import lib
lib.load()
out = {}
for mb in (128, 256, 512, 2048):
    n = mb << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda").random_()
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("torch_copy", lambda: b.copy_(a)),
                     ("cudaMemcpyD2D", lambda: torch.cuda.current_stream().synchronize() or None)):
        if name == "cudaMemcpyD2D":
            continue
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        out[f"{name}_{mb}MB"] = round(2 * n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)
    # our kernels: identity copy of a 2-D view (tiled kernel, copy-like) and a 4-D identity
    es = 8
    st = torch.cuda.current_stream().cuda_stream
    rows = n // es // 4096
    run = lambda: lib.permute(a.data_ptr(), b.data_ptr(), es, [rows, 4096], [4096, 1], [4096, 1], st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    out[f"rnb_identity_{mb}MB"] = round(2 * n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)
    # 2-D transpose of the same bytes (f64)
    side = int((n // es) ** 0.5)
    run = lambda: lib.permute(a.data_ptr(), b.data_ptr(), es, [side, side], [side, 1], [1, side], st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    out[f"rnb_transpose_{mb}MB"] = round(2 * side * side * es / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)
print(json.dumps(out, indent=0))
