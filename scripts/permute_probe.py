"""Permute kernel bandwidth under different tile knobs (GPU box). One subprocess per setting."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHAPES = {
    "c1_f64": ((16, 32, 32, 32, 32), (0, 2, 4, 1, 3), 8),
    "t2d_f64": ((8192, 8192), (1, 0), 8),
    "tn_c64_r24": ((2,) * 24, (23, 3, 17, 0, 9, 21, 5, 10, 2, 18, 4, 6, 1, 20, 7, 22, 8, 11, 13, 12, 15, 14, 16, 19), 8),
    "tn_c64_contraction_like": ((2,) * 24, tuple(i for i in range(24) if i not in (1, 4, 5, 9, 12, 13, 17, 20, 22)) + (1, 4, 5, 9, 12, 13, 17, 20, 22), 8),
    "tn_c64_bond16_r6": ((16,) * 6, (0, 3, 5, 1, 2, 4), 8),
    "blockify_c128": ((8, 256, 8, 256), (0, 2, 1, 3), 16),
    "c1_c128": ((16, 32, 32, 32, 32), (0, 2, 4, 1, 3), 16),
    "t2d_f32": ((8192, 8192), (1, 0), 4),
    "t2d_c128": ((4096, 8192), (1, 0), 16),
}
if len(sys.argv) > 1:
    import numpy as np, torch
    from rosnet_b200.engine import lib
    lib.load()
    out = {}
    for name, (shape, perm, es) in SHAPES.items():
        n = int(np.prod(shape))
        src = torch.empty(n * es, dtype=torch.uint8, device="cuda").random_()
        dst = torch.empty(n * es, dtype=torch.uint8, device="cuda")
        ss = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
        oshape = [shape[p] for p in perm]
        os_ = [int(np.prod(oshape[i + 1:])) for i in range(len(oshape))]
        ds = [0] * len(shape)
        for j, p in enumerate(perm):
            ds[p] = os_[j]
        st = torch.cuda.current_stream().cuda_stream
        pd = lib.permute_desc(es, list(shape), ss, ds)
        run = lambda: lib.permute_launch(pd, src.data_ptr(), dst.data_ptr(), st)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[name] = round(2 * n * es / (ms * 1e-3) / 1e9)
    print(json.dumps(out))
else:
    settings = [{}, {"RNB_PERMUTE_KERNEL": "tiled"}, {"RNB_PERMUTE_KERNEL": "tiled", "RNB_PERMUTE_LOAD": "cpasync"}, {"RNB_PERMUTE_KERNEL": "generic"}] if "--short" in sys.argv[0:1] or os.environ.get("PROBE_SHORT") else [{}, {"RNB_PERMUTE_OD": "src"}, {"RNB_PERMUTE_TARGET_OUT": "128"}, {"RNB_PERMUTE_TARGET_OUT": "256", "RNB_PERMUTE_VMAX_KB": "64"},
                {"RNB_PERMUTE_TARGET_OUT": "128", "RNB_PERMUTE_TARGET_IN": "128", "RNB_PERMUTE_VMAX_KB": "64"},
                {"RNB_PERMUTE_STAGES": "2"}, {"RNB_PERMUTE_STAGES": "4"}, {"RNB_PERMUTE_KERNEL": "tiled"},
                {"RNB_PERMUTE_TARGET_OUT": "128", "RNB_PERMUTE_STAGES": "2"}, {"RNB_PERMUTE_TARGET_OUT": "256", "RNB_PERMUTE_VMAX_KB": "64", "RNB_PERMUTE_STAGES": "2"}]
    for s in settings:
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **s), capture_output=True, text=True)
        print(json.dumps(s), r.stdout.strip() or r.stderr[-400:])
