"""Permute bandwidth: default kernel selection vs the older kernels, and a plain device copy of the
same bytes as the practical ceiling (GPU box).  One subprocess per setting."""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
if len(sys.argv) > 1 and sys.argv[1] == "copy":
    import torch
    out = {}
    for name, nbytes in (("128MiB", 128 << 20), ("256MiB", 256 << 20), ("512MiB", 512 << 20)):
        a = torch.empty(nbytes, dtype=torch.uint8, device="cuda").random_()
        b = torch.empty_like(a)
        for _ in range(3): b.copy_(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): b.copy_(a)
        e1.record(); torch.cuda.synchronize()
        out[name] = round(2 * nbytes / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)
    print(json.dumps(out))
else:
    r = subprocess.run([sys.executable, __file__, "copy"], capture_output=True, text=True)
    print("torch copy", r.stdout.strip() or r.stderr[-400:])
    settings = [{}, {"RNB_TMA_INTERLEAVE": "0"}, {"RNB_TMA_IN": "4", "RNB_TMA_OUT": "2"}, {"RNB_TMA_IN": "2", "RNB_TMA_OUT": "2", "RNB_TMA_CTAS": "3"},
                {"RNB_TMA_IN": "3", "RNB_TMA_OUT": "4"}, {"RNB_TMA_IN": "2", "RNB_TMA_OUT": "2", "RNB_TMA_CTAS": "2"},
                {"RNB_TMA_IN": "6", "RNB_TMA_OUT": "4", "RNB_TMA_CTAS": "1"}, {"RNB_PERMUTE_KERNEL": "micro"}]
    for s in settings:
        r = subprocess.run([sys.executable, os.path.join(HERE, "permute_probe.py"), "child"], env=dict(os.environ, **s), capture_output=True, text=True)
        print(json.dumps(s), r.stdout.strip() or r.stderr[-600:])
