"""Accuracy and speed of the 3xTF32 kernel vs the TMEM chunk length (run on the GPU box)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch
    import rosnet_b200 as rn
    rng = np.random.default_rng(11)
    out = {"chunk": os.environ.get("RNB_TF32_CHUNK")}
    for dtype in ("float32", "complex64"):
        m, n, k = 384, 320, 4096
        a = rng.standard_normal((m, k)); b = rng.standard_normal((k, n))
        if dtype == "complex64":
            a = a + 1j * rng.standard_normal((m, k)); b = b + 1j * rng.standard_normal((k, n))
        a, b = a.astype(dtype), b.astype(dtype)
        wide = "complex128" if dtype == "complex64" else "float64"
        truth = a.astype(wide) @ b.astype(wide)
        A = rn.array(a, blockshape=(192, 2048), inner="cuda"); B = rn.array(b, blockshape=(2048, 160), inner="cuda")
        got = np.asarray(np.tensordot(A, B, [(1,), (0,)])).astype(wide)
        out[dtype + "_err"] = float(np.linalg.norm(got - truth) / np.linalg.norm(truth))
        # speed: one 4096^3 block pair
        M = 4096
        x = rn.array(rng.standard_normal((M, M)).astype(dtype), inner="cuda"); y = rn.array(rng.standard_normal((M, M)).astype(dtype), inner="cuda")
        for _ in range(2): np.tensordot(x, y, [(1,), (1,)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): np.tensordot(x, y, [(1,), (1,)])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[dtype + "_tflops"] = (8 if dtype == "complex64" else 2) * M**3 / (ms * 1e-3) / 1e12
    print(json.dumps(out))
else:
    for ch in ("1", "2", "4", "8", "16", "1000000"):
        env = dict(os.environ, RNB_TF32_CHUNK=ch)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:])
