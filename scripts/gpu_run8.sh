mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bench_contraction_c128 --complex-mode 3m > gpurun_out/bench_c128_3m.log 2>&1; echo "c128 3m rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bench_contraction_c128 > gpurun_out/bench_c128_4m.log 2>&1; echo "c128 4m rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bench_contraction_c64 > gpurun_out/bench_c64.log 2>&1; echo "c64 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bench_contraction_f32 > gpurun_out/bench_f32.log 2>&1; echo "f32 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
for f in ("bench_c128_3m","bench_c128_4m","bench_c64","bench_f32","bench"):
    try:
        d=json.loads(open("gpurun_out/"+f+".log").read().strip().splitlines()[-1])
        print(f, "TF", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "frac", d["roofline"]["frac"], "perm", round(d["roofline_permute"]["achieved"]), "rel", d["parity_rel_frobenius"], d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/"+f+".log").read()[-1200:])
PY
