mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --impl reference > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
for f in ("bench.log","bench_ref.log"):
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"], d.get("roofline",{}).get("frac"), d.get("roofline_permute",{}).get("frac"), d.get("gpu_launches"), d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/"+f).read()[-800:])
PY
