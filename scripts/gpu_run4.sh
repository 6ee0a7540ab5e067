mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_permute.py -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_permute.log 2>&1; echo "pytest permute rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_tensordot.py -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_tensordot.log 2>&1; echo "pytest tensordot rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload bench_contraction_permute > gpurun_out/bench_perm.log 2>&1; echo "bench perm rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/pytest_permute.log; tail -5 gpurun_out/pytest_tensordot.log; python - <<'PY'
import json
for f in ("bench.log","bench_perm.log"):
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_permute"]["achieved"], d["roofline_permute"]["frac"], d["roofline_permute"]["bit_exact"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
