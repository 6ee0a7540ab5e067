mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_tn.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_tn.log 2>&1; echo "pytest tn rc=$?" >> gpurun_out/summary.txt
for w in bench_qft bench_sycamore bench_random_tn; do
timeout 900 python bench.py --steps 3 --warmup 1 --workload $w > gpurun_out/$w.log 2>&1; echo "$w rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_tn.log; python - <<'PY'
import json
for f in ("bench_qft","bench_sycamore","bench_random_tn"):
    try:
        d=json.loads(open("gpurun_out/"+f+".log").read().strip().splitlines()[-1])
        print(f, "TF", round(d["value"],2), "ms/step", round(d["ms_per_step"],1), "launches", d["gpu_launches"], d["config"], d.get("cpu_baseline"), d.get("result"))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/"+f+".log").read()[-1500:])
PY
