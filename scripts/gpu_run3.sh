mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python scripts/tf32_probe.py > gpurun_out/tf32_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --workload bench_contraction_c64 > gpurun_out/bench_c64.log 2>&1; echo "bench c64 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --workload bench_contraction_c128 > gpurun_out/bench_c128.log 2>&1; echo "bench c128 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/tf32_probe.log; tail -15 gpurun_out/pytest_gpu.log; tail -c 1800 gpurun_out/bench_c64.log;  tail -c 1500 gpurun_out/bench_c128.log
