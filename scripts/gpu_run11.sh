mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/permute_probe2.py > gpurun_out/permute_probe2.txt 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload bench_sycamore --steps 8 --warmup 2 > gpurun_out/bench_sycamore.log 2>&1; echo "syc rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload bench_random_tn --steps 8 --warmup 2 > gpurun_out/bench_random_tn.log 2>&1; echo "rtn rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 30 gpurun_out/pytest_gpu.log; cat gpurun_out/permute_probe2.txt; tail -n 1 gpurun_out/bench_sycamore.log | cut -c1-330;  tail -n 1 gpurun_out/bench_random_tn.log | cut -c1-330
