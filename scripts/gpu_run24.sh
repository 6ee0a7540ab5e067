mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 300 python scripts/permute_probe.py child > gpurun_out/permute_probe_copy.txt 2>&1; echo "probe rc=$?" >> $S
RNB_PERMUTE_KERNEL=tiled timeout 300 python scripts/permute_probe.py child > gpurun_out/permute_probe_copy_tiled.txt 2>&1; echo "probe tiled rc=$?" >> $S
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; echo "default rc=$?" >> $S
cat $S; tail -n 6 gpurun_out/pytest_gpu.log; cat gpurun_out/permute_probe_copy.txt gpurun_out/permute_probe_copy_tiled.txt; tail -n 1 gpurun_out/bench_default.log | cut -c1-200
