mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 600 python bench.py > gpurun_out/bench_default_noflags.log 2>&1; echo "default(no flags) rc=$?" >> $S
cat $S; tail -n 12 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/bench_default_noflags.log | cut -c1-400
