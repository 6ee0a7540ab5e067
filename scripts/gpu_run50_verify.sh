mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 400 python bench.py > gpurun_out/bench_default_s4.log 2>&1; echo "bench rc=$?" >> $S
cat $S; tail -n 25 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log | cut -c1-200; tail -n 2 gpurun_out/bench_default_s4.log | cut -c1-400
