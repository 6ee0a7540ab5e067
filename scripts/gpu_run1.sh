mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_permute.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_permute.log 2>&1; echo "permute rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_tensordot.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_tensordot.log 2>&1; echo "tensordot rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_permute.log; tail -15 gpurun_out/pytest_tensordot.log; tail -3 gpurun_out/bench.log
