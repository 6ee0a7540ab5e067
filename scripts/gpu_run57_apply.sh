mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_apply.log 2>&1; echo "apply tests rc=$?"; tail -n 5 gpurun_out/pytest_apply.log
timeout 300 python scripts/slice_calls.py bench_sycamore+searched > gpurun_out/slice_calls_sycamore_searched.txt 2>&1; echo rc=$?
awk '$3>0.06' gpurun_out/slice_calls_sycamore_searched.txt | cut -c1-160 | head -70; tail -n 2 gpurun_out/slice_calls_sycamore_searched.txt
