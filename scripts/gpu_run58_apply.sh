mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_apply.log 2>&1; echo "apply tests rc=$?"; tail -n 5 gpurun_out/pytest_apply.log
timeout 600 python scripts/searched_path_probe.py 128 flops > gpurun_out/searched_path_probe3.txt 2>&1; echo rc=$?
cut -c1-400 gpurun_out/searched_path_probe3.txt
timeout 300 python scripts/slice_calls.py bench_sycamore+searched > gpurun_out/slice_calls_sycamore_searched.txt 2>&1; echo rc=$?
awk '$3>0.07' gpurun_out/slice_calls_sycamore_searched.txt | cut -c1-160 | head -40; tail -n 2 gpurun_out/slice_calls_sycamore_searched.txt
