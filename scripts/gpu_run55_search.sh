mkdir -p gpurun_out
timeout 600 python scripts/searched_path_probe.py 128 flops > gpurun_out/searched_path_probe.txt 2>&1; echo rc=$?
cut -c1-600 gpurun_out/searched_path_probe.txt
python scripts/slice_calls.py bench_sycamore_searched > gpurun_out/slice_calls_searched.txt 2>&1 || true
