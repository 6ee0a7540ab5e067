mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_apply.log 2>&1; echo "apply tests rc=$?"; tail -n 30 gpurun_out/pytest_apply.log
timeout 600 python scripts/searched_path_probe.py 128 flops > gpurun_out/searched_path_probe2.txt 2>&1; echo rc=$?
cut -c1-700 gpurun_out/searched_path_probe2.txt
