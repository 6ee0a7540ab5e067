mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_permute.py tests/test_gpu_skinny_splitk.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 900 python scripts/permute_probe2.py > gpurun_out/permute_probe2.txt 2>&1; echo "probe rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 8 gpurun_out/pytest_gpu.log; cat gpurun_out/permute_probe2.txt
