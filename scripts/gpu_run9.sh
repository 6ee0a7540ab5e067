mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/tn_profile.py bench_sycamore bench_random_tn > gpurun_out/tn_profile.txt 2>&1; echo "tn_profile rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_gpu.log; head -60 gpurun_out/tn_profile.txt
