mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_apply.log 2>&1; echo "apply tests rc=$?"; tail -n 3 gpurun_out/pytest_apply.log
for j in 1 2 auto; do
  if [ $j = auto ]; then unset RNB_APPLY_J; else export RNB_APPLY_J=$j; fi
  echo "== RNB_APPLY_J=$j"
  timeout 600 python scripts/searched_path_probe.py 128 flops 2>&1 | grep -E "graph ms|eager ms|families"
done
unset RNB_APPLY_J
timeout 300 python scripts/slice_calls.py bench_sycamore+searched > gpurun_out/slice_calls_sycamore_searched.txt 2>&1; echo rc=$?
awk '$3>0.07' gpurun_out/slice_calls_sycamore_searched.txt | cut -c1-160 | sed -n 1,8p; awk '$3>0.2' gpurun_out/slice_calls_sycamore_searched.txt | cut -c1-160; tail -n 2 gpurun_out/slice_calls_sycamore_searched.txt
