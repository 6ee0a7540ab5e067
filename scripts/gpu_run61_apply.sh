mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_apply.log 2>&1; echo "apply tests rc=$?"; tail -n 3 gpurun_out/pytest_apply.log
( for j in 1 2 4 auto; do
  if [ $j = auto ]; then unset RNB_APPLY_J; else export RNB_APPLY_J=$j; fi
  echo "== RNB_APPLY_J=$j"
  python scripts/apply_one.py 4 2>&1 | tail -4
  timeout 600 python scripts/searched_path_probe.py 128 flops 2>&1 | grep -E "graph ms|eager ms|rel diff"
done ) > gpurun_out/apply_j_sweep.txt 2>&1
cat gpurun_out/apply_j_sweep.txt
