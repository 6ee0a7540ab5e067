mkdir -p gpurun_out
( for kv in "X=0" "RNB_COMPLEX_MODE=4m" "RNB_COMPLEX_MODE=3m" "RNB_PRECISION=tf32_bf16x2" "RNB_SIMT=0"; do
  echo "== $kv"
  env $kv timeout 300 python scripts/searched_path_probe.py 2 time 0.61461059008156 0.02284424442732209 1783767436 2>&1 | grep -E "graph ms|rel diff"
done ) > gpurun_out/searched_tree_accuracy_modes.txt 2>&1
cat gpurun_out/searched_tree_accuracy_modes.txt | cut -c1-220
