mkdir -p gpurun_out
( for c in 1 2 4 8 1000000; do
  echo "== RNB_TF32_CHUNK=$c"
  RNB_TF32_CHUNK=$c timeout 300 python scripts/searched_path_probe.py 2 time 0.61461059008156 0.02284424442732209 1783767436 2>&1 | grep -E "graph ms|rel diff"
done ) > gpurun_out/searched_tree_accuracy_chunk.txt 2>&1
cut -c1-200 gpurun_out/searched_tree_accuracy_chunk.txt
