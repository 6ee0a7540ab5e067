mkdir -p gpurun_out
( for c in 2 1; do
  echo "== cross terms first per k-block, RNB_TF32_CHUNK=$c"
  RNB_TF32_CHUNK=$c timeout 300 python scripts/searched_path_probe.py 2 time 0.61461059008156 0.02284424442732209 1783767436 2>&1 | grep -E "graph ms|rel diff"
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-opt-in --no-peak-probe 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('headline ms/slice', round(d['ms_per_step'],2), 'TFLOP/s', round(d['value'],1), 'result', d['result'])"
) > gpurun_out/tf32_mma_order.txt 2>&1
cut -c1-200 gpurun_out/tf32_mma_order.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -3
