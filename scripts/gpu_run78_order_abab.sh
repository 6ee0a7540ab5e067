mkdir -p gpurun_out
( for o in interleaved cross_first interleaved cross_first interleaved cross_first; do
  RNB_TF32_ORDER=$o timeout 200 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-opt-in --no-peak-probe 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$o', 'ms/slice', round(d['ms_per_step'],2), 'TFLOP/s', round(d['value'],1), 'gemm', d['roofline']['per_family_ms'].get('gemm'), 'clocks', d['clocks']['sm_mhz'])"
done ) 2>&1 | tee gpurun_out/tf32_mma_order_abab.txt
