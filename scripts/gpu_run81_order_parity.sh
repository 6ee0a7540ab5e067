mkdir -p gpurun_out
( for wl in bench_contraction_c64 bench_contraction_f32; do
  for o in interleaved cross_first; do
    RNB_TF32_ORDER=$o timeout 200 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu-baseline --no-opt-in --no-peak-probe --no-extra 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$wl', '$o', 'TFLOP/s', round(d['value'],1), 'parity_rel_frobenius', d.get('parity_rel_frobenius'))"
  done
done ) 2>&1 | tee gpurun_out/tf32_mma_order_parity.txt
