mkdir -p gpurun_out
( for c in 2 4 8 2 4 8; do
  RNB_TF32_CHUNK=$c timeout 200 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-opt-in --no-peak-probe 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('chunk $c', 'ms/slice', round(d['ms_per_step'],2), 'TFLOP/s', round(d['value'],1), 'gemm ms', r['per_family_ms'].get('gemm'), 'result', d['result'], 'clocks', d['clocks']['sm_mhz'])" || echo "chunk $c failed"
done ) 2>&1 | tee gpurun_out/tf32_chunk_sweep_s4.txt
