mkdir -p gpurun_out
for mode in auto 3m auto 3m; do
  for wlk in bench_qft bench_schmidt_matrix; do
    RNB_COMPLEX_MODE=$mode python bench.py --workload $wlk --steps 6 --warmup 3 --no-extra --no-cpu-baseline --no-opt-in --no-peak-probe --no-side-configs 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$mode', '$wlk', 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'parity', d.get('parity_rel_frobenius'), d.get('extra',{}).get('schmidt',{}).get('parity_rel_frobenius_all_blocks'), 'result', d.get('result'))" || echo "$mode $wlk failed"
  done
done
RNB_COMPLEX_MODE=3m python scripts/slice_calls.py bench_qft > gpurun_out/slice_calls_qft_3m.txt 2>&1
