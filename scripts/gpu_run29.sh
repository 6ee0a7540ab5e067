mkdir -p gpurun_out
for C in 3 4 6; do
RNB_TF32_CHUNK=$C timeout 200 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline --no-peak-probe --precision tf32_bf16x2 2>&1 | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('chunk $C', d['dtype'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'parity',d.get('parity_rel_frobenius'))"
done
RNB_TF32_CHUNK=3 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-peak-probe --precision tf32_bf16x2 2>&1 | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('sycamore chunk 3', 'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),d.get('result'))"
