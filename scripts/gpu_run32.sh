mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
RNB_TF32_KB=16 timeout 400 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_skinny_splitk.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_kb16.log 2>&1; echo "pytest kb16 rc=$?" >> $S
RNB_TF32_KB=16 RNB_PRECISION=tf32_bf16x2 timeout 400 python -m pytest tests/test_gpu_tensordot.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_kb16_bf16.log 2>&1; echo "pytest kb16 bf16x2 rc=$?" >> $S
cat $S; tail -n 3 gpurun_out/pytest_kb16.log; tail -n 3 gpurun_out/pytest_kb16_bf16.log
for KB in 16 32; do for PR in default tf32_bf16x2; do
RNB_TF32_KB=$KB timeout 200 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline --no-peak-probe --precision $PR 2>&1 | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('kb $KB', d['dtype'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'parity',d.get('parity_rel_frobenius'))"
done; done
RNB_TF32_KB=16 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-peak-probe 2>&1 | tail -n 1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('sycamore kb16 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),d.get('result'), d['opt_in_precision']['ms_per_step'])"
