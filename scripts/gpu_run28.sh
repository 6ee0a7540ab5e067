mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
RNB_PRECISION=tf32_bf16x2 timeout 600 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_tn.py tests/test_gpu_skinny_splitk.py -m gpu -q -s --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_bf16x2.log 2>&1; echo "pytest bf16x2 rc=$?" >> $S
timeout 300 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline --precision tf32_bf16x2 > gpurun_out/bench_c64_bf16x2.log 2>&1; echo "c64 bf16x2 rc=$?" >> $S
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --precision tf32_bf16x2 > gpurun_out/bench_sycamore_bf16x2.log 2>&1; echo "syc bf16x2 rc=$?" >> $S
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sycamore_3x.log 2>&1; echo "syc 3x rc=$?" >> $S
cat $S; grep "sycamore m=14\|passed\|failed" gpurun_out/pytest_bf16x2.log | tail -5
for f in c64_bf16x2 sycamore_bf16x2 sycamore_3x; do tail -n 1 gpurun_out/bench_$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['dtype'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),d.get('parity_rel_frobenius'),d.get('result'))"; done
