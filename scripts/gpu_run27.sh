mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_tensordot.py -m gpu -q -s -k "3xtf32_accuracy" -p no:cacheprovider > gpurun_out/pytest_acc.log 2>&1; echo "acc rc=$?" >> $S
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 300 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline --precision tf32_bf16x2 > gpurun_out/bench_c64_bf16x2.log 2>&1; echo "c64 bf16x2 rc=$?" >> $S
timeout 300 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c64.log 2>&1; echo "c64 rc=$?" >> $S
cat $S; grep "3xTF32\|TF32+2" gpurun_out/pytest_acc.log | head; tail -n 3 gpurun_out/pytest_acc.log; tail -n 4 gpurun_out/pytest_gpu.log
for f in c64_bf16x2 c64; do tail -n 1 gpurun_out/bench_$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['dtype'],'value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'frac',d['roofline']['frac'],'parity',d['parity_rel_frobenius'])"; done
