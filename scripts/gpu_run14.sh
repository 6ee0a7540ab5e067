mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction --steps 10 --warmup 3 > gpurun_out/bench_c1.log 2>&1; echo "c1 rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction_c128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c128.log 2>&1; echo "c128 rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c64.log 2>&1; echo "c64 rc=$?" >> $S
timeout 600 python scripts/permute_probe2.py > gpurun_out/permute_probe3.txt 2>&1; echo "probe rc=$?" >> $S
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_default.log 2>&1; echo "default rc=$?" >> $S
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-peak-probe"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 38 -c 19 -o gpurun_out/prof_gemm_tf32_syc $B > gpurun_out/ncu_gemm_syc.log 2>&1; echo "ncu gemm rc=$?" >> $S
BP="python bench.py --workload bench_contraction --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_tma -s 2 -c 1 -o gpurun_out/prof_permute_tma $BP > gpurun_out/ncu_permute_tma.log 2>&1; echo "ncu permute rc=$?" >> $S
cat $S; tail -n 4 gpurun_out/pytest_gpu.log; cat gpurun_out/permute_probe3.txt
for f in c1 c128 c64 default; do echo "== $f"; tail -n 1 gpurun_out/bench_$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'frac',d['roofline']['frac'])
print('permute',d.get('roofline_permute'))
"; done
