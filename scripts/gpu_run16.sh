mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_default.log 2>&1; echo "default rc=$?" >> $S
timeout 600 python bench.py --workload bench_qft --steps 5 --warmup 3 > gpurun_out/bench_qft.log 2>&1; echo "qft rc=$?" >> $S
timeout 600 python bench.py --workload bench_random_tn --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_random_tn.log 2>&1; echo "rtn rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction_f32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f32.log 2>&1; echo "f32 rc=$?" >> $S
timeout 600 python scripts/tn_profile.py bench_qft > gpurun_out/tn_profile_qft.txt 2>&1; echo "tnprof rc=$?" >> $S
cat $S; tail -n 4 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log | cut -c1-600
for f in default qft random_tn f32; do echo "== $f"; tail -n 1 gpurun_out/bench_$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'frac',d['roofline']['frac'], d['roofline'].get('per_family_ms'))
rp=d.get('roofline_permute'); print('permute', rp and (round(rp['achieved']),round(rp['frac'],3)))
"; done
