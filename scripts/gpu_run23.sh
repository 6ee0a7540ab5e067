mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_permute.py tests/test_gpu_tensordot.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 300 python scripts/permute_probe.py child > gpurun_out/permute_probe_f32.txt 2>&1; echo "probe rc=$?" >> $S
timeout 300 python bench.py --workload bench_contraction_f32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f32.log 2>&1; echo "f32 rc=$?" >> $S
cat $S; tail -n 6 gpurun_out/pytest_gpu.log; cat gpurun_out/permute_probe_f32.txt; tail -n 1 gpurun_out/bench_f32.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],2),'frac',d['roofline']['frac']); rp=d['roofline_permute']; print('permute',round(rp['achieved']),round(rp['frac'],3),rp['bit_exact'],rp['kernel'][:40])"
