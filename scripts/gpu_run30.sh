mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_default_final.log 2>&1; echo "rc=$?"
tail -n 1 gpurun_out/bench_default_final.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'launches',d['gpu_launches']); print(d['opt_in_precision']); print(d['cpu_baseline'])"
