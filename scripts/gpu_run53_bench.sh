mkdir -p gpurun_out
timeout 400 python bench.py --no-side-configs > gpurun_out/bench_s4_hot.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench_s4_hot.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2))
print('roofline achieved',r['achieved'],'frac',r['frac'],'frac_of_burst',r['frac_of_burst'],'kernel_ms',r['kernel_ms_per_step'],'share',r['share_of_gpu_time']); print(r['per_family_ms']); print(json.dumps(d['roofline_permute'])[:1500]); print(d['clocks'])" || tail -n 30 gpurun_out/bench_s4_hot.log
