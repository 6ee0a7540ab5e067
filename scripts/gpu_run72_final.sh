mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_k5.log 2>&1; echo "bench k5 rc=$?" >> $S
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/ref_k2.log 2>&1; echo "ref rc=$?" >> $S
cat $S; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log | cut -c1-80
tail -n 1 gpurun_out/bench_k5.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value'],2),'steps',d['steps'],'warmup',d['warmup'],'ms',round(d['ms_per_step'],3),'launches',d['gpu_launches'], 'searched', d['extra']['searched_tree'].get('ms_per_result'))"
tail -n 1 gpurun_out/ref_k2.log | cut -c1-160
