mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/scale_n2_final.log 2>&1; echo "n2 rc=$?"
tail -n 1 gpurun_out/scale_n2_final.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); print(d['opt_in_precision'])"
