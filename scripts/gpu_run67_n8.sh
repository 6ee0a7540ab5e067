mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 $TR --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 > gpurun_out/s4_n8.log 2>&1; echo "n8 rc=$?"
tail -n 1 gpurun_out/s4_n8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=8 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); print(json.dumps(d['extra'])[:2500])" || tail -n 30 gpurun_out/s4_n8.log
