mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s3_n8.log 2>&1; echo "n8 rc=$?"
tail -n 1 gpurun_out/s3_n8.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=8 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2), d['clocks']); print(json.dumps(d['extra'])[:3000])" || tail -n 30 gpurun_out/s3_n8.log
