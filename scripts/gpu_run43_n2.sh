mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/s3_n2b.log 2>&1; echo "n2 rc=$?"
tail -n 1 gpurun_out/s3_n2b.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); print(json.dumps(d['extra'].get('output_blocks'))[:1500])" || tail -n 30 gpurun_out/s3_n2b.log
