mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/s3_multirank.log 2>&1; echo "multirank rc=$?"; tail -n 3 gpurun_out/s3_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29541 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/s3_ref_n2.log 2>&1; echo "ref n2 rc=$?"; tail -n 1 gpurun_out/s3_ref_n2.log | cut -c1-400
timeout 300 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s3_n2.log 2>&1; echo "n2 rc=$?"
tail -n 1 gpurun_out/s3_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); print(json.dumps(d['extra'])[:1500])"
