mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multirank_s4.log 2>&1; echo "multirank rc=$?"; tail -n 3 gpurun_out/pytest_multirank_s4.log
timeout 400 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 > gpurun_out/s4_n2.log 2>&1; echo "n2 rc=$?"
tail -n 1 gpurun_out/s4_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); print(json.dumps(d['extra'])[:3000])" || tail -n 30 gpurun_out/s4_n2.log
timeout 300 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/s4_ref_n2.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/s4_ref_n2.log | cut -c1-300
