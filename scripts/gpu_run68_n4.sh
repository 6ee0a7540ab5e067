mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 $TR --nproc-per-node 4 --master-port 29544 bench.py --gpus 4 > gpurun_out/s4_n4.log 2>&1; echo "n4 rc=$?"
tail -n 1 gpurun_out/s4_n4.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=4 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); e=d['extra']; print(e['amplitude_vs_n1']['rel_diff'], e['schmidt'].get('nccl_ms'), e['schmidt'].get('fused_ms'), e['schmidt'].get('fused_parity_rel_frobenius_all_blocks'), e['output_blocks'].get('tflops'), e['output_blocks'].get('parity_rel_frobenius_all_blocks'))" || tail -n 30 gpurun_out/s4_n4.log
timeout 300 python -m pytest tests/test_gpu_tn.py tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider -k "searched or multirank or world" > gpurun_out/pytest_n4.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/pytest_n4.log
