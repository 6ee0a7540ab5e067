mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 > gpurun_out/s4_n2_final.log 2>&1; echo "n2 rc=$?"
tail -n 1 gpurun_out/s4_n2_final.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=2 value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2)); e=d['extra']; print(e['amplitude_vs_n1'], e['schmidt'].get('fused_ms'), e['schmidt'].get('fused_parity_rel_frobenius_all_blocks'), e['output_blocks'].get('tflops'))" || tail -n 30 gpurun_out/s4_n2_final.log
