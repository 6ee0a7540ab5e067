mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/scale_n8_graphs.log 2>&1; echo "n8 rc=$?" >> $S
timeout 120 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference > gpurun_out/ref_n8.log 2>&1; echo "ref n8 rc=$?" >> $S
cat $S
tail -n 1 gpurun_out/scale_n8_graphs.log | cut -c1-400; tail -n 1 gpurun_out/ref_n8.log | cut -c1-300
