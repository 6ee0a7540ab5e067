mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/scale_n2_graphs.log 2>&1; echo "n2 rc=$?" >> $S
timeout 300 $TR --nproc-per-node 2 --master-port 29503 bench.py --gpus 2 --steps 5 --warmup 3 --workload bench_qft > gpurun_out/qft_n2_graphs.log 2>&1; echo "qft n2 rc=$?" >> $S
cat $S
for f in scale_n2_graphs qft_n2_graphs; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-500; done
