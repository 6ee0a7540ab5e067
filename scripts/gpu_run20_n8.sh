mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --workload bench_schmidt_matrix --exchange fused > gpurun_out/schmidt_fused_bulk_n8.log 2>&1; echo "schmidt fused bulk n8 rc=$?" >> $S
cat $S
tail -n 1 gpurun_out/schmidt_fused_bulk_n8.log | cut -c1-1500
