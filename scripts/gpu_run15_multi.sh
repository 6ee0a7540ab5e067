mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8 >> $S; nproc >> $S
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/scale_n1.log 2>&1; echo "n1 rc=$?" >> $S
for N in 2 4 8; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/scale_n$N.log 2>&1; echo "n$N rc=$?" >> $S
done
timeout 300 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 8 --warmup 3 --workload bench_random_tn > gpurun_out/rtn_n8.log 2>&1; echo "rtn n8 rc=$?" >> $S
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --workload bench_contraction > gpurun_out/c1_n8.log 2>&1; echo "c1 n8 rc=$?" >> $S
timeout 300 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --workload bench_schmidt_matrix --exchange nccl > gpurun_out/schmidt_nccl_n8.log 2>&1; echo "schmidt nccl n8 rc=$?" >> $S
timeout 300 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --workload bench_schmidt_matrix --exchange fused > gpurun_out/schmidt_fused_n8.log 2>&1; echo "schmidt fused n8 rc=$?" >> $S
cat $S
for f in scale_n1 scale_n2 scale_n4 scale_n8 rtn_n8 c1_n8 schmidt_nccl_n8 schmidt_fused_n8; do echo "== $f"; tail -n 1 gpurun_out/$f.log | cut -c1-600; done
