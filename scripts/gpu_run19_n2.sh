mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29504 bench.py --gpus 2 --steps 3 --warmup 3 --workload bench_schmidt_small --exchange fused > gpurun_out/schmidt_small_bulk_n2.log 2>&1; echo "small bulk rc=$?" >> $S
RNB_PEER_BULK=0 timeout 200 $TR --nproc-per-node 2 --master-port 29505 bench.py --gpus 2 --steps 3 --warmup 3 --workload bench_schmidt_small --exchange fused > gpurun_out/schmidt_small_red_n2.log 2>&1; echo "small red rc=$?" >> $S
timeout 300 $TR --nproc-per-node 2 --master-port 29506 bench.py --gpus 2 --steps 3 --warmup 3 --workload bench_schmidt_matrix --exchange fused > gpurun_out/schmidt_bulk_n2.log 2>&1; echo "bulk rc=$?" >> $S
cat $S
for f in schmidt_small_bulk_n2 schmidt_small_red_n2 schmidt_bulk_n2; do echo "== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-300; grep -o '"exchange_ms": [0-9.e-]*\|"parity_rel_frobenius": [0-9.e-]*\|"avg_launch_ms": [0-9.e-]*' gpurun_out/$f.log | tr '\n' ' '; echo; done
