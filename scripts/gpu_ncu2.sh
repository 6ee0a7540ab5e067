mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe"
$B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:micro8 -s 2 -c 2 -o gpurun_out/prof_permute_micro $B > gpurun_out/ncu_permute_micro.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe --workload bench_contraction_c64"
$B2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 3 -c 2 -o gpurun_out/prof_gemm_tf32 $B2 > gpurun_out/ncu_gemm_tf32.log 2>&1
$B2 > gpurun_out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c64.csv $B2 > gpurun_out/ncu_launches_c64.log 2>&1
ls -la gpurun_out | tail -8
