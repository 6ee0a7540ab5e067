mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
$B > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 3 -c 2 -o gpurun_out/prof_gemm_f64 $B > gpurun_out/ncu_gemm.log 2>&1
$B > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:permute_tiled -s 2 -c 2 -o gpurun_out/prof_permute $B > gpurun_out/ncu_permute.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu_gemm.log; tail -3 gpurun_out/ncu_permute.log
