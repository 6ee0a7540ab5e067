mkdir -p gpurun_out
B="python bench.py --workload bench_contraction_c64 --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe --precision tf32_bf16x2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 3 -c 1 -o gpurun_out/prof_gemm_bf16x2 $B > gpurun_out/ncu_gemm_bf16x2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_gemm_bf16x2.ncu-rep
