mkdir -p gpurun_out
G="python scripts/gemm_one.py --dtype complex64 --m 8192 --n 32768 --k 4096 --reps 2"
$G 2>&1 | tail -1 | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_v2_kernel -s 1 -c 1 -o gpurun_out/prof_gemm_tf32_3m_s4 $G > gpurun_out/ncu_gemm_tf32_s4.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_gemm_tf32_3m_s4.ncu-rep
