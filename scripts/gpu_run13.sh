mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_default.log 2>&1; echo "default rc=$?" >> $S
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?" >> $S
timeout 600 python bench.py --workload bench_random_tn --steps 8 --warmup 3 > gpurun_out/bench_random_tn.log 2>&1; echo "rtn rc=$?" >> $S
timeout 600 python bench.py --workload bench_qft --steps 5 --warmup 3 > gpurun_out/bench_qft.log 2>&1; echo "qft rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction --steps 10 --warmup 3 > gpurun_out/bench_c1.log 2>&1; echo "c1 rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction_permute --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1p.log 2>&1; echo "c1p rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction_c64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c64.log 2>&1; echo "c64 rc=$?" >> $S
timeout 600 python bench.py --workload bench_contraction_c128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c128.log 2>&1; echo "c128 rc=$?" >> $S
# ncu: launch list of the default bench, then full captures of the new kernels
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-peak-probe"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_sycamore.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $S
BP="python bench.py --workload bench_contraction --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_tma -s 2 -c 1 -o gpurun_out/prof_permute_tma $BP > gpurun_out/ncu_permute_tma.log 2>&1; echo "ncu permute rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 700 -c 40 -o gpurun_out/prof_gemm_tf32_syc $B > gpurun_out/ncu_gemm_syc.log 2>&1; echo "ncu gemm rc=$?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_kernel -c 2 -o gpurun_out/prof_skinny $B > gpurun_out/ncu_skinny.log 2>&1; echo "ncu skinny rc=$?" >> $S
cat $S; tail -n 4 gpurun_out/pytest_gpu.log
for f in default reference random_tn qft c1 c1p c64 c128; do echo "== $f"; tail -n 1 gpurun_out/bench_$f.log | cut -c1-2500; done
ls -la gpurun_out | tail -12
