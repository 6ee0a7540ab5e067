mkdir -p gpurun_out
S="python scripts/one_slice.py bench_sycamore"
$S > gpurun_out/one_slice_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:permute_tiled_kernel -c 40 -o gpurun_out/prof_permute_split $S > gpurun_out/ncu_permute_split.log 2>&1; echo "ncu1 rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe --no-opt-in --no-side-configs --no-extra"
$B > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_sycamore_s3.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/
