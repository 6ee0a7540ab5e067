mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-peak-probe --no-opt-in --no-side-configs --no-extra"
$B > gpurun_out/plain_s4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_sycamore_s4.csv $B > gpurun_out/ncu_launches_s4.log 2>&1; echo "launch list rc=$?"
P="python scripts/searched_path_probe.py 2 time 0.61461059008156 0.02284424442732209 1783767436"
$P > gpurun_out/plain_probe.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_sycamore_searched_tree.csv $P > gpurun_out/ncu_launches_searched.log 2>&1; echo "searched launch list rc=$?"
ls -la gpurun_out/*.csv | tail -3
