mkdir -p gpurun_out
timeout 600 python scripts/searched_path_probe.py 2 time 0.61461059008156 0.02284424442732209 1783767436 > gpurun_out/searched_path_probe_time_tree.txt 2>&1; echo rc=$?
cut -c1-500 gpurun_out/searched_path_probe_time_tree.txt
