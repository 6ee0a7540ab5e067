mkdir -p gpurun_out
timeout 300 python scripts/gap_probe.py bench_sycamore > gpurun_out/gap_probe_sycamore.txt 2>&1; echo rc=$?
cat gpurun_out/gap_probe_sycamore.txt | cut -c1-400
