mkdir -p gpurun_out
python scripts/apply_one.py 1 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:apply_nd -s 3 -c 2 -o gpurun_out/prof_apply_nd python scripts/apply_one.py 2 > gpurun_out/ncu_apply.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/prof_apply_nd.ncu-rep
