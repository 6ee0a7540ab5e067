mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tn.py -m gpu -q -p no:cacheprovider -k "searched" > gpurun_out/pytest_searched.log 2>&1; echo "searched test rc=$?"; tail -n 15 gpurun_out/pytest_searched.log
python scripts/apply_one.py 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:apply_nd -s 3 -c 2 -o gpurun_out/prof_apply_nd_v4 python scripts/apply_one.py 2 > gpurun_out/ncu_apply_v4.log 2>&1; echo ncu rc=$?
