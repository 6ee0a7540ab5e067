mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_permute.py tests/test_gpu_large.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_permute_ws.log 2>&1; echo "permute tests rc=$?"; tail -n 4 gpurun_out/pytest_permute_ws.log
timeout 300 python -m pytest tests/test_gpu_tn.py tests/test_gpu_tensordot.py tests/test_gpu_fuzz.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_tn_ws.log 2>&1; echo "tn tests rc=$?"; tail -n 4 gpurun_out/pytest_tn_ws.log
for kv in "RNB_PERMUTE_WS=0" "RNB_PERMUTE_STAGES=2" "RNB_PERMUTE_STAGES=3" "RNB_PERMUTE_STAGES=4" "RNB_PERMUTE_STAGES=3 RNB_PERMUTE_VMAX_KB=24" "RNB_PERMUTE_STAGES=4 RNB_PERMUTE_VMAX_KB=24" "RNB_PERMUTE_STAGES=4 RNB_PERMUTE_VMAX_KB=16" "RNB_PERMUTE_STAGES=6 RNB_PERMUTE_VMAX_KB=16" "RNB_PERMUTE_STAGES=2 RNB_PERMUTE_VMAX_KB=64" "RNB_PERMUTE_WS=0"; do
  env $kv timeout 200 python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline --no-opt-in --no-peak-probe 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; p=d['roofline_permute']
print('$kv', 'ms/slice', round(d['ms_per_step'],2), 'permute_split ms', r['per_family_ms'].get('permute_split'), 'frac', round(p['frac'],3), 'isolated', round(p['isolated']['frac'],3), round(p['isolated']['ms_per_launch'],3))" || echo "$kv failed"
done 2>&1 | tee gpurun_out/permute_ws_sweep.txt
