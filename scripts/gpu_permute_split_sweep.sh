# sweep of the tiled permute kernel's tile knobs on the fused matricize + split launches of a Sycamore slice
for kv in "X=0" "RNB_PERMUTE_VMAX_KB=24" "RNB_PERMUTE_VMAX_KB=64" "RNB_PERMUTE_VMAX_KB=96" "RNB_PERMUTE_TARGET_OUT=128" "RNB_PERMUTE_TARGET_IN=128" "RNB_PERMUTE_TARGET_IN=128 RNB_PERMUTE_TARGET_OUT=128 RNB_PERMUTE_VMAX_KB=64" "RNB_PERMUTE_LOAD=cpasync"; do
  env $kv python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline --no-opt-in --no-peak-probe 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; p=d['roofline_permute']
print('$kv', 'ms/slice', round(d['ms_per_step'],2), 'permute_split ms', r['per_family_ms'].get('permute_split'), 'frac', round(p['frac'],3))"
done
