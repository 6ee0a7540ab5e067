mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/s4_reconf_default.log 2>&1; echo "bench rc=$? wall ${SECONDS}s"
tail -n 1 gpurun_out/s4_reconf_default.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2), 'frac', r['frac'])
st=d['extra'].get('searched_tree'); print({k:v for k,v in st.items() if k not in ('what','per_family_calls')})
for k,v in d['extra'].get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','error','searched_tree')})" || tail -n 30 gpurun_out/s4_reconf_default.log
timeout 300 python -m pytest tests/test_gpu_tn.py -m gpu -q -p no:cacheprovider -k "searched" 2>&1 | tail -3
