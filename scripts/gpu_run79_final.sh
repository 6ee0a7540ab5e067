mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py > gpurun_out/s4_final_default3.log 2>&1; echo "bench rc=$?" >> $S
cat $S; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log | cut -c1-80
tail -n 1 gpurun_out/s4_final_default3.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2), 'frac', r['frac'], 'burst', r['frac_of_burst'])
st=d['extra']['searched_tree']; print(st['ms_per_result'], st['speedup_time_to_result'], st.get('rel_diff_vs_headline_tree'), st['host_double']['rel_diff'])
for k,v in d['extra'].get('configs',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('searched_tree') or {}).get('ms_per_result'))"
