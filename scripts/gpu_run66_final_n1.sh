mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 600 python bench.py --impl reference > gpurun_out/s4_final_reference.log 2>&1; echo "reference rc=$?" >> $S
timeout 600 python bench.py > gpurun_out/s4_final_default.log 2>&1; echo "bench rc=$?" >> $S
cat $S; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/smoke.log | cut -c1-100; tail -n 1 gpurun_out/s4_final_reference.log | cut -c1-200
tail -n 1 gpurun_out/s4_final_default.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2), 'frac', r['frac'], 'of burst', r['frac_of_burst'], r['per_family_ms'])
print('permute', d['roofline_permute']['frac'], d['roofline_permute']['isolated']['frac'])
st=d['extra'].get('searched_tree'); print({k:v for k,v in st.items() if k in ('log2_flops','ms_per_amplitude','tflops','speedup_time_to_amplitude','rel_diff_vs_headline_tree','per_family_ms','error')}); print(st.get('host_complex128')); print(st.get('roofline_apply'))
for k,v in d['extra'].get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','e2e_value','roofline_frac','parity_rel_frobenius','error')})" || tail -n 30 gpurun_out/s4_final_default.log
