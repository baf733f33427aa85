python -m pytest tests -m gpu -q > gpurun_out/r02z_gpu_tests.log 2>&1; tail -2 gpurun_out/r02z_gpu_tests.log
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err || tail -5 gpurun_out/r02z_bench.err
ncu --set full --clock-control none -k regex:"marg_backward_kernel|marg_forward_tail_kernel" -c 2 -f -o /tmp/r02z_bt python tools/profile_driver.py > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02z_bt.ncu-rep gpurun_out/r02z_ncu_backward_tail_summary.json > /dev/null 2>&1 && echo summary-ok
python tools/sass_census.py > gpurun_out/r02z_sass_census.json 2>/dev/null && echo sass-ok
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench.json'))
print('value', d['value'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['kernels_ms'])
print(json.dumps(d['single_window_latency_us']['event_c_abi_us']), d['parity_max_rel_err'])
for c in d['configs']:
    print(c['config'], c['value'], c.get('e2e',{}).get('value'), c.get('parity',{}).get('max_rel_err'))
s=json.load(open('gpurun_out/r02z_ncu_backward_tail_summary.json'))
for k,v in s.items(): print(k, v['gpu__time_duration.sum'], v['smsp__inst_executed.sum'], v['smsp__issue_active.avg.pct_of_peak_sustained_active'], v['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'], v['launch__registers_per_thread'])
PY
