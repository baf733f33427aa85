python -m pytest tests -m gpu -q > gpurun_out/r02z_gpu_tests.log 2>&1; tail -2 gpurun_out/r02z_gpu_tests.log
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err || tail -5 gpurun_out/r02z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench.json'))
print('value', d['value'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'ceil', d['e2e']['copy_ceiling']['ms_per_step'], d['roofline']['traffic'], d['roofline']['traffic_source'])
print(json.dumps(d['single_window_latency_us']['event_c_abi_us']))
for c in d['configs']:
    print(c['config'], c['value'], c.get('e2e',{}).get('value'))
PY
