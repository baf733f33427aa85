python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err || tail -20 gpurun_out/r02v_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v_bench.json'))
print('value', d['value'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'abi2', d['e2e']['abi2']['value'], 'abi1', d['e2e']['abi1']['value'], 'ceil', d['e2e']['copy_ceiling']['ms_per_step'])
print(json.dumps(d['single_window_latency_us']))
print(d['kernels_ms'], d['parity_max_rel_err'], d['gpu_launches'])
for c in d['configs']:
    print(c['config'], c['value'], c.get('e2e',{}).get('value'), c.get('parity',{}).get('max_rel_err'))
PY
