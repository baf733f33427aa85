python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err || tail -20 gpurun_out/r02z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench.json'))
e=d['e2e']
print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], e['input_abi'], 'ceil', e['copy_ceiling']['ms_per_step'], 'abi3', e['abi3']['value'])
for c in d['configs']:
    print(c['config'], c['value'], c.get('e2e',{}).get('value'), c.get('e2e',{}).get('ms_per_step'), c.get('e2e',{}).get('input_abi'))
PY
