for L in 1000 150; do python tools/e2e_chunks.py $L 9472; python tools/e2e_chunks.py $L 9472; done
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err || tail -20 gpurun_out/r02z_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench.json'))
e=d['e2e']
print('value', d['value'], 'e2e', e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], 'ceil', e['copy_ceiling']['ms_per_step'], 'abi3', e['abi3']['value'], 'abi2', e['abi2']['value'], 'abi1', e['abi1']['value'], e['max_rel_diff_vs_device_path'])
for c in d['configs']:
    print(c['config'], c['value'], c.get('e2e',{}).get('value'), c.get('e2e',{}).get('ms_per_step'))
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
