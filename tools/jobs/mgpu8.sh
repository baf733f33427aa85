python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --no-configs > gpurun_out/r02z_bench_8gpu.json 2> gpurun_out/r02z_bench_8gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench_8gpu.json')); e=d['e2e']
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', e['value'], e['ms_per_step'], 'ceil', e['copy_ceiling']['ms_per_step'], e['copy_ceiling']['value'], 'abi3', e['abi3']['value'], 'abi2', e['abi2']['value'], 'abi1', e['abi1']['value'])
PY
