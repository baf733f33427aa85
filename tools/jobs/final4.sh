python -m pytest tests -m gpu -q > gpurun_out/r02z_gpu_tests.log 2>&1; tail -2 gpurun_out/r02z_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --no-configs > gpurun_out/r02z_bench_2gpu.json 2> gpurun_out/r02z_bench_2gpu.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench_2gpu.json')); e=d['e2e']
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', e['value'], e['ms_per_step'], 'ceil', e['copy_ceiling']['ms_per_step'], 'abi3', e['abi3']['value'])
PY
python tools/sass_census.py > gpurun_out/r02z_sass_census.json 2>/dev/null && echo sass-ok
