#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gpu_tests.log 2>&1; tail -5 gpurun_out/r02s_gpu_tests.log
python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err || tail -20 gpurun_out/r02s_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02s_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'abi2', d['e2e']['abi2']['value'], 'abi1', d['e2e']['abi1']['value'], 'ceil', d['e2e']['copy_ceiling'])
print(d['single_window_latency_us'], d['kernels_ms'], d['parity_max_rel_err'])
PY
