#!/bin/bash
python -m pytest tests/test_window_parity_gpu.py tests/test_parity_sweep_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02e_tests.log
SWEEP_LIBS="is_vins_b200/variants/v0_base.so is_vins_b200/variants/v4_accum2.so" SWEEP_L="1000 150" tools/gpu_variant_sweep.sh > gpurun_out/r02e_sweep.txt 2>&1
SWEEP_W=11840 tools/gpu_variant_sweep.sh > gpurun_out/r02e_sweep_11840.txt 2>&1
for c in 2 4 8 16; do
  ISV_HOST_CHUNKS=$c python bench.py --no-configs --no-cpu --sustain 0 --parity-windows 8 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($c, d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['abi1']['value'])"
done > gpurun_out/r02e_chunks.txt 2>&1
