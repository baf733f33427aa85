#!/bin/bash
python -m pytest tests/test_window_parity_gpu.py tests/test_parity_sweep_gpu.py tests/test_sequence_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02f_tests.log
SWEEP_LIBS="is_vins_b200/variants/v4_accum2.so is_vins_b200/variants/v6_bwd2.so" SWEEP_L="1000" tools/gpu_variant_sweep.sh > gpurun_out/r02f_sweep.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:marg_backward -c 2 -f -o gpurun_out/r02f_bwd python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r02f_ncu.log 2>&1
