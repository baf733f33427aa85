#!/bin/bash
SWEEP_L="1000 150" tools/gpu_variant_sweep.sh > gpurun_out/r02l_sweep.txt 2>&1
ISV_NO_ISO=1 SWEEP_LIBS=is_vins_b200/variants/v11_iso.so SWEEP_L="1000 150" tools/gpu_variant_sweep.sh > gpurun_out/r02l_sweep_noiso.txt 2>&1
