#!/bin/bash
SWEEP_LIBS="is_vins_b200/variants/v7_accw1.so is_vins_b200/variants/v9_acc3.so is_vins_b200/variants/v9_acc3_r224.so is_vins_b200/variants/v9_acc3_r200.so" SWEEP_L="1000 150" tools/gpu_variant_sweep.sh > gpurun_out/r02h_sweep.txt 2>&1
