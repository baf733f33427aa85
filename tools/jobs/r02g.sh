#!/bin/bash
SWEEP_LIBS="is_vins_b200/variants/v6_bwd2.so is_vins_b200/variants/v7_accw1.so is_vins_b200/variants/v7_accw2.so is_vins_b200/variants/v7_accw4.so is_vins_b200/variants/v8_accw1_mb9.so is_vins_b200/variants/v8_accw1_mb10.so is_vins_b200/variants/v8_accw1_mb12.so" SWEEP_L="1000" tools/gpu_variant_sweep.sh > gpurun_out/r02g_sweep.txt 2>&1
