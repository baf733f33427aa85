#!/bin/bash
# tools/build_variant.sh NAME [-DMACRO=VALUE ...]  ->  is_vins_b200/variants/NAME.so  (same flags as the product build)
set -e
name=$1; shift
mkdir -p is_vins_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" \
  -shared -o is_vins_b200/variants/$name.so is_vins_b200/csrc/isv_capi.cu
