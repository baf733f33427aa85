#!/bin/bash
# Times every build variant under is_vins_b200/variants (occupancy / launch-bounds / algorithm sweeps) with
# `bench.py --quick` (device-timed value + per-kernel times).  Build variants with tools/build_variant.sh.
for lib in ${SWEEP_LIBS:-is_vins_b200/variants/*.so}; do
  for L in ${SWEEP_L:-1000}; do
    ISV_B200_LIB=$PWD/$lib python bench.py --quick --steps 10 --warmup 3 --features $L --windows ${SWEEP_W:-9472} | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', $L, '%.0f' % d['value'], {k: round(v, 4) for k, v in d['kernels_ms'].items()})"
  done
done
