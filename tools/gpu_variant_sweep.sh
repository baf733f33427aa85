#!/bin/bash
# times every build variant under is_vins_b200/variants (occupancy / launch-bounds sweep)
for lib in is_vins_b200/variants/*.so; do
  for L in ${SWEEP_L:-1000}; do
    ISV_B200_LIB=$PWD/$lib python bench.py --steps 10 --warmup 3 --no-cpu --features $L | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', $L, '%.0f' % d['value'], d['kernels_ms'])"
  done
done
