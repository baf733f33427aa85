#!/bin/bash
# occupancy experiment: per-kernel time at reduced residency (extra dynamic smem throttles CTAs/SM)
run() { echo "== $*"; env "$@" python bench.py --quick --steps 10 --warmup 3 --features ${FEAT:-1000} --windows 9472 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f' % d['value'], {k: round(v, 4) for k, v in d['kernels_ms'].items()})"; }
run X=0
# accum: 5.8 KB/CTA, 8 CTAs/SM (regs).  +40000 -> 4 CTAs/SM, +25000 -> 6, +70000 -> 2
run ISV_EXP_ACC_SMEM=25000
run ISV_EXP_ACC_SMEM=40000
run ISV_EXP_ACC_SMEM=70000
# backward: 37.8 KB/CTA of 4 warps, 4 CTAs/SM.  +20000 -> 3 CTAs, +40000 -> 2 CTAs, +100000 -> 1
run ISV_EXP_BWD_SMEM=20000
run ISV_EXP_BWD_SMEM=40000
run ISV_EXP_BWD_SMEM=100000
# tail: 28.9 KB/CTA, 4 CTAs/SM (regs).  +30000 -> 3, +50000 -> 2, +100000 -> 1
run ISV_EXP_TAIL_SMEM=30000
run ISV_EXP_TAIL_SMEM=50000
run ISV_EXP_TAIL_SMEM=100000
