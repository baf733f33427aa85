#!/bin/bash
# persistent landmark kernel (ISV_ACC_PERSIST warps per SM) next to the backward kernel: whole-step time
run() { echo "== $*"; env "$@" python bench.py --quick --steps 10 --warmup 3 --features ${FEAT:-1000} --windows 9472 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f' % d['value'], round(d['ms_per_step'],4), {k: round(v, 4) for k, v in d['kernels_ms'].items()})"; }
for FEAT in 1000 150; do
  export FEAT
  run ISV_ACC_PERSIST=0 ISV_NO_CARVEOUT=1
  for p in 0 3 4 5 6; do run ISV_ACC_PERSIST=$p; done
done
