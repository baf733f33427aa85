run() { echo "== $*"; env "$@" python bench.py --quick --steps 10 --warmup 3 --features ${FEAT:-1000} --windows 9472 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f' % d['value'], round(d['ms_per_step'],4), {k: round(v, 4) for k, v in d['kernels_ms'].items()})"; }
for FEAT in 1000 150; do export FEAT; run X=1; run X=2; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
