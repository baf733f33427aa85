timeout 300 python -m pytest tests/test_window_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for L in 1000 150; do
  timeout 120 python tools/latency_breakdown.py $L 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L', d['L'], 'staged:', json.dumps(d['isv_marg_event']))"
  ISV_EVENT_NO_STAGE=1 timeout 120 python tools/latency_breakdown.py $L 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('L', d['L'], 'NOT staged:', json.dumps(d['isv_marg_event']))"
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
