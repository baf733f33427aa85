for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
timeout 300 python tools/sanitize_driver.py 2>&1 | tail -1
for i in 1 2 3 4 5; do timeout 120 python -m pytest tests/test_parity_sweep_gpu.py -m gpu -x -q -k "event_call or fused" 2>&1 | tail -1; done
