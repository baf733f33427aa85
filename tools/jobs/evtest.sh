timeout 600 python -m pytest tests/test_parity_sweep_gpu.py -m gpu -x -q -k "event_call or fused" 2>&1 | tail -5
