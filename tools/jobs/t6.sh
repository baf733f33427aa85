timeout 300 python -m pytest tests/test_window_parity_gpu.py -m gpu -x -q -k "packed_host" 2>&1 | tail -4
