timeout 600 python -m pytest tests/test_parity_sweep_gpu.py -m gpu -x -q -k "raw_imu" 2>&1 | tail -3
