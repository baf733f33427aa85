for L in 1000 150; do
  python tools/e2e_chunks.py $L 9472
  ISV_HOST_EVEN_CHUNKS=1 python tools/e2e_chunks.py $L 9472
  python tools/e2e_chunks.py $L 9472
  ISV_HOST_EVEN_CHUNKS=1 python tools/e2e_chunks.py $L 9472
done
timeout 300 python -m pytest tests/test_parity_sweep_gpu.py tests/test_window_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
