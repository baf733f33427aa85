for L in 1000 150; do
  python tools/e2e_chunks.py $L 9472
  ISV_HOST_EVEN_CHUNKS=1 python tools/e2e_chunks.py $L 9472
  ISV_HOST_CHUNKS=8 python tools/e2e_chunks.py $L 9472
done
ISV_HOST_TRACE=1 python tools/e2e_chunks.py 1000 9472 2>&1 | grep -B0 -A5 "chunk 0" | sed -n 7,11p
timeout 300 python -m pytest tests/test_parity_sweep_gpu.py tests/test_window_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
