timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for L in 1000 150; do
  python tools/e2e_chunks.py $L 9472
  ISV_HOST_NO_ZC_OUT=1 python tools/e2e_chunks.py $L 9472
  python tools/e2e_chunks.py $L 9472
  ISV_HOST_NO_ZC_OUT=1 python tools/e2e_chunks.py $L 9472
done
