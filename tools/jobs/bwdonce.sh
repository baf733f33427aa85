timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for L in 1000 150; do python tools/e2e_chunks.py $L 9472; python tools/e2e_chunks.py $L 9472; done
ISV_HOST_CHUNKS=8 python tools/e2e_chunks.py 150 9472
ISV_HOST_CHUNKS=2 python tools/e2e_chunks.py 150 9472
python tools/e2e_chunks.py 150 4096
python tools/e2e_chunks.py 2000 4736
