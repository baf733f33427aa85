python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 4 8; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 150 9472; done
for c in 2 4 8; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 1000 9472; done
ISV_HOST_TRACE=1 python tools/e2e_chunks.py 1000 9472 2>&1 | grep -A5 "chunk 0" | head -6
ISV_HOST_TRACE=1 python tools/e2e_chunks.py 150 9472 2>&1 | grep -A5 "chunk 0" | head -6
