for c in 1 2 4; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 150 4096; done
for c in 2 4; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 150 2048; done
for c in 1 2; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 150 512; done
