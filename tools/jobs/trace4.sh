ISV_HOST_TRACE=1 python tools/e2e_chunks.py 150 9472 2>&1 | grep -B0 -A5 "chunk 0" | sed -n 7,12p
