python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 4 8; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 150 9472; done
for c in 4 8; do ISV_HOST_CHUNKS=$c python tools/e2e_chunks.py 1000 9472; done
ISV_HOST_TRACE=1 python tools/e2e_chunks.py 150 9472 2>&1 | grep -A5 "chunk 0" | head -6
python - <<'PY'
import torch, numpy as np, ctypes as C, sys
sys.path.insert(0, '.')
import bench
from is_vins_b200 import MargBackend, capi
be = MargBackend(0)
b = bench.make_batch(150, 9472, 3)
raw = torch.from_numpy(b.imu_raw).cuda(); init = torch.from_numpy(b.imu_init).cuda()
out = torch.zeros((9472, 467), dtype=torch.float64, device='cuda')
pi = capi.isv_preint_in(9472, int(raw.shape[1]), None, raw.data_ptr(), init.data_ptr())
for _ in range(3): capi.check(be.lib.isv_preintegrate_batch(be.h, C.byref(pi), C.c_void_p(out.data_ptr())), 'p')
be.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20): be.lib.isv_preintegrate_batch(be.h, C.byref(pi), C.c_void_p(out.data_ptr()))
be.synchronize()
print('preintegrate_kernel ms per 9472 windows (K=%d):' % raw.shape[1], (time.perf_counter() - t0) / 20 * 1e3)
PY
