#!/usr/bin/env python
"""tools/fused_stamps.py [L] -- where one MARGIN_OLD event spends its time inside marg_event_fused_kernel: the SM cycle
counter at the phase boundaries of the eight warps (isv_test_fused_stamps), in microseconds from kernel entry."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from is_vins_b200 import DeviceBatch, MargBackend, capi
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    be = MargBackend(0)
    b = bench.make_batch(L, 1, 5, ragged=0.0)
    db = DeviceBatch(b, "cuda:0")
    st = torch.zeros(80, dtype=torch.int64, device="cuda:0")
    bi, bo = db.structs()
    for _ in range(5):
        capi.check(be.lib.isv_test_fused_stamps(be.h, C.byref(bi), C.byref(bo), C.c_void_p(st.data_ptr())), "stamps")
        be.synchronize()
    mhz = 1965.0
    raw = st.cpu().numpy().astype(np.float64)
    s = raw[:64].reshape(8, 8)
    t0 = s[:, 0].min()
    us = np.where(s > 0, (s - t0) / mhz, np.nan)
    names = ["entry", "start barrier", "task done / backward done (w0)", "landmark share done", "fwd barrier passed (w7)",
             "tail done (w7)", "-", "final barrier"]
    bw = [round(float(x - t0) / mhz, 2) for x in raw[64:70]]
    print(json.dumps({"L": L, "backward_warp_us": dict(zip(["cholesky", "jacobians arrived", "householder", "lq", "backsolve", "cov blocks"], bw)), "sm_mhz_assumed": mhz, "columns": names, "us_from_entry": [[None if np.isnan(x) else round(float(x), 2) for x in r] for r in us]}))
    be.close()


if __name__ == "__main__":
    main()
