#!/usr/bin/env python
"""tools/latency_breakdown.py [L] -- one MARGIN_OLD event alone on the GPU: wall time (host call -> results ready) of each
kernel group of isv_marg_window_batch, inputs resident.  Median of 300 calls; prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from is_vins_b200 import DeviceBatch, MargBackend, capi
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    torch.cuda.set_device(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    be = MargBackend(0)
    be.use_torch_stream()
    b = bench.make_batch(L, n, 5, ragged=0.0)
    db = DeviceBatch(b, "cuda:0")
    res = {"L": L, "windows": n}
    for name, which in (("both", capi.RUN_BOTH), ("forward", capi.RUN_FORWARD), ("backward", capi.RUN_BACKWARD),
                        ("factor_jac(2 launches)", 16), ("accum", 4), ("tail", 8), ("backward_kernel", 32)):
        for _ in range(30):
            be.marg_window_batch(db, which)
        torch.cuda.synchronize()
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev = []
        for _ in range(300):
            t0 = time.perf_counter()
            e0.record()
            be.marg_window_batch(db, which)
            e1.record()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            dev.append(e0.elapsed_time(e1) * 1e3)
        res[name] = {"wall_us": float(np.median(ts) * 1e6), "device_us": float(np.median(dev))}
    print(json.dumps(res))
    be.close()


if __name__ == "__main__":
    main()
