#!/usr/bin/env python
"""tools/latency_breakdown.py [L] [n] -- one MARGIN_OLD event (or a small batch) alone on the GPU.

  resident: wall time (host call -> results ready, stream synchronised) and device time of isv_marg_window_batch with the
            inputs resident, on the one-launch fused kernel and on each kernel group of the batch route;
  event:    isv_marg_event through its three routes (ISV_TUNE_EVENT_MODE 0 / 1 / 2), timed INSIDE the library
            (isv_test_event_latency: what a C++ estimator pays), and once more through the Python binding.
Median of 300 calls; prints one JSON line."""
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
    for name, which, fused in (("fused_kernel", capi.RUN_BOTH, 1 << 20), ("both", capi.RUN_BOTH, 0), ("forward", capi.RUN_FORWARD, 0),
                               ("backward", capi.RUN_BACKWARD, 0), ("factor_jac(2 launches)", 16, 0), ("accum", 4, 0),
                               ("tail", 8, 0), ("backward_kernel", 32, 0)):
        be.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, fused)
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
    # an empty measurement: what event record + launch + synchronise cost with nothing to do
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts, dev = [], []
    for _ in range(300):
        t0 = time.perf_counter()
        e0.record()
        e1.record()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        dev.append(e0.elapsed_time(e1) * 1e3)
    res["empty"] = {"wall_us": float(np.median(ts) * 1e6), "device_us": float(np.median(dev))}
    if n == 1:
        be.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 148)
        o = b.lm_obs
        args1 = (b.pose_fwd[0, 0], b.pose_fwd[0, 1], b.ex_pose if b.ex_pose.ndim == 1 else b.ex_pose[0], o[5],
                 np.ascontiguousarray(o[0:3].T), np.ascontiguousarray(np.vstack([o[3:5], np.ones((1, o.shape[1]))]).T),
                 b.prior_se3[0], b.prior_rel[0], None if b.prior_rp is None else b.prior_rp[0])
        args2 = (b.pose_bwd[0, 0], b.sb_bwd[0, 0], b.pose_bwd[0, 1], b.sb_bwd[0, 1], b.prior_vb[0], b.preint[0])
        ev = {}
        for mode, label in ((0, "zero_copy_fused"), (1, "staged_fused"), (2, "staged_batch_kernels")):
            be.set_tuning(capi.TUNE_EVENT_MODE, mode)
            be.event_latency_us(args1, args2, 50)
            us = be.event_latency_us(args1, args2, 300)
            ev[label] = {"median_us": float(np.median(us)), "p10_us": float(np.percentile(us, 10)),
                         "p90_us": float(np.percentile(us, 90)), "p99_us": float(np.percentile(us, 99))}
        be.set_tuning(capi.TUNE_EVENT_MODE, 0)
        ts = []
        for it in range(320):
            t0 = time.perf_counter()
            be.marg_event(args1, args2)
            if it >= 20:
                ts.append(time.perf_counter() - t0)
        ev["zero_copy_fused_through_python_binding_us"] = float(np.median(ts) * 1e6)
        res["isv_marg_event"] = ev
    print(json.dumps(res))
    be.close()


if __name__ == "__main__":
    main()
