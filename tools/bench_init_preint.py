#!/usr/bin/env python
"""tools/bench_init_preint.py -- device time of the two one-off kernels of the path that bench.py does not time on their own:
init_sparsify_kernel (the initFactorGraph tail, src/estimator.cpp:745-1001: one warp per 120 x 120 initial window) and
preintegrate_kernel (IntegrationBase::push_back over one frame interval).  Inputs resident; prints one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def timed(fn, sync, reps=20):
    import torch
    for _ in range(3):
        fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / reps


def main():
    import torch
    from is_vins_b200 import MargBackend, capi
    from oracle import sim
    torch.cuda.set_device(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    be = MargBackend(0)
    be.use_torch_stream()
    res = {}
    # ---- init sparsification: 3 distinct initial windows (V = 8) tiled
    chains = [sim.make_chain(sim.seed_for(6, b), L=8, rounds=1) for b in range(3)]
    poses = np.array([c.init_in.poses for c in chains])
    sbs = np.array([c.init_in.sbs for c in chains])
    pre = np.array([[p.pack() for p in c.init_in.pres] for c in chains])
    V = poses.shape[1]
    for n in (148, 2368):
        reps = (n + 2) // 3
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1))[:n])).cuda()
        dp, ds, dr = t(poses), t(sbs), t(pre)
        o_rel = torch.zeros((n, V - 1, capi.REL_REC), dtype=torch.float64, device="cuda")
        o_se3 = torch.zeros((n, capi.SE3_REC), dtype=torch.float64, device="cuda")
        o_vb = torch.zeros((n, capi.VB_REC), dtype=torch.float64, device="cuda")
        o_rank = torch.zeros((n,), dtype=torch.int32, device="cuda")
        o_st = torch.zeros((n,), dtype=torch.int32, device="cuda")
        ii = capi.isv_init_in(n, dp.data_ptr(), ds.data_ptr(), dr.data_ptr())
        oo = capi.isv_init_out(o_rel.data_ptr(), o_se3.data_ptr(), o_vb.data_ptr(), o_rank.data_ptr(), o_st.data_ptr())
        ms = timed(lambda: capi.check(be.lib.isv_init_sparsify_batch(be.h, C.byref(ii), C.byref(oo)), "init"), torch.cuda.synchronize)
        assert int(o_rank[0]) == 42 and not bool(o_st.any())
        res[f"init_sparsify_{n}_windows_ms"] = ms
        res[f"init_sparsify_{n}_windows_per_s"] = n / (ms * 1e-3)
    # ---- pre-integration: the bench's 9472 intervals of 10 samples
    b = bench.make_batch(150, 9472, 3)
    raw, init = torch.from_numpy(b.imu_raw).cuda(), torch.from_numpy(b.imu_init).cuda()
    out = torch.zeros((9472, capi.PREINT_REC), dtype=torch.float64, device="cuda")
    pi = capi.isv_preint_in(9472, int(raw.shape[1]), None, raw.data_ptr(), init.data_ptr())
    ms = timed(lambda: capi.check(be.lib.isv_preintegrate_batch(be.h, C.byref(pi), C.c_void_p(out.data_ptr())), "preint"), torch.cuda.synchronize)
    res["preintegrate_9472_intervals_ms"] = ms
    res["preintegrate_intervals_per_s"] = 9472 / (ms * 1e-3)
    res["preintegrate_samples_per_interval"] = int(raw.shape[1])
    print(json.dumps(res))
    be.close()


if __name__ == "__main__":
    main()
