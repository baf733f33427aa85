#!/usr/bin/env python
"""tools/profile_driver.py -- one launch of every kernel of the window path, for `ncu --set full` (profiles/r02x_*):
the batch kernels on the bench's headline batch (9472 windows, L ~ 1000, inputs resident), preintegrate_kernel (raw-IMU
inputs), the fused single-event kernel on 148 windows and on one staged event (isv_marg_event)."""
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
    b = bench.make_batch(L, 9472, 1000)
    db = DeviceBatch(b, "cuda:0")
    for _ in range(2):
        be.marg_window_batch(db, capi.RUN_BOTH)
    be.synchronize()
    dr = DeviceBatch(b, "cuda:0", raw_imu=True, z_one=True, xy_f32=True)
    be.marg_window_batch(dr, capi.RUN_BOTH)
    be.synchronize()
    s = b.slice(0, 148)
    ds = DeviceBatch(s, "cuda:0")
    for _ in range(2):
        be.marg_window_batch(ds, capi.RUN_BOTH)
    be.synchronize()
    o = b.slice(0, 1)
    ob = o.lm_obs
    args1 = (o.pose_fwd[0, 0], o.pose_fwd[0, 1], o.ex_pose, ob[5], np.ascontiguousarray(ob[0:3].T),
             np.ascontiguousarray(np.vstack([ob[3:5], np.ones((1, ob.shape[1]))]).T), o.prior_se3[0], o.prior_rel[0], o.prior_rp[0])
    args2 = (o.pose_bwd[0, 0], o.sb_bwd[0, 0], o.pose_bwd[0, 1], o.sb_bwd[0, 1], o.prior_vb[0], o.preint[0])
    for _ in range(2):
        be.marg_event(args1, args2)
    be.close()
    print("profile driver done")


if __name__ == "__main__":
    main()
