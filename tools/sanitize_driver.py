#!/usr/bin/env python
"""tools/sanitize_driver.py -- a small pass over every kernel of the window path for compute-sanitizer (memcheck / racecheck /
synccheck): batch route, fused route (device batch and the staged zero-copy event), raw-IMU + FP32-xy inputs, host path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from is_vins_b200 import DeviceBatch, MargBackend, capi
    from is_vins_b200.backend import xy_as_f32
    be = MargBackend(0)
    counts = np.array([0, 1, 5, 31, 32, 33, 64, 129, 200, 7, 300, 150])
    b = bench.make_batch(100, len(counts), 77, ragged=1.0, counts=counts)
    db = DeviceBatch(b, "cuda:0")
    for fused in (0, 148):
        be.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, fused)
        be.marg_window_batch(db, capi.RUN_BOTH)
        be.synchronize()
    dr = DeviceBatch(b, "cuda:0", raw_imu=True, z_one=True, xy_f32=True)
    be.marg_window_batch(dr, capi.RUN_BOTH)
    be.synchronize()
    be.marg_window_batch_host(b, capi.RUN_BOTH, raw_imu=True, z_one=True, xy_f32=xy_as_f32(b.lm_obs))
    for w in (0, 2, 8):
        o = b.slice(w, w + 1)
        ob = o.lm_obs
        a1 = (o.pose_fwd[0, 0], o.pose_fwd[0, 1], o.ex_pose, ob[5], np.ascontiguousarray(ob[0:3].T),
              np.ascontiguousarray(np.vstack([ob[3:5], np.ones((1, ob.shape[1]))]).T), o.prior_se3[0], o.prior_rel[0], o.prior_rp[0])
        a2 = (o.pose_bwd[0, 0], o.sb_bwd[0, 0], o.pose_bwd[0, 1], o.sb_bwd[0, 1], o.prior_vb[0], o.preint[0])
        for mode in (0, 1, 2):
            be.set_tuning(capi.TUNE_EVENT_MODE, mode)
            be.marg_event(a1, a2)
    be.close()
    print("sanitize driver done")


if __name__ == "__main__":
    main()
