"""ctypes wrapper of oracle/libisv_ref.so (plain-C restatement, oracle/isv_ref.c).

TEST INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libisv_ref.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.isv_ref_marg_window_batch.restype = C.c_int
        _lib.isv_ref_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(load().isv_ref_max_threads())


def marg_window_batch(batch, which: int = 3, n_threads: int = 0, structured: bool = False, cfg=None):
    """Run the C restatement on a is_vins_b200.batch.WindowBatch (host arrays).  Returns WindowOutputs."""
    from is_vins_b200 import capi
    from is_vins_b200.batch import WindowOutputs
    lib = load()
    n = batch.n
    if cfg is None:
        cfg = capi.isv_config()
        cfg.alpha = 0.1
        cfg.proj_sqrt_info[0] = cfg.proj_sqrt_info[3] = 460.0
        cfg.g[2] = 9.81007
        cfg.acc_n, cfg.gyr_n, cfg.acc_w, cfg.gyr_w = 0.22627, 0.003988, 0.001, 0.0001
        cfg.vo_size, cfg.all_buf_size, cfg.qr_rank_eps_log10 = 8, 18, -16
    out = WindowOutputs(np.zeros((n, capi.SE3_REC)), np.zeros((n, capi.PG_REC)), np.zeros((n, capi.REL_REC)),
                        np.zeros((n, capi.VB_REC)), np.zeros((n, capi.RP_REC)), np.zeros((n, 2), np.int32),
                        np.zeros((n,), np.int32))
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    bi = capi.isv_batch_in(n, 1 if batch.ex_pose.ndim == 1 else 0, p(batch.lm_offset), p(batch.lm_obs),
                           batch.lm_obs.shape[1], p(batch.pose_fwd), p(batch.ex_pose), p(batch.prior_se3),
                           p(batch.prior_rel), p(batch.prior_rp), p(batch.pose_bwd), p(batch.sb_bwd),
                           p(batch.prior_vb), p(batch.preint))
    bo = capi.isv_batch_out(p(out.se3), p(out.pg), p(out.rel), p(out.vb), p(out.rp), p(out.rank), p(out.status))
    rc = lib.isv_ref_marg_window_batch(C.byref(cfg), C.byref(bi), C.byref(bo), int(which), int(n_threads),
                                       1 if structured else 0)
    assert rc == 0
    return out
