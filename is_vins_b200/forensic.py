"""Forensic mode: the reference's DENSE route and its KLD diagnostics on the GPU (include/isv_capi.h, "forensic mode").

Not the hot path.  `forensic_batch` runs the product kernels and also returns their intermediates (Lamda_prior of
MargForward, the factor G of MargBackward) and the reference's KLD values; `literal_forward` rebuilds one window's dense
`Lamda` (/root/reference/src/estimator.cpp:1164-1242) through a DIFFERENT set of kernels -- the ceres-`Evaluate` kernels
and `ne_build_kernel`, the generic engine's `ThreadsConstructA` analogue -- and eliminates the whole (L + 6)^2 block the way
the reference does (`isv_literal_schur`: full-pivot inverse, :1286-1288).  structured == literal localises a parity break
without any CPU code.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from . import capi
from .backend import DeviceBatch, MargBackend
from .batch import WindowBatch
from .marginalization import MarginalizationInfo, ResidualBlockInfo


def forensic_batch(be: MargBackend, db: DeviceBatch) -> Dict[str, np.ndarray]:
    """isv_marg_forensic_batch on a device-resident batch; db.out receives the normal outputs."""
    import torch
    n = db.n
    kw = dict(dtype=torch.float64, device=db.device)
    t = {"lamda_prior_fwd": torch.zeros((n, 36), **kw), "g_bwd": torch.zeros((n, 315), **kw),
         "kld_fwd": torch.zeros((n,), **kw), "kld_bwd": torch.zeros((n,), **kw),
         "lamda_prior_bwd": torch.zeros((n, 441), **kw), "eig_bwd": torch.zeros((n, 21), **kw),
         "info_abs": torch.zeros((n, 9), **kw), "info_yaw": torch.zeros((n,), **kw)}
    fo = capi.isv_forensic_out(*[t[k].data_ptr() for k in ("lamda_prior_fwd", "g_bwd", "kld_fwd", "kld_bwd",
                                                           "lamda_prior_bwd", "eig_bwd", "info_abs", "info_yaw")])
    bi, bo = db.structs()
    capi.check(be.lib.isv_marg_forensic_batch(be.h, C.byref(bi), C.byref(bo), C.byref(fo)), "isv_marg_forensic_batch")
    be.synchronize()
    r = {k: v.cpu().numpy() for k, v in t.items()}
    r["lamda_prior_fwd"] = r["lamda_prior_fwd"].reshape(n, 6, 6).transpose(0, 2, 1)      # column-major records
    r["lamda_prior_bwd"] = r["lamda_prior_bwd"].reshape(n, 21, 21).transpose(0, 2, 1)
    r["g_bwd"] = r["g_bwd"].reshape(n, 15, 21)
    r["info_abs"] = r["info_abs"].reshape(n, 3, 3).transpose(0, 2, 1)
    return r


def literal_forward(be: MargBackend, batch: WindowBatch, w: int) -> Dict[str, np.ndarray]:
    """Window w of `batch`: the dense Lamda ((12 + L)^2, OrderMap order T1@0, T0@6, landmark k@12+k, :1153-1162) assembled
    on the GPU from every factor's ceres-Evaluate Jacobians, and Lamda_prior by the full-pivot dense route."""
    import torch
    a, b = int(batch.lm_offset[w]), int(batch.lm_offset[w + 1])
    L = b - a
    mi = MarginalizationInfo(be, eps=0.0, cauchy_a=0.0, constant=[("ex_pose", 0)])     # no robust loss in Marg* (Q2)
    for k in range(L):
        o = batch.lm_obs[:, a + k]
        mi.addResidualBlockInfo(ResidualBlockInfo("projection", [("pose", 0), ("pose", 1), ("ex_pose", 0), ("feature", k)],
                                                  drop_set=[0, 3], pts_i=o[0:3], pts_j=np.array([o[3], o[4], 1.0])))
    ps, pr = batch.prior_se3[w], batch.prior_rel[w]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", [("pose", 0)], drop_set=[0], t=ps[0:3], R=ps[3:12].reshape(3, 3).T,
                                              sqrt_info=ps[12:48].reshape(6, 6).T))
    mi.addResidualBlockInfo(ResidualBlockInfo("rel", [("pose", 0), ("pose", 1)], drop_set=[0], delta_t=pr[0:3],
                                              delta_R=pr[3:12].reshape(3, 3).T, sqrt_info=pr[12:48].reshape(6, 6).T))
    ex = batch.ex_pose if batch.ex_pose.ndim == 1 else batch.ex_pose[w]
    mi.preMarginalize({"pose": batch.pose_fwd[w], "speed_bias": np.zeros((2, 9)), "ex_pose": ex.reshape(1, 7),
                       "feature": batch.lm_obs[5, a:b].copy()})
    mi.marginalize(build_only=True)
    idx = mi.parameter_block_idx
    perm = [idx[("pose", 1)] + i for i in range(6)] + [idx[("pose", 0)] + i for i in range(6)] + \
           [idx[("feature", k)] for k in range(L)]
    p = torch.tensor(perm, dtype=torch.int64, device=mi.A_dev.device)
    lam = mi.A_dev.index_select(0, p).index_select(1, p).contiguous()                  # re-ordering only: no arithmetic
    n = 12 + L
    prior = torch.zeros((36,), dtype=torch.float64, device=lam.device)
    rank = torch.zeros((1,), dtype=torch.int32, device=lam.device)
    capi.check(be.lib.isv_literal_schur(be.h, 1, n, 6, C.c_void_p(lam.data_ptr()), C.c_void_p(prior.data_ptr()), None,
                                        C.c_void_p(rank.data_ptr())), "isv_literal_schur")
    be.synchronize()
    return {"Lamda": lam.cpu().numpy(), "Lamda_prior": prior.cpu().numpy().reshape(6, 6).T, "rank_mm": int(rank.item()),
            "status": mi.status}
