"""Host mirror of the batched ceres-`Evaluate` entry points (include/isv_capi.h `isv_eval_*`).

`FactorProblem` holds, as plain arrays, the factor list `Estimator::problemSolve()` hands to ceres
(/root/reference/src/estimator.cpp:1004-1146): parameter blocks + per-type factor records.
`DeviceProblem` puts it in HBM (torch tensors) with preallocated outputs; `MargBackend.eval_problem`
launches the three kernels (projection / IMU / recovered-prior factors).  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import capi


def rel_record(dt, dR, sqrt_info) -> np.ndarray:
    """[delta_t 3 | delta_R 9 col-major | sqrt_info 36 col-major]  (also the SE3 prior record)."""
    return np.concatenate([np.asarray(dt, float).ravel(), np.asarray(dR, float).T.ravel(),
                           np.asarray(sqrt_info, float).T.ravel()])


def vb_record(vb, sqrt_info) -> np.ndarray:
    return np.concatenate([np.asarray(vb, float).ravel(), np.asarray(sqrt_info, float).T.ravel()])


def rp_record(R, sqrt_info) -> np.ndarray:
    return np.concatenate([np.asarray(R, float).T.ravel(), np.asarray(sqrt_info, float).T.ravel()])


def yaw_record(yaw_meas, sqrt_info) -> np.ndarray:
    return np.concatenate([np.asarray(yaw_meas, float).ravel(), np.asarray(sqrt_info, float).ravel()])


@dataclass
class FactorProblem:
    pose: np.ndarray                       # [n_pose,7]
    speed_bias: np.ndarray                 # [n_sb,9]
    ex_pose: np.ndarray                    # [n_ex,7]
    feature: np.ndarray                    # [n_feat]
    proj_idx: np.ndarray                   # int32 [4,P]
    proj_obs: np.ndarray                   # [5,P]
    imu_idx: np.ndarray                    # int32 [n_imu,2]
    imu_preint: np.ndarray                 # [n_imu,467]
    rel_idx: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    rel_rec: np.ndarray = field(default_factory=lambda: np.zeros((0, capi.REL_REC)))
    se3_idx: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    se3_rec: np.ndarray = field(default_factory=lambda: np.zeros((0, capi.SE3_REC)))
    vb_idx: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    vb_rec: np.ndarray = field(default_factory=lambda: np.zeros((0, capi.VB_REC)))
    rp_idx: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    rp_rec: np.ndarray = field(default_factory=lambda: np.zeros((0, capi.RP_REC)))
    yaw_idx: np.ndarray = field(default_factory=lambda: np.zeros((0,), np.int32))
    yaw_rec: np.ndarray = field(default_factory=lambda: np.zeros((0, capi.YAW_REC)))

    @classmethod
    def from_factors(cls, p) -> "FactorProblem":
        """From an object exposing the reference's member names (poses/sbs/ex/feat arrays, proj_idx,
        proj_obs, imu_idx, imu_pre[k].pack(), and factor lists rel/se3/vb/rp/yaw with delta_t, delta_R,
        t, R, VB, yaw_meas, sqrt_info, imu_i/imu_j/index)."""
        i32 = lambda a, shape: np.asarray(a, np.int32).reshape(shape)
        f64 = lambda rows, w: np.asarray(rows, float).reshape(-1, w)
        return cls(
            np.ascontiguousarray(p.poses, float), np.ascontiguousarray(p.sbs, float),
            np.ascontiguousarray(p.ex, float).reshape(-1, 7), np.ascontiguousarray(p.feat, float),
            np.ascontiguousarray(p.proj_idx, np.int32), np.ascontiguousarray(p.proj_obs, float),
            np.ascontiguousarray(p.imu_idx, np.int32), f64([q.pack() for q in p.imu_pre], capi.PREINT_REC),
            i32([(f.imu_i, f.imu_j) for f in p.rel], (-1, 2)),
            f64([rel_record(f.delta_t, f.delta_R, f.sqrt_info) for f in p.rel], capi.REL_REC),
            i32([f.index for f in p.se3], (-1,)), f64([rel_record(f.t, f.R, f.sqrt_info) for f in p.se3], capi.SE3_REC),
            i32([f.index for f in p.vb], (-1,)), f64([vb_record(f.VB, f.sqrt_info) for f in p.vb], capi.VB_REC),
            i32([f.index for f in p.rp], (-1,)), f64([rp_record(f.R, f.sqrt_info) for f in p.rp], capi.RP_REC),
            i32([f.index for f in p.yaw], (-1,)), f64([yaw_record(f.yaw_meas, f.sqrt_info) for f in p.yaw], capi.YAW_REC))

    @classmethod
    def load(cls, path: str) -> "FactorProblem":
        """From an .npz holding the fields of this class (tests/golden/problem_*.npz)."""
        z = np.load(path)
        return cls(*[z[k] for k in ("pose", "speed_bias", "ex_pose", "feature", "proj_idx", "proj_obs", "imu_idx",
                                    "imu_preint", "rel_idx", "rel_rec", "se3_idx", "se3_rec", "vb_idx", "vb_rec",
                                    "rp_idx", "rp_rec", "yaw_idx", "yaw_rec")])

    def tile(self, reps: int) -> "FactorProblem":
        """`reps` independent copies (parameter blocks concatenated, indices offset)."""
        n_pose, n_sb, n_ex, n_feat = len(self.pose), len(self.speed_bias), len(self.ex_pose), len(self.feature)
        t = lambda a: np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))
        P = self.proj_idx.shape[1]
        pidx = np.tile(self.proj_idx, (1, reps)).astype(np.int64)
        r = np.repeat(np.arange(reps), P)
        pidx[0] += r * n_pose
        pidx[1] += r * n_pose
        pidx[2] += r * n_ex
        pidx[3] += r * n_feat

        def off(idx, step):
            k = idx.shape[0]
            o = np.repeat(np.arange(reps), k) * step
            tt = np.tile(idx, (reps,) + (1,) * (idx.ndim - 1)).astype(np.int64)
            return (tt + (o[:, None] if idx.ndim == 2 else o)).astype(np.int32)
        return FactorProblem(t(self.pose), t(self.speed_bias), t(self.ex_pose), t(self.feature),
                             np.ascontiguousarray(pidx.astype(np.int32)),
                             np.ascontiguousarray(np.tile(self.proj_obs, (1, reps))),
                             off(self.imu_idx, n_pose), t(self.imu_preint), off(self.rel_idx, n_pose), t(self.rel_rec),
                             off(self.se3_idx, n_pose), t(self.se3_rec), off(self.vb_idx, n_sb), t(self.vb_rec),
                             off(self.rp_idx, n_pose), t(self.rp_rec), off(self.yaw_idx, n_pose), t(self.yaw_rec))


class DeviceProblem:
    """A FactorProblem resident in HBM plus preallocated Evaluate outputs."""
    IN = ("pose", "speed_bias", "ex_pose", "feature", "proj_idx", "proj_obs", "imu_idx", "imu_preint", "rel_idx",
          "rel_rec", "se3_idx", "se3_rec", "vb_idx", "vb_rec", "rp_idx", "rp_rec", "yaw_idx", "yaw_rec")

    def __init__(self, p: FactorProblem, device, want_ex_jac: bool = True, want_jac: bool = True):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.p = p
        self.t: Dict[str, "torch.Tensor"] = {k: torch.from_numpy(np.ascontiguousarray(getattr(p, k))).to(self.device)
                                             for k in self.IN}
        P, ni = p.proj_idx.shape[1], p.imu_idx.shape[0]
        z = lambda *s: torch.zeros(s, dtype=torch.float64, device=self.device)
        self.n_proj, self.n_imu = P, ni
        self.out = {"proj_res": z(P, 2), "imu_res": z(ni, 15), "rel_res": z(len(p.rel_idx), 6),
                    "se3_res": z(len(p.se3_idx), 6), "vb_res": z(len(p.vb_idx), 9), "rp_res": z(len(p.rp_idx), 2),
                    "yaw_res": z(len(p.yaw_idx), 1)}
        if want_jac:
            self.out.update({"proj_ji": z(P, 14), "proj_jj": z(P, 14), "proj_jf": z(P, 2),
                             "imu_jac": z(ni, capi.IMU_JAC_REC), "rel_jac": z(len(p.rel_idx), 84),
                             "se3_jac": z(len(p.se3_idx), 42), "vb_jac": z(len(p.vb_idx), 81),
                             "rp_jac": z(len(p.rp_idx), 14), "yaw_jac": z(len(p.yaw_idx), 7)})
            if want_ex_jac:
                self.out["proj_je"] = z(P, 14)
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.device)

    def _o(self, k) -> Optional[int]:
        return self.out[k].data_ptr() if k in self.out else None

    def param_blocks(self) -> capi.isv_param_blocks:
        t = self.t
        return capi.isv_param_blocks(t["pose"].shape[0], t["speed_bias"].shape[0], t["ex_pose"].shape[0],
                                     t["feature"].shape[0], t["pose"].data_ptr(), t["speed_bias"].data_ptr(),
                                     t["ex_pose"].data_ptr(), t["feature"].data_ptr())

    def host(self) -> Dict[str, np.ndarray]:
        return {k: v.cpu().numpy() for k, v in self.out.items()}


def eval_problem(backend, dp: DeviceProblem, cauchy_a: float = 0.0, fused_call: bool = True) -> None:
    """Stream-ordered: the three Evaluate kernels over dp; results land in dp.out, flags in dp.status.
    fused_call: one `isv_eval_problem` (IMU / prior kernels overlapped on side streams) instead of the
    three per-class entry points back to back."""
    lib, h, t = backend.lib, backend.h, dp.t
    pb = dp.param_blocks()
    st = C.c_void_p(dp.status.data_ptr())
    pf = capi.isv_proj_factors(dp.n_proj, dp.n_proj, t["proj_idx"].data_ptr(), t["proj_obs"].data_ptr(), cauchy_a)
    po = capi.isv_proj_eval(dp._o("proj_res"), dp._o("proj_ji"), dp._o("proj_jj"), dp._o("proj_je"), dp._o("proj_jf"))
    mf = capi.isv_imu_factors(dp.n_imu, t["imu_idx"].data_ptr(), t["imu_preint"].data_ptr())
    mo = capi.isv_imu_eval(dp._o("imu_res"), dp._o("imu_jac"))
    n = [t[k].shape[0] for k in ("rel_idx", "se3_idx", "vb_idx", "rp_idx", "yaw_idx")]
    sf = capi.isv_small_factors(*n, t["rel_idx"].data_ptr(), t["rel_rec"].data_ptr(), t["se3_idx"].data_ptr(),
                                t["se3_rec"].data_ptr(), t["vb_idx"].data_ptr(), t["vb_rec"].data_ptr(),
                                t["rp_idx"].data_ptr(), t["rp_rec"].data_ptr(), t["yaw_idx"].data_ptr(),
                                t["yaw_rec"].data_ptr(), cauchy_a)
    so = capi.isv_small_eval(dp._o("rel_res"), dp._o("rel_jac"), dp._o("se3_res"), dp._o("se3_jac"),
                             dp._o("vb_res"), dp._o("vb_jac"), dp._o("rp_res"), dp._o("rp_jac"),
                             dp._o("yaw_res"), dp._o("yaw_jac"))
    if fused_call:
        capi.check(lib.isv_eval_problem(h, C.byref(pb), C.byref(pf) if dp.n_proj else None, C.byref(po),
                                        C.byref(mf) if dp.n_imu else None, C.byref(mo),
                                        C.byref(sf) if sum(n) else None, C.byref(so), st), "isv_eval_problem")
        return
    if dp.n_proj:
        capi.check(lib.isv_eval_projection_batch(h, C.byref(pb), C.byref(pf), C.byref(po), st), "isv_eval_projection_batch")
    if dp.n_imu:
        capi.check(lib.isv_eval_imu_batch(h, C.byref(pb), C.byref(mf), C.byref(mo), st), "isv_eval_imu_batch")
    if sum(n):
        capi.check(lib.isv_eval_small_batch(h, C.byref(pb), C.byref(sf), C.byref(so), st), "isv_eval_small_batch")
