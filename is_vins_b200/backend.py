"""`MargBackend`: one handle of libisv_b200.so bound to one GPU (include/isv_capi.h).

Mirrors the call sites of the reference's backend (`Estimator::backendOptimization`,
/root/reference/src/estimator.cpp:1541-1562): `marg_forward` / `marg_backward` for one
MARGIN_OLD event, and the batched form over independent windows.  All arithmetic happens in the
CUDA library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import capi
from .batch import WindowBatch, WindowOutputs


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(capi.c_double_p)


def xy_as_f32(lm_obs: np.ndarray) -> np.ndarray:
    """Components 0, 1 of lm_obs (pts_i.x, pts_i.y) narrowed to FP32 [2, n_lm]; they are cv::Point2f values in the reference
    (feature_tracker_simple.h:55, System.cpp:119-122), so the narrowing must be exact -- anything else is a caller error."""
    xy = np.ascontiguousarray(lm_obs[0:2].astype(np.float32))
    if not np.array_equal(xy.astype(np.float64), lm_obs[0:2]):
        raise ValueError("xy_f32: pts_i.x / pts_i.y are not FP32-representable; pass them as doubles")
    return xy


class DeviceBatch:
    """A WindowBatch resident in HBM (torch CUDA tensors) plus preallocated outputs."""

    def __init__(self, batch: WindowBatch, device, pinned_src: bool = False, raw_imu: bool = False, z_one: bool = False,
                 xy_f32: bool = False):
        """raw_imu: hand the library the raw IMU samples (imu_raw / imu_init) instead of the pre-integration record (ABI
        2: it runs preintegrate_kernel first); z_one: promise pts_i.z == 1 (ISV_IN_PTS_I_Z_ONE); xy_f32: hand pts_i.x / pts_i.y
        over as FP32 (ABI 3, isv_batch_in::lm_xy_f32 -- they ARE floats in the reference; raises if narrowing is lossy)."""
        import torch
        self.xy_f32 = None
        if xy_f32:
            self.xy_f32 = torch.from_numpy(xy_as_f32(batch.lm_obs)).to(device)
        self.raw_imu, self.flags = bool(raw_imu), (capi.IN_PTS_I_Z_ONE if z_one else 0)
        if raw_imu and (batch.imu_raw is None or batch.imu_init is None):
            raise ValueError("raw_imu=True needs batch.imu_raw / batch.imu_init")
        self.torch = torch
        self.device = torch.device(device)
        self.n = batch.n
        self.t: Dict[str, "torch.Tensor"] = {}
        for f in WindowBatch.FIELDS:
            a = getattr(batch, f)
            if a is not None:
                self.t[f] = torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.ex_shared = batch.ex_pose.ndim == 1
        self.n_lm = batch.n_landmarks
        n = batch.n
        kw = dict(dtype=torch.float64, device=self.device)
        self.out = {
            "se3": torch.empty((n, capi.SE3_REC), **kw), "pg": torch.empty((n, capi.PG_REC), **kw),
            "rel": torch.empty((n, capi.REL_REC), **kw), "vb": torch.empty((n, capi.VB_REC), **kw),
            "rp": torch.empty((n, capi.RP_REC), **kw),
            "rank": torch.zeros((n, 2), dtype=torch.int32, device=self.device),
            "status": torch.zeros((n,), dtype=torch.int32, device=self.device),
        }

    def structs(self):
        t = self.t
        g = lambda k: t[k].data_ptr() if k in t else None
        bi = capi.isv_batch_in(self.n, 1 if self.ex_shared else 0, g("lm_offset"), g("lm_obs"), self.n_lm,
                               g("pose_fwd"), g("ex_pose"), g("prior_se3"), g("prior_rel"), g("prior_rp"),
                               g("pose_bwd"), g("sb_bwd"), g("prior_vb"), None if self.raw_imu else g("preint"))
        if self.raw_imu:
            bi.imu_raw, bi.imu_init, bi.imu_k_max = g("imu_raw"), g("imu_init"), int(t["imu_raw"].shape[1])
        bi.flags = self.flags
        if self.xy_f32 is not None:
            bi.lm_xy_f32 = self.xy_f32.data_ptr()
        o = self.out
        bo = capi.isv_batch_out(o["se3"].data_ptr(), o["pg"].data_ptr(), o["rel"].data_ptr(), o["vb"].data_ptr(),
                                o["rp"].data_ptr(), o["rank"].data_ptr(), o["status"].data_ptr())
        return bi, bo

    def outputs(self) -> WindowOutputs:
        o = {k: v.cpu().numpy() for k, v in self.out.items()}
        return WindowOutputs(o["se3"], o["pg"], o["rel"], o["vb"], o["rp"], o["rank"], o["status"])


class MargBackend:
    def __init__(self, device: int = 0, config: Optional[capi.isv_config] = None):
        self.lib = capi.load()
        self.cfg = config if config is not None else capi.default_config()
        self.h = C.c_void_p()
        capi.check(self.lib.isv_create(C.byref(self.cfg), int(device), C.byref(self.h)), "isv_create")
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.isv_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ---------------------------------------------------------------------------
    def use_torch_stream(self):
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        if not s:
            # isv_set_stream(h, NULL) means "back to the handle's own non-blocking stream": torch's legacy default stream
            # (handle 0) can therefore not be shared, and silently staying unordered with torch would let
            # DeviceBatch.outputs() read results before the kernels finish
            raise capi.IsvError("use_torch_stream(): torch's current stream is the default stream (handle 0), which the "
                                "library cannot share; run under `torch.cuda.stream(torch.cuda.Stream())` or call "
                                "MargBackend.synchronize() before reading results")
        capi.check(self.lib.isv_set_stream(self.h, C.c_void_p(s)), "isv_set_stream")

    def synchronize(self):
        capi.check(self.lib.isv_synchronize(self.h), "isv_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.isv_launch_count(self.h))

    def set_tuning(self, knob: int, value: int) -> None:
        """isv_set_tuning: capi.TUNE_FUSED_MAX_WINDOWS (0 = always the batch kernels) / capi.TUNE_EVENT_MODE (0, 1, 2)."""
        capi.check(self.lib.isv_set_tuning(self.h, int(knob), int(value)), "isv_set_tuning")

    # ---- batched, device-resident ---------------------------------------------------------------
    def marg_window_batch(self, dbatch: DeviceBatch, which: int = capi.RUN_BOTH) -> None:
        """Stream-ordered MargForward+MargBackward over dbatch; results land in dbatch.out."""
        bi, bo = dbatch.structs()
        capi.check(self.lib.isv_marg_window_batch(self.h, C.byref(bi), C.byref(bo), which), "isv_marg_window_batch")

    # ---- batched, host pointers (H2D + kernels + D2H inside the call) ---------------------------
    def marg_window_batch_host(self, batch: WindowBatch, which: int = capi.RUN_BOTH,
                               out: Optional[WindowOutputs] = None, raw_imu: bool = False,
                               z_one: bool = False, xy_f32: Optional[np.ndarray] = None,
                               tri_in: Optional[Dict[str, Optional[np.ndarray]]] = None, tri_out: bool = False) -> WindowOutputs:
        """raw_imu / z_one: see DeviceBatch (fewer bytes cross PCIe: 12 + 7 K doubles instead of the 467-double
        pre-integration record, 3 instead of 4 doubles per landmark)."""
        n = batch.n
        if out is None and tri_out:
            from .batch import packed_outputs
            out = packed_outputs(n)
        if out is None:
            out = WindowOutputs(np.zeros((n, capi.SE3_REC)), np.zeros((n, capi.PG_REC)), np.zeros((n, capi.REL_REC)),
                                np.zeros((n, capi.VB_REC)), np.zeros((n, capi.RP_REC)),
                                np.zeros((n, 2), np.int32), np.zeros((n,), np.int32))
        bi = capi.isv_batch_in(n, 1 if batch.ex_pose.ndim == 1 else 0, _p(batch.lm_offset), _p(batch.lm_obs),
                               batch.lm_obs.shape[1], _p(batch.pose_fwd), _p(batch.ex_pose), _p(batch.prior_se3),
                               _p(batch.prior_rel), _p(batch.prior_rp), _p(batch.pose_bwd), _p(batch.sb_bwd),
                               _p(batch.prior_vb), None if raw_imu else _p(batch.preint))
        if raw_imu:
            bi.imu_raw, bi.imu_init, bi.imu_k_max = _p(batch.imu_raw), _p(batch.imu_init), int(batch.imu_raw.shape[1])
        bi.flags = (capi.IN_PTS_I_Z_ONE if z_one else 0) | (capi.OUT_TRI_RECORDS if tri_out else 0)
        if tri_in is not None:   # ABI 4: batch.pack_tri_inputs(batch) -- the prior records without their structural zeros
            bi.flags |= capi.IN_TRI_RECORDS
            bi.prior_se3, bi.prior_rel, bi.prior_vb = _p(tri_in["prior_se3"]), _p(tri_in["prior_rel"]), _p(tri_in["prior_vb"])
            bi.prior_rp = _p(tri_in["prior_rp"])
        if xy_f32 is not None:   # ABI 3: [2, n_lm] float32 from xy_as_f32(batch.lm_obs) (pin it for the e2e measurement)
            assert xy_f32.dtype == np.float32 and xy_f32.shape == (2, batch.lm_obs.shape[1]) and xy_f32.flags.c_contiguous
            bi.lm_xy_f32 = _p(xy_f32)
        bo = capi.isv_batch_out(_p(out.se3), _p(out.pg), _p(out.rel), _p(out.vb), _p(out.rp), _p(out.rank),
                                _p(out.status))
        capi.check(self.lib.isv_marg_window_batch_host(self.h, C.byref(bi), C.byref(bo), which),
                   "isv_marg_window_batch_host")
        return out

    # ---- one MARGIN_OLD event (what Estimator::MargForward / MargBackward call) -------------------
    def marg_forward(self, pose0, pose1, ex_pose, inv_dep, pts_i, pts_j, prior_se3, prior_rel, prior_rp=None):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pose0, pose1, ex_pose, inv_dep, pts_i, pts_j,
                                                                 prior_se3, prior_rel)]
        rp = None if prior_rp is None else np.ascontiguousarray(prior_rp, dtype=np.float64)
        fi = capi.isv_fwd_in(int(a[3].shape[0]), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), _dp(a[5]),
                             _dp(a[6]), _dp(a[7]), None if rp is None else _dp(rp))
        fo = capi.isv_fwd_out()
        capi.check(self.lib.isv_marg_forward(self.h, C.byref(fi), C.byref(fo)), "isv_marg_forward")
        return np.array(fo.se3), np.array(fo.pg), int(fo.rank), int(fo.status)

    def marg_backward(self, pose_i, sb_i, pose_j, sb_j, prior_vb, preint):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (pose_i, sb_i, pose_j, sb_j, prior_vb, preint)]
        bi = capi.isv_bwd_in(*[_dp(x) for x in a])
        bo = capi.isv_bwd_out()
        capi.check(self.lib.isv_marg_backward(self.h, C.byref(bi), C.byref(bo)), "isv_marg_backward")
        return np.array(bo.rel), np.array(bo.vb), np.array(bo.rp), int(bo.rank), int(bo.status)

    def marg_event(self, fwd_args, bwd_args):
        """`MargForward(); MargBackward();` of one MARGIN_OLD event in one blocking call (isv_marg_event): fwd_args /
        bwd_args are the argument tuples of marg_forward / marg_backward; returns their two result tuples."""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in fwd_args[:8]]
        rp = None if len(fwd_args) < 9 or fwd_args[8] is None else np.ascontiguousarray(fwd_args[8], dtype=np.float64)
        fi = capi.isv_fwd_in(int(a[3].shape[0]), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), _dp(a[5]),
                             _dp(a[6]), _dp(a[7]), None if rp is None else _dp(rp))
        b = [np.ascontiguousarray(x, dtype=np.float64) for x in bwd_args]
        bi = capi.isv_bwd_in(*[_dp(x) for x in b])
        fo, bo = capi.isv_fwd_out(), capi.isv_bwd_out()
        capi.check(self.lib.isv_marg_event(self.h, C.byref(fi), C.byref(bi), C.byref(fo), C.byref(bo)), "isv_marg_event")
        return ((np.array(fo.se3), np.array(fo.pg), int(fo.rank), int(fo.status)),
                (np.array(bo.rel), np.array(bo.vb), np.array(bo.rp), int(bo.rank), int(bo.status)))

    def event_latency_us(self, fwd_args, bwd_args, iters: int = 300) -> np.ndarray:
        """isv_test_event_latency: `iters` isv_marg_event calls of one event timed inside the library (no ctypes / NumPy
        marshalling in the timed region) -> microseconds per call."""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in fwd_args[:8]]
        rp = None if len(fwd_args) < 9 or fwd_args[8] is None else np.ascontiguousarray(fwd_args[8], dtype=np.float64)
        fi = capi.isv_fwd_in(int(a[3].shape[0]), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), _dp(a[5]),
                             _dp(a[6]), _dp(a[7]), None if rp is None else _dp(rp))
        b = [np.ascontiguousarray(x, dtype=np.float64) for x in bwd_args]
        bi = capi.isv_bwd_in(*[_dp(x) for x in b])
        fo, bo = capi.isv_fwd_out(), capi.isv_bwd_out()
        us = np.zeros(iters)
        capi.check(self.lib.isv_test_event_latency(self.h, C.byref(fi), C.byref(bi), C.byref(fo), C.byref(bo), iters, _dp(us)),
                   "isv_test_event_latency")
        return us

    # ---- initFactorGraph sparsification tail (one-time, src/estimator.cpp:745-1001) ---------------
    def init_sparsify(self, poses, sbs, preint):
        """poses [n,V,7], sbs [n,V,9], preint [n,V-1,467] (host) -> dict of recovered-factor records."""
        V = int(self.cfg.vo_size)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, V, 7)
        n = poses.shape[0]
        sbs = np.ascontiguousarray(sbs, dtype=np.float64).reshape(n, V, 9)
        preint = np.ascontiguousarray(preint, dtype=np.float64).reshape(n, V - 1, capi.PREINT_REC)
        out = {"rel": np.zeros((n, V - 1, capi.REL_REC)), "se3": np.zeros((n, capi.SE3_REC)),
               "vb": np.zeros((n, capi.VB_REC)), "rank": np.zeros((n,), np.int32), "status": np.zeros((n,), np.int32)}
        ii = capi.isv_init_in(n, _p(poses), _p(sbs), _p(preint))
        oo = capi.isv_init_out(_p(out["rel"]), _p(out["se3"]), _p(out["vb"]), _p(out["rank"]), _p(out["status"]))
        capi.check(self.lib.isv_init_sparsify_host(self.h, C.byref(ii), C.byref(oo)), "isv_init_sparsify_host")
        return out

    # ---- IMU pre-integration (include/factor/integration_base.h:30-158) --------------------------
    def preintegrate(self, imu_raw, imu_init, k_count=None):
        """imu_raw [n,K,7] (dt, acc, gyr), imu_init [n,12] (acc_0, gyr_0, lin_ba, lin_bg) -> [n,467]."""
        imu_raw = np.ascontiguousarray(imu_raw, dtype=np.float64)
        n, K = imu_raw.shape[0], imu_raw.shape[1]
        imu_init = np.ascontiguousarray(imu_init, dtype=np.float64).reshape(n, 12)
        kc = None if k_count is None else np.ascontiguousarray(k_count, dtype=np.int32)
        out = np.zeros((n, capi.PREINT_REC))
        pi = capi.isv_preint_in(n, K, _p(kc), _p(imu_raw), _p(imu_init))
        capi.check(self.lib.isv_preintegrate_host(self.h, C.byref(pi), _p(out)), "isv_preintegrate_host")
        return out
