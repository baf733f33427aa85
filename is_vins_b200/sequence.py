"""`SequenceState`: the estimator's frame-to-frame factor members for n independent sequences, resident
in HBM (include/isv_capi.h `isv_seq_*`).  One method per reference step of
`Estimator::backendOptimization` / `slideWindow` (/root/reference/src/estimator.cpp:1541-1562,
:1133-1144, :520-550, :1605-1638) and of the pose-graph accumulator
(/root/reference/src/pose_graph/pose_graph_builder.cpp:157-214).  Host arrays passed to the methods are
moved to the device with torch (plumbing); all arithmetic runs in libisv_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import capi
from .batch import WindowBatch


class SequenceState:
    def __init__(self, backend, n_sequences: int):
        import torch
        self.torch = torch
        self.be = backend
        self.n = int(n_sequences)
        self.V = int(backend.cfg.vo_size)
        self.dev = torch.device("cuda", backend.device)
        self.s = C.c_void_p()
        capi.check(backend.lib.isv_seq_create(backend.h, self.n, C.byref(self.s)), "isv_seq_create")
        self._keep = []

    def close(self):
        if self.s.value:
            self.be.lib.isv_seq_destroy(self.be.h, self.s)
            self.s = C.c_void_p()

    def _d(self, a, dtype=np.float64):
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.dev)
        self._keep.append(t)      # keep alive until the next synchronising call
        return t.data_ptr()

    # ---- initFactorGraph tail -------------------------------------------------------------------
    def init(self, poses, sbs, preint) -> np.ndarray:
        """poses [n,V,7], sbs [n,V,9], preint [n,V-1,467] -> rank [n] of the init marginal."""
        rank = self.torch.zeros((self.n,), dtype=self.torch.int32, device=self.dev)
        ii = capi.isv_init_in(self.n, self._d(poses), self._d(sbs), self._d(preint))
        capi.check(self.be.lib.isv_seq_init(self.be.h, self.s, C.byref(ii), rank.data_ptr()), "isv_seq_init")
        self.be.synchronize()
        self._keep.clear()
        return rank.cpu().numpy()

    # ---- factor->update(...) ----------------------------------------------------------------------
    def update(self, old_P, old_R, old_vb, pose, speed_bias) -> None:
        """old_P [n,V,3], old_R [n,V,3,3] (row-major numpy matrices; transposed to the column-major ABI
        here), old_vb [n,9]; pose [n,V,7], speed_bias [n,9] after the solve."""
        oR = np.ascontiguousarray(np.swapaxes(np.asarray(old_R, float).reshape(self.n, self.V, 3, 3), -1, -2))
        ui = capi.isv_seq_update_in(self._d(old_P), self._d(oR), self._d(old_vb), self._d(pose), self._d(speed_bias))
        capi.check(self.be.lib.isv_seq_update(self.be.h, self.s, C.byref(ui)), "isv_seq_update")

    # ---- double2vector's rotation of the priors ----------------------------------------------------
    def yaw(self, old_R0, pose0) -> np.ndarray:
        """old_R0 [n,3,3], pose0 [n,7] -> rot_diff [n,3,3] (also applied to the two priors on the device)."""
        oR = np.ascontiguousarray(np.swapaxes(np.asarray(old_R0, float).reshape(self.n, 3, 3), -1, -2))
        rot = self.torch.zeros((self.n, 9), dtype=self.torch.float64, device=self.dev)
        capi.check(self.be.lib.isv_seq_yaw(self.be.h, self.s, self._d(oR), self._d(pose0), rot.data_ptr()), "isv_seq_yaw")
        self.be.synchronize()
        self._keep.clear()
        return np.swapaxes(rot.cpu().numpy().reshape(self.n, 3, 3), -1, -2).copy()

    # ---- MARGIN_OLD + pose-graph accumulation + slideWindow rotation ---------------------------------
    def marginalize(self, batch: WindowBatch, ts=None, Ri=None, ti=None, pg_cut_distance: float = 0.1):
        """batch: the per-frame inputs (its prior_* members are ignored: priors come from the state).
        Returns (kf_flag [n], kf_records [n,ACC_REC]) when pose-graph members are given, else None."""
        pg = ts is not None
        fr = capi.isv_seq_frame(1 if batch.ex_pose.ndim == 1 else 0, self._d(batch.lm_offset, np.int64),
                                self._d(batch.lm_obs), batch.lm_obs.shape[1], self._d(batch.pose_fwd),
                                self._d(batch.ex_pose), self._d(batch.pose_bwd), self._d(batch.sb_bwd),
                                self._d(batch.preint), None, None, None)
        kf = flag = None
        if pg:
            RiT = np.ascontiguousarray(np.swapaxes(np.asarray(Ri, float).reshape(self.n, 3, 3), -1, -2))
            fr.ts, fr.Ri, fr.ti = self._d(ts), self._d(RiT), self._d(ti)
            kf = self.torch.zeros((self.n, capi.ACC_REC), dtype=self.torch.float64, device=self.dev)
            flag = self.torch.zeros((self.n,), dtype=self.torch.int32, device=self.dev)
        capi.check(self.be.lib.isv_seq_marginalize(self.be.h, self.s, C.byref(fr), float(pg_cut_distance),
                                                   kf.data_ptr() if pg else None, flag.data_ptr() if pg else None),
                   "isv_seq_marginalize")
        self.be.synchronize()
        self._keep.clear()
        return (flag.cpu().numpy(), kf.cpu().numpy()) if pg else None

    # ---- checkpoint / resume ---------------------------------------------------------------------------
    def _host_struct(self, h: Dict[str, np.ndarray]) -> capi.isv_seq_host:
        p = lambda k: h[k].ctypes.data_as(C.c_void_p) if k in h else None
        return capi.isv_seq_host(*[p(k) for k, _ in capi.isv_seq_host._fields_])

    def export(self) -> Dict[str, np.ndarray]:
        n, V = self.n, self.V
        h = {"rel": np.zeros((V, n, capi.REL_REC)), "se3": np.zeros((n, capi.SE3_REC)), "vb": np.zeros((n, capi.VB_REC)),
             "rp": np.zeros((V, n, capi.RP_REC)), "rp_valid": np.zeros((V, n), np.int32),
             "acc": np.zeros((n, capi.ACC_REC)), "pg_count": np.zeros((n,), np.int32),
             "last_se3": np.zeros((n, capi.SE3_REC)), "last_pg": np.zeros((n, capi.PG_REC)),
             "last_rel": np.zeros((n, capi.REL_REC)), "last_vb": np.zeros((n, capi.VB_REC)),
             "last_rp": np.zeros((n, capi.RP_REC)), "last_rank": np.zeros((n, 2), np.int32),
             "last_status": np.zeros((n,), np.int32)}
        hs = self._host_struct(h)
        capi.check(self.be.lib.isv_seq_export_host(self.be.h, self.s, C.byref(hs)), "isv_seq_export_host")
        return h

    def restore(self, h: Dict[str, np.ndarray]) -> None:
        h = {k: np.ascontiguousarray(v) for k, v in h.items()}
        hs = self._host_struct(h)
        capi.check(self.be.lib.isv_seq_import_host(self.be.h, self.s, C.byref(hs)), "isv_seq_import_host")
