"""SoA window batches: the memory layout the C ABI consumes (include/isv_capi.h, isv_batch_in/out).

Plain NumPy on the host; `to_device` moves a batch to torch CUDA tensors (PyTorch is only used for
device memory and streams).  Matrices inside records are column-major, i.e. the bytes of the
Eigen member they mirror.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi


def se3_record(t, R, sqrt_info) -> np.ndarray:
    """SE3PriorFactor / RelativePoseFactor members -> [t3 | R col-major 9 | sqrt_info col-major 36]."""
    return np.concatenate([np.asarray(t, float).ravel(), np.asarray(R, float).flatten(order="F"),
                           np.asarray(sqrt_info, float).flatten(order="F")])


def vb_record(vb, sqrt_info) -> np.ndarray:
    return np.concatenate([np.asarray(vb, float).ravel(), np.asarray(sqrt_info, float).flatten(order="F")])


def rp_in_record(valid: bool, sqrt_info=None) -> np.ndarray:
    r = np.zeros(capi.RP_IN_REC)
    if valid:
        r[0] = 1.0
        r[1:5] = np.asarray(sqrt_info, float).flatten(order="F")
    return r


# ---- ABI 4: records without their structural zeros (include/isv_capi.h ISV_IN_TRI_RECORDS / ISV_OUT_TRI_RECORDS) ----------
# A record = segments that are copied (kind 0) or N x N column-major blocks whose upper triangle travels column by column.
TRI_LAYOUTS = {"se3": [(0, 12), (1, 6)], "rel": [(0, 12), (1, 6)], "vb": [(0, 9), (1, 9)], "rp_in": [(0, 1), (1, 2)],
               "pg": [(0, 12), (1, 6), (1, 6), (0, 5)], "rp": [(0, 9), (1, 2)]}


def _tri_index(layout):
    """(full index of every packed element, full length): the gather map full -> packed."""
    idx, fo = [], 0
    for kind, N in layout:
        if kind == 0:
            idx += list(range(fo, fo + N))
            fo += N
        else:
            idx += [fo + i + N * j for j in range(N) for i in range(j + 1)]
            fo += N * N
    return np.asarray(idx, np.int64), fo


def pack_tri(records: np.ndarray, family: str) -> np.ndarray:
    """[n, full] records -> [n, packed] (upper triangles only); raises if a strict lower triangle of a sqrt_info is not zero."""
    idx, fl = _tri_index(TRI_LAYOUTS[family])
    records = np.asarray(records, float)
    assert records.shape[1] == fl
    mask = np.ones(fl, bool)
    mask[idx] = False
    if family != "pg" and np.any(records[:, mask] != 0.0):
        raise ValueError(f"pack_tri({family}): a strict lower triangle is not zero")
    return np.ascontiguousarray(records[:, idx])


def unpack_tri(packed: np.ndarray, family: str, symmetric_blocks: Sequence[int] = ()) -> np.ndarray:
    """[n, packed] -> [n, full] with zero strict lower triangles; blocks listed in symmetric_blocks (segment numbers) are
    mirrored instead (covRel of the pose-graph record: segment 2)."""
    layout = TRI_LAYOUTS[family]
    idx, fl = _tri_index(layout)
    packed = np.asarray(packed, float)
    full = np.zeros((packed.shape[0], fl))
    full[:, idx] = packed
    fo = 0
    for sno, (kind, N) in enumerate(layout):
        if kind == 1 and sno in symmetric_blocks:
            up = full[:, fo:fo + N * N].reshape(-1, N, N).copy()   # [n][j][i] (column-major): only i <= j is filled
            dg = up * np.eye(N)
            full[:, fo:fo + N * N] = (up + up.transpose(0, 2, 1) - dg).reshape(-1, N * N)
        fo += N * N if kind else N
    return full


@dataclass
class WindowBatch:
    """Inputs of n independent MARGIN_OLD events (MargForward + MargBackward)."""
    n: int
    lm_offset: np.ndarray            # int64 [n+1]
    lm_obs: np.ndarray               # f64 [6, n_lm]  x_i y_i z_i x_j y_j inv_dep
    pose_fwd: np.ndarray             # [n,2,7]
    ex_pose: np.ndarray              # [7] (shared) or [n,7]
    prior_se3: np.ndarray            # [n,48]
    prior_rel: np.ndarray            # [n,48]
    prior_rp: Optional[np.ndarray]   # [n,5] or None
    pose_bwd: np.ndarray             # [n,2,7]
    sb_bwd: np.ndarray               # [n,2,9]
    prior_vb: np.ndarray             # [n,90]
    preint: np.ndarray               # [n,467]
    imu_raw: Optional[np.ndarray] = None    # [n,K,7] dt,acc,gyr  (input of the pre-integration kernel)
    imu_init: Optional[np.ndarray] = None   # [n,12]  acc0, gyr0, lin_ba, lin_bg

    FIELDS = ("lm_offset", "lm_obs", "pose_fwd", "ex_pose", "prior_se3", "prior_rel", "prior_rp", "pose_bwd",
              "sb_bwd", "prior_vb", "preint", "imu_raw", "imu_init")

    @property
    def n_landmarks(self) -> int:
        return int(self.lm_offset[-1])

    def input_bytes(self) -> int:
        return sum(getattr(self, f).nbytes for f in self.FIELDS[:11] if getattr(self, f) is not None)

    def tile(self, reps: int) -> "WindowBatch":
        """Repeat the batch `reps` times (independent copies of the same windows)."""
        n = self.n * reps
        counts = np.tile(np.diff(self.lm_offset), reps)
        off = np.zeros(n + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        t = lambda a: None if a is None else np.ascontiguousarray(np.tile(a, (reps,) + (1,) * (a.ndim - 1)))
        return WindowBatch(n, off, np.ascontiguousarray(np.tile(self.lm_obs, (1, reps))), t(self.pose_fwd),
                           self.ex_pose if self.ex_pose.ndim == 1 else t(self.ex_pose), t(self.prior_se3),
                           t(self.prior_rel), t(self.prior_rp), t(self.pose_bwd), t(self.sb_bwd), t(self.prior_vb),
                           t(self.preint), t(self.imu_raw), t(self.imu_init))

    def slice(self, lo: int, hi: int) -> "WindowBatch":
        """Windows [lo, hi) -- the contiguous shard one rank owns."""
        a, b = int(self.lm_offset[lo]), int(self.lm_offset[hi])
        s = lambda x: None if x is None else np.ascontiguousarray(x[lo:hi])
        return WindowBatch(hi - lo, (self.lm_offset[lo:hi + 1] - a).astype(np.int64),
                           np.ascontiguousarray(self.lm_obs[:, a:b]), s(self.pose_fwd),
                           self.ex_pose if self.ex_pose.ndim == 1 else s(self.ex_pose), s(self.prior_se3),
                           s(self.prior_rel), s(self.prior_rp), s(self.pose_bwd), s(self.sb_bwd), s(self.prior_vb),
                           s(self.preint), s(self.imu_raw), s(self.imu_init))


    def take(self, idx: Sequence[int]) -> "WindowBatch":
        """The windows `idx` (any order, repeats allowed) as a new batch -- e.g. the sample a checker re-computes."""
        idx = np.asarray(idx, dtype=np.int64)
        counts = np.diff(self.lm_offset)[idx]
        off = np.zeros(len(idx) + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        cols = np.concatenate([np.arange(self.lm_offset[w], self.lm_offset[w + 1]) for w in idx]) if len(idx) else np.zeros(0, np.int64)
        s = lambda x: None if x is None else np.ascontiguousarray(x[idx])
        return WindowBatch(len(idx), off, np.ascontiguousarray(self.lm_obs[:, cols.astype(np.int64)]), s(self.pose_fwd),
                           self.ex_pose if self.ex_pose.ndim == 1 else s(self.ex_pose), s(self.prior_se3),
                           s(self.prior_rel), s(self.prior_rp), s(self.pose_bwd), s(self.sb_bwd), s(self.prior_vb),
                           s(self.preint), s(self.imu_raw), s(self.imu_init))


def pack_events(events: Sequence, with_rp: bool = True) -> WindowBatch:
    """Pack records exposing the reference's member names (``fwd_in``/``bwd_in`` with pose0, pose1,
    ex_pose, inv_dep, pts_i, pts_j, prior_t, prior_R, prior_sqrt_info, rel_dt, rel_dR, rel_sqrt_info,
    rp_valid, rp_sqrt_info / pose_i, sb_i, pose_j, sb_j, vb_prior, vb_sqrt_info, pre.pack())."""
    n = len(events)
    counts = [int(len(e.fwd_in.inv_dep)) for e in events]
    off = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    obs = np.zeros((6, int(off[-1])))
    pose_fwd = np.zeros((n, 2, 7))
    prior_se3 = np.zeros((n, capi.SE3_REC))
    prior_rel = np.zeros((n, capi.REL_REC))
    prior_rp = np.zeros((n, capi.RP_IN_REC))
    pose_bwd = np.zeros((n, 2, 7))
    sb_bwd = np.zeros((n, 2, 9))
    prior_vb = np.zeros((n, capi.VB_REC))
    preint = np.zeros((n, capi.PREINT_REC))
    K = max(int(e.raw_imu.shape[0]) for e in events) if hasattr(events[0], "raw_imu") else 0
    same_k = K > 0 and all(int(e.raw_imu.shape[0]) == K for e in events)
    imu_raw = np.zeros((n, K, 7)) if same_k else None
    imu_init = np.zeros((n, 12)) if same_k else None
    ex = None
    for w, e in enumerate(events):
        f, b = e.fwd_in, e.bwd_in
        a, c = int(off[w]), int(off[w + 1])
        obs[0:3, a:c] = np.asarray(f.pts_i, float).T
        obs[3:5, a:c] = np.asarray(f.pts_j, float).T[0:2]
        obs[5, a:c] = f.inv_dep
        pose_fwd[w, 0], pose_fwd[w, 1] = f.pose0, f.pose1
        ex = np.asarray(f.ex_pose, float) if ex is None else ex
        prior_se3[w] = se3_record(f.prior_t, f.prior_R, f.prior_sqrt_info)
        prior_rel[w] = se3_record(f.rel_dt, f.rel_dR, f.rel_sqrt_info)
        prior_rp[w] = rp_in_record(bool(f.rp_valid), f.rp_sqrt_info)
        pose_bwd[w, 0], pose_bwd[w, 1] = b.pose_i, b.pose_j
        sb_bwd[w, 0], sb_bwd[w, 1] = b.sb_i, b.sb_j
        prior_vb[w] = vb_record(b.vb_prior, b.vb_sqrt_info)
        preint[w] = b.pre.pack()
        if same_k:
            imu_raw[w] = e.raw_imu
            imu_init[w] = np.concatenate([e.acc0, e.gyr0, b.pre.linearized_ba, b.pre.linearized_bg])
    return WindowBatch(n, off, obs, pose_fwd, ex, prior_se3, prior_rel, prior_rp if with_rp else None, pose_bwd,
                       sb_bwd, prior_vb, preint, imu_raw, imu_init)


@dataclass
class WindowOutputs:
    se3: np.ndarray      # [n,48] forwardPosePriorEdgeToAdd
    pg: np.ndarray       # [n,89] CombinedFactors
    rel: np.ndarray      # [n,48] backwardRelativePoseEdgeToAdd
    vb: np.ndarray       # [n,90] backwardVBEdgeToAdd
    rp: np.ndarray       # [n,13] rollPitchFactor
    rank: np.ndarray     # [n,2]
    status: np.ndarray   # [n]

    # record views (column-major matrices -> numpy row-major arrays)
    def se3_sqrt_info(self, w):
        return self.se3[w, 12:48].reshape(6, 6).T

    def se3_t(self, w):
        return self.se3[w, 0:3]

    def se3_R(self, w):
        return self.se3[w, 3:12].reshape(3, 3).T

    def pg_dt(self, w):
        return self.pg[w, 0:3]

    def pg_dR(self, w):
        return self.pg[w, 3:12].reshape(3, 3).T

    def pg_sqrt_info(self, w):
        return self.pg[w, 12:48].reshape(6, 6).T

    def pg_covRel(self, w):
        return self.pg[w, 48:84].reshape(6, 6).T

    def pg_distance(self, w):
        return self.pg[w, 84]

    def pg_covAbs(self, w):
        return self.pg[w, 85:89].reshape(2, 2).T

    def rel_dt(self, w):
        return self.rel[w, 0:3]

    def rel_dR(self, w):
        return self.rel[w, 3:12].reshape(3, 3).T

    def rel_sqrt_info(self, w):
        return self.rel[w, 12:48].reshape(6, 6).T

    def vb_VB(self, w):
        return self.vb[w, 0:9]

    def vb_sqrt_info(self, w):
        return self.vb[w, 9:90].reshape(9, 9).T

    def rp_R(self, w):
        return self.rp[w, 0:9].reshape(3, 3).T

    def rp_sqrt_info(self, w):
        return self.rp[w, 9:13].reshape(2, 2).T


# sub-blocks of each output record, compared separately so that a large-magnitude block cannot mask a small one
OUTPUT_BLOCKS = {"se3": [(0, 3), (3, 12), (12, 48)], "pg": [(0, 3), (3, 12), (12, 48), (48, 84), (84, 85), (85, 89)],
                 "rel": [(0, 3), (3, 12), (12, 48)], "vb": [(0, 9), (9, 90)], "rp": [(0, 9), (9, 13)]}


def outputs_rel_diff(a: WindowOutputs, b: WindowOutputs, which: int = 3) -> np.ndarray:
    """Per window: the worst relative Frobenius difference over every block of every recovered-factor record
    (`which` & 1: MargForward's se3 / pg, & 2: MargBackward's rel / vb / rp)."""
    n = a.rank.shape[0]
    worst = np.zeros(n)
    fams = (["se3", "pg"] if which & 1 else []) + (["rel", "vb", "rp"] if which & 2 else [])
    for f in fams:
        x, y = getattr(a, f), getattr(b, f)
        for lo, hi in OUTPUT_BLOCKS[f]:
            d = np.linalg.norm(x[:, lo:hi] - y[:, lo:hi], axis=1)
            r = np.linalg.norm(y[:, lo:hi], axis=1)
            worst = np.maximum(worst, np.where(r > 0, d / np.where(r > 0, r, 1.0), d))
    return worst


def pack_tri_inputs(batch: "WindowBatch") -> Dict[str, Optional[np.ndarray]]:
    """The four prior-record arrays of a batch in the ABI 4 packed form (ISV_IN_TRI_RECORDS)."""
    return {"prior_se3": pack_tri(batch.prior_se3, "se3"), "prior_rel": pack_tri(batch.prior_rel, "rel"),
            "prior_rp": None if batch.prior_rp is None else pack_tri(batch.prior_rp, "rp_in"),
            "prior_vb": pack_tri(batch.prior_vb, "vb")}


def packed_outputs(n: int) -> "WindowOutputs":
    """Result buffers of the ABI 4 packed form (ISV_OUT_TRI_RECORDS): 191 instead of 289 doubles per window."""
    return WindowOutputs(np.zeros((n, capi.SE3_TRI_REC)), np.zeros((n, capi.PG_TRI_REC)), np.zeros((n, capi.REL_TRI_REC)),
                         np.zeros((n, capi.VB_TRI_REC)), np.zeros((n, capi.RP_TRI_REC)), np.zeros((n, 2), np.int32),
                         np.zeros((n,), np.int32))


def unpack_outputs(o: "WindowOutputs") -> "WindowOutputs":
    """Packed results -> the full records (strict lower triangles of the sqrt_info blocks zero, covRel mirrored)."""
    return WindowOutputs(unpack_tri(o.se3, "se3"), unpack_tri(o.pg, "pg", symmetric_blocks=(2,)), unpack_tri(o.rel, "rel"),
                         unpack_tri(o.vb, "vb"), unpack_tri(o.rp, "rp"), o.rank, o.status)
