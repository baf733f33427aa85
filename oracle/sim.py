"""Synthetic EuRoC-shaped sliding windows, driven through the oracle's own estimator chain.

TEST INFRASTRUCTURE (see oracle/isv_oracle.py header).  Produces, for a seed, a sequence of
``MARGIN_OLD`` events: the exact inputs `Estimator::MargForward` / `MargBackward` /
`initFactorGraph` (sparsification tail) read, with prior factors that are *realistic* because they
come from running the oracle's own `init_sparsify` followed by forward/backward rounds with the
reference's factor rotation (src/estimator.cpp:1605-1638) and pseudo-measurement re-centring
(`update`, src/estimator.cpp:1133-1144) in between.

Constants: config/euroc_config.yaml (camera :5-17, extrinsics :24-35, noise :57-61).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import isv_oracle as O

FX, FY, CX, CY = 461.6, 460.3, 363.0, 248.1
IMG_W, IMG_H = 752, 480
RIC = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422],
                [0.999557249008, 0.0149672133247, 0.025715529948],
                [-0.0257744366974, 0.00375618835797, 0.999660727178]])
TIC = np.array([-0.0216401454975, -0.064676986768, 0.00981073058949])


def seed_for(config_id: int, batch_id: int) -> int:
    """seed = 20260000 + 1000*config_id + batch_id."""
    return 20260000 + 1000 * config_id + batch_id


def ex_pose() -> np.ndarray:
    # orthonormalise the yaml matrix the way Quaterniond{ric} + normalisation would
    q = O.q_normalized(O.R_to_q(RIC))
    return np.array([TIC[0], TIC[1], TIC[2], q[1], q[2], q[3], q[0]])


def pose_vec(p, q) -> np.ndarray:
    return np.array([p[0], p[1], p[2], q[1], q[2], q[3], q[0]])


class Trajectory:
    """Smooth MH_01-like motion: speed ~0.3-1.2 m/s, yaw-rate ~0.3 rad/s."""

    def __init__(self, rng: np.random.Generator):
        self.A = rng.uniform(0.5, 2.0, 3) * np.array([1.0, 1.0, 0.3])
        self.w = rng.uniform(0.3, 0.7, 3)
        self.ph = rng.uniform(0, 2 * math.pi, 3)
        self.yaw0 = rng.uniform(-math.pi, math.pi)
        self.yaw_rate = rng.normal(0, 0.3)
        self.rp_amp = rng.uniform(0.02, 0.15, 2)
        self.rp_w = rng.uniform(0.5, 1.5, 2)
        self.rp_ph = rng.uniform(0, 2 * math.pi, 2)

    def p(self, t):
        return self.A * np.sin(self.w * t + self.ph)

    def v(self, t):
        return self.A * self.w * np.cos(self.w * t + self.ph)

    def a(self, t):
        return -self.A * self.w * self.w * np.sin(self.w * t + self.ph)

    def R(self, t):
        yaw = self.yaw0 + self.yaw_rate * t + 0.2 * math.sin(0.4 * t)
        pitch = self.rp_amp[0] * math.sin(self.rp_w[0] * t + self.rp_ph[0])
        roll = self.rp_amp[1] * math.sin(self.rp_w[1] * t + self.rp_ph[1])
        return O.ypr2R(np.array([yaw, pitch, roll]) * 180.0 / math.pi)

    def omega_body(self, t, h=1e-5):
        Rm, Rp = self.R(t - h), self.R(t + h)
        return O.SO3.from_matrix(Rm.T @ Rp).log() / (2 * h)


@dataclass
class Frame:
    t: float
    P: np.ndarray
    Q: np.ndarray        # (w,x,y,z)
    V: np.ndarray
    Ba: np.ndarray
    Bg: np.ndarray
    pre: Optional[O.IntegrationBase]   # pre-integration from the previous frame to this one
    raw: Optional[np.ndarray]          # (K,7) [dt, acc3, gyr3] feeding `pre`
    acc0: Optional[np.ndarray] = None
    gyr0: Optional[np.ndarray] = None

    def pose(self):
        return pose_vec(self.P, self.Q)

    def sb(self):
        return np.concatenate([self.V, self.Ba, self.Bg])


def make_frames(rng, cfg: O.Config, n_frames: int, K: int = 10, dt: float = 0.005,
                max_gap: int = 1) -> List[Frame]:
    traj = Trajectory(rng)
    ba_true = rng.normal(0, 0.02, 3)
    bg_true = rng.normal(0, 0.002, 3)
    G = cfg.G
    frames: List[Frame] = []
    t = rng.uniform(0, 20.0)

    def meas(tt, ba, bg):
        Rw = traj.R(tt)
        acc = Rw.T @ (traj.a(tt) + G) + ba + rng.normal(0, cfg.acc_n / math.sqrt(dt) * 0.05, 3)
        gyr = traj.omega_body(tt) + bg + rng.normal(0, cfg.gyr_n / math.sqrt(dt) * 0.05, 3)
        return acc, gyr

    acc_prev, gyr_prev = meas(t, ba_true, bg_true)
    for f in range(n_frames):
        pre = None
        raw = None
        a0 = g0 = None
        if f > 0:
            gap = int(rng.integers(1, max_gap + 1))
            steps = K * gap
            # the estimator's bias estimate at the start of the interval
            lin_ba = ba_true + rng.normal(0, 0.002, 3)
            lin_bg = bg_true + rng.normal(0, 0.0002, 3)
            a0, g0 = acc_prev.copy(), gyr_prev.copy()
            pre = O.IntegrationBase(a0, g0, lin_ba, lin_bg, cfg)
            raw = np.zeros((steps, 7))
            for s in range(steps):
                t += dt
                ba_true = ba_true + rng.normal(0, cfg.acc_w * math.sqrt(dt), 3)
                bg_true = bg_true + rng.normal(0, cfg.gyr_w * math.sqrt(dt), 3)
                acc, gyr = meas(t, ba_true, bg_true)
                pre.push_back(dt, acc, gyr)
                raw[s, 0] = dt
                raw[s, 1:4] = acc
                raw[s, 4:7] = gyr
                acc_prev, gyr_prev = acc, gyr
        # post-solve estimate = truth + small perturbation (1 cm, 0.2 deg)
        P = traj.p(t) + rng.normal(0, 0.01, 3)
        Rw = traj.R(t)
        q = O.q_normalized(O.q_mul(O.R_to_q(Rw), O.SO3.exp(rng.normal(0, 0.2 * math.pi / 180, 3)).q))
        Vv = traj.v(t) + rng.normal(0, 0.02, 3)
        Ba = ba_true + rng.normal(0, 0.002, 3)
        Bg = bg_true + rng.normal(0, 0.0002, 3)
        frames.append(Frame(t, P, q, Vv, Ba, Bg, pre, raw, a0, g0))
    return frames


def make_landmarks(rng, cfg: O.Config, f0: Frame, f1: Frame, L: int, expose=ex_pose()):
    """L features hosted in frame 0 and observed in frame 1: uniform in the frame-0 image, depth
    U(1,8) m, exact projection into frame 1 + N(0,(1/460)^2) on normalised coords, z == 1
    (src/System.cpp:346).  Returns (inv_dep[L], pts_i[L,3], pts_j[L,3])."""
    u = rng.uniform(0, IMG_W, L)
    v = rng.uniform(0, IMG_H, L)
    pts_i = np.stack([(u - CX) / FX, (v - CY) / FY, np.ones(L)], axis=1)
    depth = rng.uniform(1.0, 8.0, L)
    qic = O.quat_from_pose(expose)
    tic = expose[0:3]
    ric = O.q_to_R(qic)
    R0, R1 = O.q_to_R(f0.Q), O.q_to_R(f1.Q)
    pc = pts_i * depth[:, None]
    pw = (R0 @ (ric @ pc.T + tic[:, None])).T + f0.P
    pj = (ric.T @ (R1.T @ (pw - f1.P).T - tic[:, None])).T
    pts_j = np.stack([pj[:, 0] / pj[:, 2] + rng.normal(0, 1.0 / 460.0, L),
                      pj[:, 1] / pj[:, 2] + rng.normal(0, 1.0 / 460.0, L), np.ones(L)], axis=1)
    # the estimator's inverse depth: truth perturbed by ~1 %
    inv_dep = (1.0 / depth) * (1.0 + rng.normal(0, 0.01, L))
    return inv_dep, pts_i, pts_j


@dataclass
class MargEvent:
    """Everything one MARGIN_OLD event reads, plus the oracle's outputs."""
    fwd_in: O.ForwardInput
    bwd_in: O.BackwardInput
    raw_imu: np.ndarray          # (K,7) samples of the backward interval
    acc0: np.ndarray
    gyr0: np.ndarray
    fwd_out: Optional[O.ForwardOutput] = None
    bwd_out: Optional[O.BackwardOutput] = None
    # what the steps around Marg* read (device-resident sequence tests): the states before the solve
    # (Ps/Rs/Vs/Bas/Bgs) and after it (para_*), the yaw re-anchoring rotation, the pose-graph record
    upd: Optional[dict] = None
    rot_diff: Optional[np.ndarray] = None
    pg_meta: Optional[dict] = None
    state_after: Optional[dict] = None   # factor members after slideWindow's rotation


@dataclass
class Chain:
    cfg: O.Config
    init_in: O.InitInput
    init_out: O.InitOutput
    events: List[MargEvent]


def _perturb_and_update(rng, frames_win: List[Frame], prior: O.SE3PriorFactor, rels: List[O.RelativePoseFactor],
                        vbp: O.Linear9Factor, rps: List[O.RollPitchFactor], V: int, scale: float, rec: Optional[dict] = None):
    """Mimic problemSolve(): states move a little, then factor->update re-centres the
    pseudo-measurements  (src/estimator.cpp:1133-1144)."""
    old = [(f.P.copy(), O.q_to_R(f.Q), f.V.copy(), f.Ba.copy(), f.Bg.copy()) for f in frames_win]
    if rec is not None:
        rec["old_P"] = np.array([o[0] for o in old[:V]])
        rec["old_R"] = np.array([o[1] for o in old[:V]])
        rec["old_vb"] = np.concatenate([old[V - 1][2], old[V - 1][3], old[V - 1][4]])
    for f in frames_win:
        f.P = f.P + rng.normal(0, 0.003 * scale, 3)
        f.Q = O.q_normalized(O.q_mul(f.Q, O.SO3.exp(rng.normal(0, 0.0005 * scale, 3)).q))
        f.V = f.V + rng.normal(0, 0.005 * scale, 3)
        f.Ba = f.Ba + rng.normal(0, 0.0005 * scale, 3)
        f.Bg = f.Bg + rng.normal(0, 0.00005 * scale, 3)
    vbp.update(old[V - 1][2], old[V - 1][3], old[V - 1][4], frames_win[V - 1].sb())
    prior.update(old[0][0], old[0][1], frames_win[0].pose())
    for i in range(V - 1):
        j = i + 1
        rels[j].update(old[i][0], old[i][1], old[j][0], old[j][1], frames_win[i].pose(), frames_win[j].pose())
    for rp in rps:
        rp.update(old[rp.index][1], frames_win[rp.index].pose())
    if rec is not None:
        rec["new_pose"] = np.array([f.pose() for f in frames_win[:V]])
        rec["new_sb"] = frames_win[V - 1].sb()


def _factor_state(V, prior, rels, vbp, rps):
    return {"se3": (prior.t.copy(), prior.R.copy(), prior.sqrt_info.copy()),
            "vb": (vbp.VB.copy(), vbp.sqrt_info.copy()),
            "rel": {i: (rels[i].delta_t.copy(), rels[i].delta_R.copy(), rels[i].sqrt_info.copy()) for i in range(1, V)},
            "rp": {rp.index: (rp.R.copy(), rp.sqrt_info.copy()) for rp in rps}}


def make_chain(seed: int, L, rounds: int = 2, cfg: Optional[O.Config] = None, K: int = 10,
               max_gap: int = 1, run_oracle: bool = True, perturb: float = 1.0,
               structured: bool = False, with_yaw: bool = False) -> Chain:
    """Run init_sparsify on V frames, then `rounds` MARGIN_OLD events.  L: int or list of ints."""
    cfg = cfg or O.Config()
    V = cfg.vo_size
    rng = np.random.default_rng(seed)
    Ls = [L] * rounds if np.isscalar(L) else list(L)
    frames = make_frames(rng, cfg, V + 1 + rounds, K=K, max_gap=max_gap)
    expose = ex_pose()
    init_in = O.InitInput(np.array([f.pose() for f in frames[:V]]), np.array([f.sb() for f in frames[:V]]),
                          [frames[i + 1].pre for i in range(V - 1)])
    init_out = O.init_sparsify(init_in, cfg)
    # live factor set (src/estimator.cpp:821-858)
    rels: List[Optional[O.RelativePoseFactor]] = [None]
    for i in range(V - 1):
        rf = O.RelativePoseFactor(init_out.rel_dt[i], init_out.rel_dR[i])
        rf.sqrt_info = init_out.rel_sqrt_info[i]
        rf.setIndex(i, i + 1)
        rels.append(rf)
    prior = O.SE3PriorFactor(init_out.se3_t, R_new=init_out.se3_R)
    prior.sqrt_info = init_out.se3_sqrt_info
    prior.setIndex(0)
    vbp = O.Linear9Factor(init_out.vb)
    vbp.sqrt_info = init_out.vb_sqrt_info
    vbp.setIndex(V - 1)
    rps: List[O.RollPitchFactor] = []
    events: List[MargEvent] = []
    for r in range(rounds):
        win = frames[r:r + V + 1]
        upd: dict = {}
        _perturb_and_update(rng, win, prior, rels, vbp, rps, V, perturb, upd)
        rot_diff = None
        if with_yaw:
            # double2vector (:520-550): the prior factors are rotated by the yaw correction; Marg* then
            # read para_* (not rotated) and these rotated members (Q17)
            rot_diff = O.double2vector_rot_diff(upd["old_R"][0], win[0].pose())
            vbp.VB[6:9] = rot_diff @ vbp.VB[6:9]
            prior.R = rot_diff @ prior.R
        inv_dep, pts_i, pts_j = make_landmarks(rng, cfg, win[0], win[1], Ls[r], expose)
        rp_valid = bool(rps) and rps[0].index == 0
        fin = O.ForwardInput(win[0].pose(), win[1].pose(), expose, inv_dep, pts_i, pts_j,
                             prior.t.copy(), prior.R.copy(), prior.sqrt_info.copy(),
                             rels[1].delta_t.copy(), rels[1].delta_R.copy(), rels[1].sqrt_info.copy(),
                             rp_valid, rps[0].sqrt_info.copy() if rp_valid else None)
        bin_ = O.BackwardInput(win[V - 1].pose(), win[V - 1].sb(), win[V].pose(), win[V].sb(),
                               vbp.VB.copy(), vbp.sqrt_info.copy(), win[V].pre)
        ev = MargEvent(fin, bin_, win[V].raw, win[V].acc0, win[V].gyr0, upd=upd, rot_diff=rot_diff)
        # CombinedFactors members that do not come out of the kernels (:1275-1279)
        ev.pg_meta = {"ts": float(win[0].t), "Ri": upd["old_R"][0].copy(), "ti": upd["old_P"][0].copy(),
                      "rp": (rps[0].R.copy(), rps[0].sqrt_info.copy()) if rp_valid else None}
        fo = O.marg_forward(fin, cfg, structured=structured)
        bo = O.marg_backward(bin_, cfg)
        if run_oracle:
            ev.fwd_out, ev.bwd_out = fo, bo
        events.append(ev)
        # slideWindow factor rotation (src/estimator.cpp:1605-1638)
        for i in range(1, V):
            rels[i].shift()
        for i in range(1, V - 1):
            rels[i], rels[i + 1] = rels[i + 1], rels[i]
        kept = []
        for rp in rps:
            rp.shift()
            if rp.index >= 0:
                kept.append(rp)
        rps = kept
        nrel = O.RelativePoseFactor(bo.rel_dt, bo.rel_dR)
        nrel.sqrt_info = bo.rel_sqrt_info
        nrel.setIndex(V - 2, V - 1)
        rels[V - 1] = nrel
        prior = O.SE3PriorFactor(fo.se3_t, R_new=fo.se3_R)
        prior.sqrt_info = fo.se3_sqrt_info
        prior.setIndex(0)
        vbp = O.Linear9Factor(bo.vb)
        vbp.sqrt_info = bo.vb_sqrt_info
        vbp.setIndex(V - 1)
        nrp = O.RollPitchFactor(R=bo.rp_R)
        nrp.sqrt_info = bo.rp_sqrt_info
        nrp.setIndex(V - 2)      # pushed with index V-1 (:1516), then shifted by slideWindow
        rps.append(nrp)
        ev.state_after = _factor_state(V, prior, rels, vbp, rps)
    return Chain(cfg, init_in, init_out, events)


# ----------------------------------------------------------------------------------------------
# A whole problemSolve() factor list (src/estimator.cpp:1004-1146): the input of the batched
# ceres-Evaluate kernels and of the normal-equation builder.
# ----------------------------------------------------------------------------------------------
@dataclass
class WindowProblem:
    cfg: O.Config
    poses: np.ndarray            # [N,7]   para_Pose
    sbs: np.ndarray              # [N,9]   para_SpeedBias
    ex: np.ndarray               # [1,7]   para_Ex_Pose
    feat: np.ndarray             # [F]     para_Feature
    proj_idx: np.ndarray         # int32 [4,P]: imu_i, imu_j, ex index, feature_index
    proj_obs: np.ndarray         # [5,P]: pts_i.xyz, pts_j.xy
    imu_idx: np.ndarray          # int32 [N-1,2]
    imu_pre: List[O.IntegrationBase]
    rel: List[O.RelativePoseFactor]
    se3: List[O.SE3PriorFactor]
    vb: List[O.Linear9Factor]
    rp: List[O.RollPitchFactor]
    yaw: List[O.YawFactor]


def _rand_sqrt_info(rng, n, scale):
    a = rng.normal(0, 1.0, (n, n))
    return O.llt_upper(scale * scale * (a @ a.T + n * np.eye(n)))


def make_problem(seed: int, n_features: int = 60, cfg: Optional[O.Config] = None, K: int = 10,
                 max_track: int = 8, host0: float = 0.0) -> WindowProblem:
    """N = ALL_BUF_SIZE frames, N-1 IMU factors, `n_features` features with start_frame uniform and
    track length 2..max_track (one ProjectionFactor per later observation, :1062-1092), the prior
    factor set of :1100-1121 (priors from the oracle's init_sparsify on the first V frames, states
    then moved so that no residual is zero) plus one YawFactor (a type the path owns, yaw_factor.h)."""
    cfg = cfg or O.Config()
    V, N = cfg.vo_size, cfg.all_buf_size
    rng = np.random.default_rng(seed)
    frames = make_frames(rng, cfg, N, K=K)
    expose = ex_pose()
    init_in = O.InitInput(np.array([f.pose() for f in frames[:V]]), np.array([f.sb() for f in frames[:V]]),
                          [frames[i + 1].pre for i in range(V - 1)])
    init_out = O.init_sparsify(init_in, cfg)
    rel = []
    for i in range(V - 1):
        rf = O.RelativePoseFactor(init_out.rel_dt[i], init_out.rel_dR[i])
        rf.sqrt_info = init_out.rel_sqrt_info[i]
        rf.setIndex(i, i + 1)
        rel.append(rf)
    se3 = O.SE3PriorFactor(init_out.se3_t, R_new=init_out.se3_R)
    se3.sqrt_info = init_out.se3_sqrt_info
    se3.setIndex(0)
    vbp = O.Linear9Factor(init_out.vb)
    vbp.sqrt_info = init_out.vb_sqrt_info
    vbp.setIndex(V - 1)
    rps = []
    for idx in (2, V - 2):
        rp = O.RollPitchFactor(frames[idx].Q)
        rp.sqrt_info = _rand_sqrt_info(rng, 2, 30.0)
        rp.setIndex(idx)
        rps.append(rp)
    yw = O.YawFactor(frames[V - 1].Q)
    yw.sqrt_info = np.array([[57.0]])
    yw.index = V - 1
    # features: truth landmarks from the (unperturbed) host frame, observed with pixel noise
    qic, tic = O.quat_from_pose(expose), expose[0:3]
    ric = O.q_to_R(qic)
    pidx, pobs, feat = [], [], []
    for f in range(n_features):
        start = int(rng.integers(0, N - 1))
        if host0 > 0.0 and rng.random() < host0:     # fraction of features hosted in the oldest frame
            start = 0
        length = int(rng.integers(2, max_track + 1))
        u, v = rng.uniform(0, IMG_W), rng.uniform(0, IMG_H)
        pts_i = np.array([(u - CX) / FX, (v - CY) / FY, 1.0])
        depth = rng.uniform(1.0, 8.0)
        Ri = O.q_to_R(frames[start].Q)
        pw = Ri @ (ric @ (pts_i * depth) + tic) + frames[start].P
        feat.append((1.0 / depth) * (1.0 + rng.normal(0, 0.01)))
        for j in range(start + 1, min(start + length, N)):
            Rj = O.q_to_R(frames[j].Q)
            pj = ric.T @ (Rj.T @ (pw - frames[j].P) - tic)
            pidx.append((start, j, 0, f))
            pobs.append((pts_i[0], pts_i[1], pts_i[2], pj[0] / pj[2] + rng.normal(0, 1.0 / 460.0),
                         pj[1] / pj[2] + rng.normal(0, 1.0 / 460.0)))
    # post-solve estimates: move every state a little so all residuals are non-trivial
    for f in frames:
        f.P = f.P + rng.normal(0, 0.004, 3)
        f.Q = O.q_normalized(O.q_mul(f.Q, O.SO3.exp(rng.normal(0, 0.001, 3)).q))
        f.V = f.V + rng.normal(0, 0.01, 3)
        f.Ba = f.Ba + rng.normal(0, 0.0005, 3)
        f.Bg = f.Bg + rng.normal(0, 0.00005, 3)
    return WindowProblem(cfg, np.array([f.pose() for f in frames]), np.array([f.sb() for f in frames]),
                         expose.reshape(1, 7), np.array(feat), np.array(pidx, dtype=np.int32).T.copy(),
                         np.array(pobs).T.copy(), np.array([(i, i + 1) for i in range(N - 1)], dtype=np.int32),
                         [frames[i + 1].pre for i in range(N - 1)], rel, [se3], [vbp], rps, [yw])


def cauchy_correct(res: np.ndarray, jacs, a: float):
    """ceres::CauchyLoss(a) through ceres' Corrector (ceres 2.0.0 internal/ceres/corrector.cc,
    loss_function.cc; not under /root/reference): rho' = 1/(1+s/a^2), rho'' < 0 always, so the
    Corrector takes its `rho[2] <= 0` branch: residual and Jacobians are scaled by sqrt(rho')."""
    if not a or a <= 0:
        return res, jacs
    s = float(res @ res)
    sc = math.sqrt(1.0 / (1.0 + s / (a * a)))
    return res * sc, [None if j is None else j * sc for j in jacs]


def eval_problem_oracle(p: WindowProblem, cauchy_a: float = 0.0):
    """Every factor's ceres Evaluate output through the oracle: dict name -> (residuals, [jac blocks])."""
    out = {"proj": [], "imu": [], "rel": [], "se3": [], "vb": [], "rp": [], "yaw": []}
    s = p.cfg.proj_sqrt_info
    for k in range(p.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        pf = O.ProjectionFactor(p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0]), s)
        r, js = pf.EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1]])
        out["proj"].append(cauchy_correct(r, js, cauchy_a))
    for k, (i, j) in enumerate(p.imu_idx):
        out["imu"].append(O.IMUFactor(p.imu_pre[k]).EvaluateCeres([p.poses[i], p.sbs[i], p.poses[j], p.sbs[j]]))
    for rf in p.rel:
        out["rel"].append(cauchy_correct(*rf.EvaluateCeres([p.poses[rf.imu_i], p.poses[rf.imu_j]]), cauchy_a))
    for sf in p.se3:
        out["se3"].append(cauchy_correct(*sf.EvaluateCeres([p.poses[sf.index]]), cauchy_a))
    for vf in p.vb:
        out["vb"].append(cauchy_correct(*vf.EvaluateCeres([p.sbs[vf.index]]), cauchy_a))
    for rp in p.rp:
        out["rp"].append(cauchy_correct(*rp.EvaluateCeres([p.poses[rp.index]]), cauchy_a))
    for yf in p.yaw:
        out["yaw"].append(cauchy_correct(*yf.EvaluateCeres([p.poses[yf.index]]), cauchy_a))
    return out


# ----------------------------------------------------------------------------------------------
# BASELINE configs[3] variant (b): WINDOW_SIZE = 20 with online td estimation, VINS-Mono style (neither
# ProjectionTdFactor nor MarginalizationInfo exists in the reference, SURVEY.md section 0: parity unpinned).
# ----------------------------------------------------------------------------------------------
W20_SEED = seed_for(4, 100)
W20_KW = dict(n_features=2000, max_track=15, host0=0.08)


def w20_problem() -> WindowProblem:
    """WINDOW_SIZE = 20 (21 frames), ~2000 live features with mean track length ~8, ~250 of them hosted in the oldest
    frame (~1750 visual factors to marginalize): the problem behind tests/golden/problem_W20_F2000_td.npz."""
    return make_problem(W20_SEED, cfg=O.Config(all_buf_size=21), **W20_KW)


def make_td_observations(p: WindowProblem, seed: int):
    """Per ProjectionTdFactor members the feature tracker would supply (image velocities, per-observation td,
    rows minus ROW / 2) for every projection factor of `p`: td_obs [8,P] in the C ABI's component order, the
    current para_Td and TR / ROW."""
    rng = np.random.default_rng(seed)
    P = p.proj_idx.shape[1]
    td_obs = np.empty((8, P))
    td_obs[0:4] = rng.normal(0, 0.3, (4, P))          # velocity_i.xy, velocity_j.xy (normalized plane / s)
    td_obs[4:6] = rng.normal(0, 0.004, (2, P))        # td_i, td_j: para_Td when each observation was taken
    td_obs[6:8] = rng.uniform(-240, 240, (2, P))      # row_i, row_j minus ROW / 2
    return {"td_obs": td_obs, "td": np.array([0.007]), "tr_over_row": 0.033 / 480}


def make_window_prior(p: WindowProblem, seed: int, with_td: bool = True):
    """A `last_marginalization_info` in steady state: a prior |r0 + J dx|^2 over EVERY pose / speed-bias of the
    window, the extrinsics and (with_td) the time offset -- what makes n_keep = 15 (N - 1) + 6 + 1 after the
    oldest frame goes.  Synthetic (upper-triangular square-root information with per-family scales), linearized
    a little away from the current estimates.  Returns (keys, J, r0, x0 list)."""
    rng = np.random.default_rng(seed)
    N = p.poses.shape[0]
    keys, scale, x0 = [], [], []
    for i in range(N):
        keys += [("pose", i), ("speed_bias", i)]
        scale += [300.0] * 3 + [600.0] * 3 + [50.0] * 3 + [100.0] * 3 + [2000.0] * 3
        ps = p.poses[i].copy()
        ps[0:3] += rng.normal(0, 0.005, 3)
        q = O.q_normalized(O.q_mul(O.quat_from_pose(ps), O.SO3.exp(rng.normal(0, 0.002, 3)).q))
        ps[3:6], ps[6] = q[1:4], q[0]
        x0 += [ps, p.sbs[i] + rng.normal(0, 0.002, 9)]
    keys.append(("ex_pose", 0))
    scale += [1000.0] * 6
    ex = p.ex[0].copy()
    ex[0:3] += rng.normal(0, 0.001, 3)
    x0.append(ex)
    if with_td:
        keys.append(("td", 0))
        scale.append(1000.0)
        x0.append(np.array([0.0065]))
    n = len(scale)
    J = np.triu(rng.normal(0, 1.0, (n, n)), 1) * 3.0 + np.diag(np.asarray(scale) * rng.uniform(0.7, 1.3, n))
    r0 = rng.normal(0, 0.3, n)
    return keys, J, r0, x0
