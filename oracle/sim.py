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


@dataclass
class Chain:
    cfg: O.Config
    init_in: O.InitInput
    init_out: O.InitOutput
    events: List[MargEvent]


def _perturb_and_update(rng, frames_win: List[Frame], prior: O.SE3PriorFactor, rels: List[O.RelativePoseFactor],
                        vbp: O.Linear9Factor, rps: List[O.RollPitchFactor], V: int, scale: float):
    """Mimic problemSolve(): states move a little, then factor->update re-centres the
    pseudo-measurements  (src/estimator.cpp:1133-1144)."""
    old = [(f.P.copy(), O.q_to_R(f.Q), f.V.copy(), f.Ba.copy(), f.Bg.copy()) for f in frames_win]
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


def make_chain(seed: int, L, rounds: int = 2, cfg: Optional[O.Config] = None, K: int = 10,
               max_gap: int = 1, run_oracle: bool = True, perturb: float = 1.0,
               structured: bool = False) -> Chain:
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
        _perturb_and_update(rng, win, prior, rels, vbp, rps, V, perturb)
        inv_dep, pts_i, pts_j = make_landmarks(rng, cfg, win[0], win[1], Ls[r], expose)
        rp_valid = bool(rps) and rps[0].index == 0
        fin = O.ForwardInput(win[0].pose(), win[1].pose(), expose, inv_dep, pts_i, pts_j,
                             prior.t.copy(), prior.R.copy(), prior.sqrt_info.copy(),
                             rels[1].delta_t.copy(), rels[1].delta_R.copy(), rels[1].sqrt_info.copy(),
                             rp_valid, rps[0].sqrt_info.copy() if rp_valid else None)
        bin_ = O.BackwardInput(win[V - 1].pose(), win[V - 1].sb(), win[V].pose(), win[V].sb(),
                               vbp.VB.copy(), vbp.sqrt_info.copy(), win[V].pre)
        ev = MargEvent(fin, bin_, win[V].raw, win[V].acc0, win[V].gyr0)
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
    return Chain(cfg, init_in, init_out, events)
