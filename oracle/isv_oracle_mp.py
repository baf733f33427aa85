"""50-digit (mpmath) restatement of MargForward / MargBackward -- the adjudicator.

TEST INFRASTRUCTURE ONLY (see oracle/isv_oracle.py header).  Two correct FP64 implementations of
this path (the reference's Eigen code, oracle/isv_oracle.py, the CUDA kernels) differ from each other
by ~cond * eps: the IMU information is covariance^-1 with cond ~1e6, the Schur complement cancels
~1e9-sized numbers down to ~1e3-sized eigenvalues, and the recovered covariances are inverted once
more.  This module evaluates the SAME formulas
(/root/reference/src/estimator.cpp:1149-1539 and the factor headers it calls) in 50-digit
arithmetic, so that tests can measure each FP64 implementation against the exact answer instead of
against each other.  In exact arithmetic the literal dense Schur complement and the structured
(landmarks-first) one coincide, so the forward pass uses the cheap one.

Discrete decisions are mathematical here: eigenvalues are kept iff > ALPHA (round-off-level
eigenvalues of the FP64 code are exact zeros in this arithmetic).
"""
from __future__ import annotations

import mpmath as mp
import numpy as np

mp.mp.dps = 50


def _m(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return mp.matrix([mp.mpf(float(x)) for x in a])
    return mp.matrix([[mp.mpf(float(x)) for x in row] for row in a])


def _np(m):
    return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)])


def _q(ps):
    return [mp.mpf(float(ps[6])), mp.mpf(float(ps[3])), mp.mpf(float(ps[4])), mp.mpf(float(ps[5]))]


def qmul(a, b):
    return [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]]


def qinv(q):
    n2 = sum(x * x for x in q)
    return [q[0] / n2, -q[1] / n2, -q[2] / n2, -q[3] / n2]


def qnormalized(q):
    n = mp.sqrt(sum(x * x for x in q))
    return [x / n for x in q]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def qrot(q, v):
    u = q[1:4]
    uv = [2 * x for x in cross(u, v)]
    c = cross(u, uv)
    return [v[i] + q[0] * uv[i] + c[i] for i in range(3)]


def q2R(q):
    w, x, y, z = q
    return mp.matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def R2q(m):
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = [mp.mpf(0)] * 4
    if t > 0:
        t = mp.sqrt(t + 1)
        q[0] = t / 2
        t = 1 / (2 * t)
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = mp.sqrt(m[i, i] - m[j, j] - m[k, k] + 1)
        q[1 + i] = t / 2
        t = 1 / (2 * t)
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def skew(v):
    return mp.matrix([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def so3_log(q):
    q = qnormalized(q)
    n2 = q[1] ** 2 + q[2] ** 2 + q[3] ** 2
    if n2 == 0:
        return [mp.mpf(0)] * 3
    n = mp.sqrt(n2)
    at = mp.atan2(-n, -q[0]) if q[0] < 0 else mp.atan2(n, q[0])
    f = 2 * at / n
    return [f * q[1], f * q[2], f * q[3]]


def right_jacobian_inv(phi):
    n2 = sum(x * x for x in phi)
    H = skew(phi)
    J = mp.eye(3) + H / 2
    if n2 == 0:
        return J
    n = mp.sqrt(n2)
    return J + H * H * (1 / n2 - (1 + mp.cos(n)) / (2 * n * mp.sin(n)))


def _set(M, r0, c0, B, sign=1):
    for i in range(B.rows):
        for j in range(B.cols):
            M[r0 + i, c0 + j] = sign * B[i, j]


def relpose_jac(PSi, PSj, dR):
    Qi, Qj = _q(PSi), _q(PSj)
    Ri, Rj = q2R(Qi), q2R(Qj)
    d = [mp.mpf(float(PSj[k])) - mp.mpf(float(PSi[k])) for k in range(3)]
    tij = qrot(qinv(Qi), d)
    lg = so3_log(R2q(dR * Rj.T * Ri))
    J = right_jacobian_inv(lg)
    Ji, Jj = mp.zeros(6, 6), mp.zeros(6, 6)
    _set(Ji, 0, 0, Ri.T)
    _set(Ji, 0, 3, skew(tij), -1)
    _set(Ji, 3, 3, J)
    _set(Jj, 0, 0, Ri.T, -1)
    _set(Jj, 3, 3, J * Ri.T * Rj, -1)
    return Ji, Jj, tij


def se3prior_jac(PS, Rprior):
    ri = qnormalized(_q(PS))
    rp = R2q(Rprior)
    lg = so3_log(qmul([rp[0], -rp[1], -rp[2], -rp[3]], ri))
    J = mp.eye(6)
    _set(J, 3, 3, right_jacobian_inv(lg))
    return J


def rollpitch_jac(PS, Rmeas):
    ri = qnormalized(_q(PS))
    rm = R2q(Rmeas)
    a = qrot([ri[0], -ri[1], -ri[2], -ri[3]], [mp.mpf(0), mp.mpf(0), mp.mpf(-1)])
    r3 = qrot(rm, a)
    SR = skew(r3) * q2R(rm)
    J = mp.zeros(2, 6)
    for r in range(2):
        for c in range(3):
            J[r, 3 + c] = SR[r, c]
    return J


def chol_upper(M):
    """LLT(M).matrixL().transpose()"""
    return mp.cholesky(M).T


def _recover(J, Sigma):
    cov = J * Sigma * J.T
    return chol_upper(mp.inverse(cov))


def _trunc_pinv(A, alpha):
    A = (A + A.T) / 2
    w, V = mp.eigsy(A)
    n = A.rows
    S = mp.zeros(n, n)
    rank = 0
    for k in range(n):
        if w[k] > alpha:
            rank += 1
            v = V[:, k]
            S += v * v.T / w[k]
    return S, rank, [float(x) for x in w]


def marg_backward_mp(inp, cfg):
    """MargBackward in 50 digits.  Returns dict of float64 arrays (rel_sqrt_info, vb_sqrt_info,
    rp_sqrt_info, rel_dt, rel_dR, rank, eigvals)."""
    pre = inp.pre
    G = [mp.mpf(float(x)) for x in cfg.G]
    s = mp.mpf(float(pre.sum_dt))
    Jp = _m(pre.jacobian)
    P = _m(pre.covariance)
    Qi, Qj = _q(inp.pose_i), _q(inp.pose_j)
    Pi = [mp.mpf(float(x)) for x in inp.pose_i[0:3]]
    Pj = [mp.mpf(float(x)) for x in inp.pose_j[0:3]]
    Vi = [mp.mpf(float(x)) for x in inp.sb_i[0:3]]
    Vj = [mp.mpf(float(x)) for x in inp.sb_j[0:3]]
    Bgi = [mp.mpf(float(x)) for x in inp.sb_i[6:9]]
    lin_bg = [mp.mpf(float(x)) for x in pre.linearized_bg]
    dq = [mp.mpf(float(x)) for x in pre.delta_q]
    dq_dbg = Jp[3:6, 12:15]
    th = dq_dbg * mp.matrix([Bgi[k] - lin_bg[k] for k in range(3)])
    cq = qmul(dq, [mp.mpf(1), th[0] / 2, th[1] / 2, th[2] / 2])
    Qi_inv = qinv(Qi)
    Ri_inv = q2R(Qi_inv)
    v1 = qrot(Qi_inv, [G[k] * s * s / 2 + Pj[k] - Pi[k] - Vi[k] * s for k in range(3)])
    v2 = qrot(Qi_inv, [G[k] * s + Vj[k] - Vi[k] for k in range(3)])

    def qleft(q):
        M = mp.zeros(4, 4)
        M[0, 0] = q[0]
        for i in range(3):
            M[0, 1 + i] = -q[1 + i]
            M[1 + i, 0] = q[1 + i]
        _set(M, 1, 1, q[0] * mp.eye(3) + skew(q[1:4]))
        return M

    def qright(q):
        M = mp.zeros(4, 4)
        M[0, 0] = q[0]
        for i in range(3):
            M[0, 1 + i] = -q[1 + i]
            M[1 + i, 0] = q[1 + i]
        _set(M, 1, 1, q[0] * mp.eye(3) - skew(q[1:4]))
        return M

    QjinvQi = qmul(qinv(Qj), Qi)
    j0, j1, j2, j3 = mp.zeros(15, 6), mp.zeros(15, 9), mp.zeros(15, 6), mp.zeros(15, 9)
    _set(j0, 0, 0, Ri_inv, -1)
    _set(j0, 0, 3, skew(v1))
    _set(j0, 3, 3, (qleft(QjinvQi) * qright(cq))[1:4, 1:4], -1)
    _set(j0, 6, 3, skew(v2))
    _set(j1, 0, 0, Ri_inv * s, -1)
    _set(j1, 0, 3, Jp[0:3, 9:12], -1)
    _set(j1, 0, 6, Jp[0:3, 12:15], -1)
    _set(j1, 3, 6, qleft(qmul(QjinvQi, dq))[1:4, 1:4] * dq_dbg, -1)   # Q9
    _set(j1, 6, 0, Ri_inv, -1)
    _set(j1, 6, 3, Jp[6:9, 9:12], -1)
    _set(j1, 6, 6, Jp[6:9, 12:15], -1)
    _set(j1, 9, 3, mp.eye(3), -1)
    _set(j1, 12, 6, mp.eye(3), -1)
    _set(j2, 0, 0, Ri_inv)
    _set(j2, 3, 3, qleft(qmul(qmul(qinv(cq), Qi_inv), Qj))[1:4, 1:4])
    _set(j3, 6, 0, Ri_inv)
    _set(j3, 9, 3, mp.eye(3))
    _set(j3, 12, 6, mp.eye(3))
    Jall = mp.zeros(15, 30)   # OrderMap: T_V@0, VB_V@6, T_{V-1}@15, VB_{V-1}@21
    _set(Jall, 0, 15, j0)
    _set(Jall, 0, 21, j1)
    _set(Jall, 0, 0, j2)
    _set(Jall, 0, 6, j3)
    Om = mp.inverse(P)
    Lam = Jall.T * Om * Jall
    svb = _m(inp.vb_sqrt_info)
    pri = svb.T * svb
    for i in range(9):
        for j in range(9):
            Lam[21 + i, 21 + j] += pri[i, j]
    Lrr, Lrm, Lmm = Lam[0:21, 0:21], Lam[0:21, 21:30], Lam[21:30, 21:30]
    Lprior = Lrr - Lrm * mp.inverse(Lmm) * Lrm.T
    Sigma, rank, w = _trunc_pinv(Lprior, mp.mpf(float(cfg.alpha)))
    Rij = q2R(qmul(qinv(Qi), Qj))
    Ji, Jj, tij = relpose_jac(inp.pose_i, inp.pose_j, Rij)
    Jrel = mp.zeros(6, 21)
    _set(Jrel, 0, 15, Ji)
    _set(Jrel, 0, 0, Jj)
    Jvb = mp.zeros(9, 21)
    _set(Jvb, 0, 6, mp.eye(9))
    Jrp = mp.zeros(2, 21)
    _set(Jrp, 0, 15, rollpitch_jac(inp.pose_i, q2R(Qi)))
    return {"rel_sqrt_info": _np(_recover(Jrel, Sigma)), "vb_sqrt_info": _np(_recover(Jvb, Sigma)),
            "rp_sqrt_info": _np(_recover(Jrp, Sigma)), "rel_dt": np.array([float(x) for x in tij]),
            "rel_dR": _np(Rij), "rank": rank, "eigvals": np.array(w)}


def marg_forward_mp(inp, cfg):
    """MargForward in 50 digits (landmarks eliminated first; exact arithmetic makes this identical to
    the dense FullPivLU route).  Full-rank branch only (qr.rank()==6)."""
    L = int(inp.inv_dep.shape[0])
    Qi, Qj, qic = _q(inp.pose0), _q(inp.pose1), _q(inp.ex_pose)
    Pi = [mp.mpf(float(x)) for x in inp.pose0[0:3]]
    Pj = [mp.mpf(float(x)) for x in inp.pose1[0:3]]
    tic = [mp.mpf(float(x)) for x in inp.ex_pose[0:3]]
    Ri, Rj, ric = q2R(Qi), q2R(Qj), q2R(qic)
    sp = _m(cfg.proj_sqrt_info)
    Om = sp.T * sp
    B = ric.T * Rj.T
    C = B * Ri
    Ap = C * ric
    H = mp.zeros(12, 12)
    S = mp.zeros(12, 12)
    for k in range(L):
        pts_i = [mp.mpf(float(x)) for x in inp.pts_i[k]]
        lam = mp.mpf(float(inp.inv_dep[k]))
        pc = [x / lam for x in pts_i]
        pim = [a + b for a, b in zip(qrot(qic, pc), tic)]
        pw = [a + b for a, b in zip(qrot(Qi, pim), Pi)]
        pj = qrot(qinv(Qj), [pw[i] - Pj[i] for i in range(3)])
        cj = qrot(qinv(qic), [pj[i] - tic[i] for i in range(3)])
        dep = cj[2]
        red = mp.matrix([[1 / dep, 0, -cj[0] / (dep * dep)], [0, 1 / dep, -cj[1] / (dep * dep)]])
        jaco_i = mp.zeros(3, 6)
        _set(jaco_i, 0, 0, B)
        _set(jaco_i, 0, 3, C * skew(pim), -1)
        jaco_j = mp.zeros(3, 6)
        _set(jaco_j, 0, 0, B, -1)
        _set(jaco_j, 0, 3, ric.T * skew(pj))
        Jw = mp.zeros(2, 12)
        _set(Jw, 0, 0, red * jaco_j)
        _set(Jw, 0, 6, red * jaco_i)
        jl = red * (Ap * mp.matrix(pts_i)) * (-1 / (lam * lam))
        Hk = Jw.T * Om * Jw
        bk = Jw.T * Om * jl
        dk = (jl.T * Om * jl)[0, 0]
        H += Hk
        S += Hk - bk * bk.T / dk
    Jp = se3prior_jac(inp.pose0, _m(inp.prior_R))
    s_p = _m(inp.prior_sqrt_info)
    add = mp.zeros(12, 12)
    _set(add, 6, 6, Jp.T * (s_p.T * s_p) * Jp)
    Jri, Jrj, _ = relpose_jac(inp.pose0, inp.pose1, _m(inp.rel_dR))
    s_r = _m(inp.rel_sqrt_info)
    Jst = mp.zeros(6, 12)
    _set(Jst, 0, 0, Jrj)
    _set(Jst, 0, 6, Jri)
    add += Jst.T * (s_r.T * s_r) * Jst
    H += add
    S += add
    Rij = q2R(qmul(qinv(Qi), Qj))
    Gi, Gj, tij = relpose_jac(inp.pose0, inp.pose1, Rij)
    J = mp.zeros(6, 12)
    _set(J, 0, 0, Gi)
    _set(J, 0, 6, Gj)
    Jpinv = J.T * mp.inverse(J * J.T)
    rpOmega = Jpinv.T * H * Jpinv
    Lprior = S[0:6, 0:6] - S[0:6, 6:12] * mp.inverse(S[6:12, 6:12]) * S[6:12, 0:6]
    Jr = se3prior_jac(inp.pose1, Rj)
    covi = Jr * mp.inverse(Lprior) * Jr.T
    return {"se3_sqrt_info": _np(chol_upper(mp.inverse(covi))), "pg_sqrt_info": _np(chol_upper(rpOmega)),
            "pg_covRel": _np(mp.inverse(rpOmega)), "pg_dt": np.array([float(x) for x in tij]), "pg_dR": _np(Rij)}
