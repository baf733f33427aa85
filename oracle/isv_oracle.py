"""CPU oracle for the IS-VINS marginalization + sparsification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker / the reported CPU baseline.  The product path
(``is_vins_b200``) never imports this module and fails loudly without its CUDA library.

PARITY UNPINNED: the reference (lyeemax/IS-VINS) ships no tests, golden vectors or fixtures for
this path and cannot be compiled here (Eigen 3.3.4, Ceres 2.0.0, Sophus,
OpenCV are not vendored and absent from the image).  This file is therefore a *restatement* in
NumPy float64 of the reference's algorithm, following the reference line by line (each function
cites the file:line it follows, paths relative to /root/reference).  Trust is earned by
(1) the reference's own finite-difference recipes re-run in tests/, (2) algebraic identities
(Schur == block of inverse, marginal preservation, KLD >= 0), (3) an independent C restatement
(oracle/isv_ref.c) and (4) an mpmath 50-digit mode (oracle/isv_oracle_mp.py) agreeing with it.

Conventions (all from the reference):
  pose block   = [px,py,pz,qx,qy,qz,qw]   (src/estimator.cpp:474-487)
  speed-bias   = [v(3), ba(3), bg(3)]     (src/estimator.cpp:489-499)
  quaternions are handled internally as (w, x, y, z) numpy arrays with Eigen semantics
  (no implicit normalisation; q*v uses the unit-quaternion formula; inverse = conj/|q|^2).
Third-party arithmetic restated here from the published algorithms (not under /root/reference):
  Eigen 3.3.4 (README.md:22): Quaternion ops, FullPivLU::solve, PartialPivLU inverse, LLT,
  FullPivHouseholderQR rank/solve, SelfAdjointEigenSolver (LAPACK dsyev used instead: any
  backward-stable solver is equivalent up to the conditioning-scaled tolerance), BDCSVD
  (LAPACK dgesdd).  Sophus (unpinned): SO3 ctor/log/exp/inverse/matrix, SE3::Adj.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------
# constants: config/euroc_config.yaml:57-61,83,86 ; include/parameters.h:35-40,89-96
# ----------------------------------------------------------------------------------------------
O_P, O_R, O_V, O_BA, O_BG = 0, 3, 6, 9, 12
SOPHUS_EPS = 1e-10          # Sophus::Constants<double>::epsilon()
SOPHUS_EPS_SQRT = 1e-5      # Sophus::Constants<double>::epsilonSqrt()
ESTIMATOR_EPS = 1e-16       # src/estimator.cpp:8


@dataclass
class Config:
    """Hot-path knobs.  Defaults = config/euroc_config.yaml."""
    alpha: float = 0.1                 # yaml:86, eigenvalue cut (strict >)
    pixel_sqrt_info: float = 460.0     # yaml:83 -> ProjectionFactor::sqrt_info = 460*I2 (estimator.cpp:35)
    acc_n: float = 0.22627
    gyr_n: float = 0.003988
    acc_w: float = 0.001
    gyr_w: float = 0.0001
    g_norm: float = 9.81007
    vo_size: int = 8                   # include/parameters.h:35
    all_buf_size: int = 18             # include/parameters.h:40

    @property
    def G(self) -> np.ndarray:         # src/parameters.cpp:19,96 : G = (0,0,+g)
        return np.array([0.0, 0.0, self.g_norm])

    @property
    def proj_sqrt_info(self) -> np.ndarray:
        return self.pixel_sqrt_info * np.eye(2)


# ----------------------------------------------------------------------------------------------
# Eigen quaternion semantics
# ----------------------------------------------------------------------------------------------
def quat_from_pose(ps) -> np.ndarray:
    """Quaterniond(PS[6], PS[3], PS[4], PS[5]) -> (w,x,y,z)."""
    return np.array([ps[6], ps[3], ps[4], ps[5]], dtype=np.float64)


def q_mul(a, b) -> np.ndarray:
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ])


def q_conj(q) -> np.ndarray:
    return np.array([q[0], -q[1], -q[2], -q[3]])


def q_inv(q) -> np.ndarray:
    """Eigen Quaternion::inverse(): conjugate / squaredNorm (Q13)."""
    n2 = float(np.dot(q, q))
    return q_conj(q) / n2


def q_normalized(q) -> np.ndarray:
    return np.asarray(q, dtype=np.float64) / math.sqrt(float(np.dot(q, q)))


def q_rot(q, v) -> np.ndarray:
    """Eigen QuaternionBase::_transformVector: v + w*(2 u x v) + u x (2 u x v) (Q13)."""
    u = np.asarray(q[1:4])
    uv = 2.0 * np.cross(u, v)
    return np.asarray(v) + q[0] * uv + np.cross(u, uv)


def q_to_R(q) -> np.ndarray:
    """Eigen QuaternionBase::toRotationMatrix() (assumes, does not enforce, unit norm)."""
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1 - (txx + tyy)],
    ])


def R_to_q(m) -> np.ndarray:
    """Eigen quaternion-from-matrix (Shepperd branch on trace, then largest diagonal) (Q14)."""
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)  # w,x,y,z
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
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
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


# ----------------------------------------------------------------------------------------------
# include/utility/utility.h
# ----------------------------------------------------------------------------------------------
def skew(v) -> np.ndarray:
    """Utility::skewSymmetric  utility.h:26-34"""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def deltaQ(theta) -> np.ndarray:
    """Utility::deltaQ  utility.h:11-24 : (1, theta/2), NOT normalised."""
    return np.array([1.0, theta[0] / 2.0, theta[1] / 2.0, theta[2] / 2.0])


def Qleft(q) -> np.ndarray:
    """Utility::Qleft  utility.h:47-54"""
    ans = np.zeros((4, 4))
    ans[0, 0] = q[0]
    ans[0, 1:] = -np.asarray(q[1:4])
    ans[1:, 0] = q[1:4]
    ans[1:, 1:] = q[0] * np.eye(3) + skew(q[1:4])
    return ans


def Qright(p) -> np.ndarray:
    """Utility::Qright  utility.h:57-64"""
    ans = np.zeros((4, 4))
    ans[0, 0] = p[0]
    ans[0, 1:] = -np.asarray(p[1:4])
    ans[1:, 0] = p[1:4]
    ans[1:, 1:] = p[0] * np.eye(3) - skew(p[1:4])
    return ans


def R2ypr(R) -> np.ndarray:
    """Utility::R2ypr  utility.h:66-82 (degrees)."""
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = math.atan2(n[1], n[0])
    p = math.atan2(-n[2], n[0] * math.cos(y) + n[1] * math.sin(y))
    r = math.atan2(a[0] * math.sin(y) - a[1] * math.cos(y), -o[0] * math.sin(y) + o[1] * math.cos(y))
    return np.array([y, p, r]) / math.pi * 180.0


def ypr2R(ypr) -> np.ndarray:
    """Utility::ypr2R  utility.h:84-109 (degrees)."""
    y, p, r = (ypr[0] / 180.0 * math.pi, ypr[1] / 180.0 * math.pi, ypr[2] / 180.0 * math.pi)
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1.0]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1.0, 0], [-math.sin(p), 0, math.cos(p)]])
    Rx = np.array([[1.0, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    return Rz @ Ry @ Rx


def pseudo_inverse(a, epsilon=np.finfo(np.float64).eps) -> np.ndarray:
    """Utility::pseudoInverse  utility.h:144-156 : thin SVD, Eigen *relative* threshold
    epsilon*max(rows,cols) of the largest singular value (Eigen SVDBase::rank())."""
    a = np.asarray(a, dtype=np.float64)
    U, s, Vt = np.linalg.svd(a, full_matrices=False)
    thr = epsilon * max(a.shape)
    if s.size == 0 or s[0] == 0:
        rank = 0
    else:
        premult = max(s[0] * thr, np.finfo(np.float64).tiny)
        rank = int(np.sum(s > premult))
    tmp = U[:, :rank].T
    tmp = (1.0 / s[:rank])[:, None] * tmp
    return Vt[:rank, :].T @ tmp


# ----------------------------------------------------------------------------------------------
# Sophus SO3 (external; formulas per SURVEY.md section 9)
# ----------------------------------------------------------------------------------------------
class SO3:
    __slots__ = ("q",)

    def __init__(self, q):
        self.q = np.asarray(q, dtype=np.float64)

    @staticmethod
    def from_quat(q) -> "SO3":
        """SO3d(Quaterniond): normalises."""
        return SO3(q_normalized(q))

    @staticmethod
    def from_matrix(R) -> "SO3":
        """SO3d(Matrix3d): asserts orthogonality (not re-checked here), Eigen matrix->quaternion."""
        return SO3(R_to_q(R))

    def inverse(self) -> "SO3":
        return SO3(q_conj(self.q))

    def __mul__(self, other):
        if isinstance(other, SO3):
            q = q_mul(self.q, other.q)
            # Sophus SO3::operator* renormalises when the product drifted from unit norm
            n2 = float(np.dot(q, q))
            if n2 != 1.0:
                q = q * (2.0 / (1.0 + n2))
            return SO3(q)
        return q_rot(self.q, np.asarray(other, dtype=np.float64))

    def matrix(self) -> np.ndarray:
        return q_to_R(self.q)

    def log(self) -> np.ndarray:
        w = self.q[0]
        v = self.q[1:4]
        n2 = float(np.dot(v, v))
        if n2 < SOPHUS_EPS * SOPHUS_EPS:
            f = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w)
        else:
            n = math.sqrt(n2)
            at = math.atan2(-n, -w) if w < 0 else math.atan2(n, w)
            f = 2.0 * at / n
        return f * v

    @staticmethod
    def exp(omega) -> "SO3":
        omega = np.asarray(omega, dtype=np.float64)
        t2 = float(np.dot(omega, omega))
        if t2 < SOPHUS_EPS * SOPHUS_EPS:
            t4 = t2 * t2
            im = 0.5 - t2 / 48.0 + t4 / 3840.0
            re = 1.0 - t2 / 8.0 + t4 / 384.0
        else:
            t = math.sqrt(t2)
            im = math.sin(0.5 * t) / t
            re = math.cos(0.5 * t)
        return SO3(np.array([re, im * omega[0], im * omega[1], im * omega[2]]))


def right_jacobian_inv_SO3(phi) -> np.ndarray:
    """Sophus::rightJacobianInvSO3  include/utility/sophus_utils.hpp:194-236"""
    phi = np.asarray(phi, dtype=np.float64)
    n2 = float(np.dot(phi, phi))
    ph = skew(phi)
    ph2 = ph @ ph
    J = np.eye(3) + ph / 2.0
    if n2 > SOPHUS_EPS:
        n = math.sqrt(n2)
        assert n <= math.pi + SOPHUS_EPS
        if n < math.pi - SOPHUS_EPS_SQRT:
            J = J + ph2 * (1.0 / n2 - (1.0 + math.cos(n)) / (2.0 * n * math.sin(n)))
        else:
            J = J + ph2 / (math.pi * math.pi)
    else:
        J = J + ph2 / 12.0
    return J


def se3_adj(R, t) -> np.ndarray:
    """Sophus::SE3d::Adj() for tangent order (upsilon, omega)."""
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[:3, 3:] = skew(t) @ R
    A[3:, 3:] = R
    return A


# ----------------------------------------------------------------------------------------------
# Eigen dense decompositions restated (literal algorithm classes)
# ----------------------------------------------------------------------------------------------
def full_piv_lu_solve_identity(A) -> np.ndarray:
    """A.fullPivLu().solve(Identity)  (src/estimator.cpp:814,1286,1417).

    Eigen FullPivLU: at step k pick the largest |entry| of the remaining corner, swap row and
    column, eliminate (no blocking).  solve() uses the rank found with Eigen's default threshold
    (eps * diagonal size) and zero-fills the dependent unknowns.
    """
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    lu = A.copy()
    rowp = np.arange(n)
    colp = np.arange(n)
    maxpivot = 0.0
    nonzero = n
    for k in range(n):
        sub = np.abs(lu[k:, k:])
        idx = int(np.argmax(sub))
        r, c = divmod(idx, n - k)
        big = sub[r, c]
        if big == 0.0:
            nonzero = k
            break
        maxpivot = max(maxpivot, big)
        r += k
        c += k
        if r != k:
            lu[[k, r], :] = lu[[r, k], :]
            rowp[[k, r]] = rowp[[r, k]]
        if c != k:
            lu[:, [k, c]] = lu[:, [c, k]]
            colp[[k, c]] = colp[[c, k]]
        if k < n - 1:
            lu[k + 1:, k] /= lu[k, k]
            lu[k + 1:, k + 1:] -= np.outer(lu[k + 1:, k], lu[k, k + 1:])
    thr = np.finfo(np.float64).eps * n
    diag = np.abs(np.diag(lu))
    rank = int(np.sum(diag[:nonzero] > maxpivot * thr))
    # solve P A Q = L U  ->  X = Q * U^-1 L^-1 P * I
    B = np.eye(n)[rowp, :]
    L = np.tril(lu, -1) + np.eye(n)
    import scipy.linalg as sla
    c = sla.solve_triangular(L, B, lower=True, unit_diagonal=True)
    X = np.zeros((n, n))
    y = sla.solve_triangular(lu[:rank, :rank], c[:rank, :], lower=False)
    X[colp[:rank], :] = y
    return X


def partial_piv_inverse(A) -> np.ndarray:
    """MatrixXd::inverse() = PartialPivLU inverse (Q11). LAPACK dgetrf/dgetri is the same class."""
    return np.linalg.inv(np.asarray(A, dtype=np.float64))


def llt_upper(M) -> np.ndarray:
    """Eigen::LLT<MatrixXd>(M).matrixL().transpose(): upper U with U^T U = M; reads only the
    lower triangle of M."""
    M = np.asarray(M, dtype=np.float64)
    Ml = np.tril(M) + np.tril(M, -1).T
    n = M.shape[0]
    L = np.zeros((n, n))
    for j in range(n):
        d = Ml[j, j] - np.dot(L[j, :j], L[j, :j])
        if not d > 0:
            # Eigen LLT reports NumericalIssue and leaves garbage; oracle makes it explicit.
            L[j:, j] = np.nan
            return L.T
        L[j, j] = math.sqrt(d)
        if j + 1 < n:
            L[j + 1:, j] = (Ml[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L.T


def self_adjoint_eig(M) -> Tuple[np.ndarray, np.ndarray]:
    """Eigen::SelfAdjointEigenSolver: reads the lower triangle, eigenvalues ascending."""
    M = np.asarray(M, dtype=np.float64)
    Ml = np.tril(M) + np.tril(M, -1).T
    w, V = np.linalg.eigh(Ml)
    return w, V


def full_piv_householder_qr(A, threshold):
    """Eigen::FullPivHouseholderQR(A) with setThreshold(threshold): returns (rank, solve_fn).

    rank() counts pivots |R_kk| > |maxpivot| * threshold  (src/estimator.cpp:1304-1307).
    """
    A = np.array(A, dtype=np.float64)
    n_r, n_c = A.shape
    size = min(n_r, n_c)
    qr = A.copy()
    hcoeffs = np.zeros(size)
    col_perm = np.arange(n_c)
    row_trans = np.arange(size)
    maxpivot = 0.0
    nonzero = size
    precision = np.finfo(np.float64).eps * size
    biggest = 0.0
    for k in range(size):
        sub = np.abs(qr[k:, k:])
        idx = int(np.argmax(sub))
        r, c = divmod(idx, n_c - k)
        big_in_corner = sub[r, c]
        r += k
        c += k
        if k == 0:
            biggest = big_in_corner
        if big_in_corner <= abs(biggest) * precision or big_in_corner == 0.0:
            nonzero = k
            for i in range(k, size):
                row_trans[i] = i
                hcoeffs[i] = 0.0
            break
        row_trans[k] = r
        if r != k:
            qr[[k, r], k:] = qr[[r, k], k:]
        if c != k:
            qr[:, [k, c]] = qr[:, [c, k]]
            col_perm[[k, c]] = col_perm[[c, k]]
        # Householder on qr[k:, k]
        x = qr[k:, k].copy()
        tail2 = float(np.dot(x[1:], x[1:]))
        c0 = x[0]
        if tail2 <= np.finfo(np.float64).tiny:
            tau = 0.0
            beta = c0
            ess = np.zeros(x.size - 1)
        else:
            beta = math.sqrt(c0 * c0 + tail2)
            if c0 >= 0:
                beta = -beta
            ess = x[1:] / (c0 - beta)
            tau = (beta - c0) / beta
        qr[k, k] = beta
        qr[k + 1:, k] = ess
        hcoeffs[k] = tau
        maxpivot = max(maxpivot, abs(beta))
        if tau != 0.0 and k + 1 < n_c:
            v = np.concatenate(([1.0], ess))
            blk = qr[k:, k + 1:]
            blk -= tau * np.outer(v, v @ blk)
    diag = np.abs(np.diag(qr))[:nonzero]
    rank = int(np.sum(diag > maxpivot * threshold))

    def solve(B):
        B = np.array(B, dtype=np.float64)
        c = B.copy()
        for k in range(min(nonzero, size)):
            r = row_trans[k]
            if r != k:
                c[[k, r], :] = c[[r, k], :]
            v = np.concatenate(([1.0], qr[k + 1:, k]))
            c[k:, :] -= hcoeffs[k] * np.outer(v, v @ c[k:, :])
        import scipy.linalg as sla
        y = sla.solve_triangular(qr[:rank, :rank], c[:rank, :], lower=False)
        X = np.zeros((n_c, B.shape[1]))
        X[col_perm[:rank], :] = y
        return X

    return rank, solve


# ----------------------------------------------------------------------------------------------
# include/factor/integration_base.h
# ----------------------------------------------------------------------------------------------
class IntegrationBase:
    """include/factor/integration_base.h:10-208"""

    def __init__(self, acc_0, gyr_0, linearized_ba, linearized_bg, cfg: Config):
        self.cfg = cfg
        self.acc_0 = np.array(acc_0, dtype=np.float64)
        self.gyr_0 = np.array(gyr_0, dtype=np.float64)
        self.linearized_acc = self.acc_0.copy()
        self.linearized_gyr = self.gyr_0.copy()
        self.linearized_ba = np.array(linearized_ba, dtype=np.float64)
        self.linearized_bg = np.array(linearized_bg, dtype=np.float64)
        self.jacobian = np.eye(15)
        self.covariance = np.zeros((15, 15))
        self.sum_dt = 0.0
        self.delta_p = np.zeros(3)
        self.delta_q = np.array([1.0, 0, 0, 0])
        self.delta_v = np.zeros(3)
        n = np.zeros((18, 18))                                # :21-27
        n[0:3, 0:3] = cfg.acc_n ** 2 * np.eye(3)
        n[3:6, 3:6] = cfg.gyr_n ** 2 * np.eye(3)
        n[6:9, 6:9] = cfg.acc_n ** 2 * np.eye(3)
        n[9:12, 9:12] = cfg.gyr_n ** 2 * np.eye(3)
        n[12:15, 12:15] = cfg.acc_w ** 2 * np.eye(3)
        n[15:18, 15:18] = cfg.gyr_w ** 2 * np.eye(3)
        self.noise = n
        self.dt_buf: List[float] = []
        self.acc_buf: List[np.ndarray] = []
        self.gyr_buf: List[np.ndarray] = []

    def push_back(self, dt, acc, gyr):                        # :30-36
        self.dt_buf.append(float(dt))
        self.acc_buf.append(np.array(acc, dtype=np.float64))
        self.gyr_buf.append(np.array(gyr, dtype=np.float64))
        self.propagate(dt, acc, gyr)

    def repropagate(self, ba, bg):                            # :38-52
        self.sum_dt = 0.0
        self.acc_0 = self.linearized_acc.copy()
        self.gyr_0 = self.linearized_gyr.copy()
        self.delta_p = np.zeros(3)
        self.delta_q = np.array([1.0, 0, 0, 0])
        self.delta_v = np.zeros(3)
        self.linearized_ba = np.array(ba, dtype=np.float64)
        self.linearized_bg = np.array(bg, dtype=np.float64)
        self.jacobian = np.eye(15)
        self.covariance = np.zeros((15, 15))
        for dt, a, g in zip(self.dt_buf, self.acc_buf, self.gyr_buf):
            self.propagate(dt, a, g)

    def propagate(self, _dt, _acc_1, _gyr_1):                 # :130-158 + :54-128
        _acc_1 = np.asarray(_acc_1, dtype=np.float64)
        _gyr_1 = np.asarray(_gyr_1, dtype=np.float64)
        dt = float(_dt)
        a0, g0 = self.acc_0, self.gyr_0
        dq, dp, dv = self.delta_q, self.delta_p, self.delta_v
        ba, bg = self.linearized_ba, self.linearized_bg
        un_acc_0 = q_rot(dq, a0 - ba)
        un_gyr = 0.5 * (g0 + _gyr_1) - bg
        rq = q_mul(dq, np.array([1.0, un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2]))
        un_acc_1 = q_rot(rq, _acc_1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        rp = dp + dv * dt + 0.5 * un_acc * dt * dt
        rv = dv + un_acc * dt
        # jacobian / covariance  (:76-126)
        w_x = 0.5 * (g0 + _gyr_1) - bg
        a_0_x = a0 - ba
        a_1_x = _acc_1 - ba
        R_w_x, R_a_0_x, R_a_1_x = skew(w_x), skew(a_0_x), skew(a_1_x)
        Rd = q_to_R(dq)
        Rr = q_to_R(rq)           # result_delta_q is NOT yet normalised here
        I3 = np.eye(3)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * Rd @ R_a_0_x * dt * dt + -0.25 * Rr @ R_a_1_x @ (I3 - R_w_x * dt) * dt * dt
        F[0:3, 6:9] = I3 * dt
        F[0:3, 9:12] = -0.25 * (Rd + Rr) * dt * dt
        F[0:3, 12:15] = -0.25 * Rr @ R_a_1_x * dt * dt * -dt
        F[3:6, 3:6] = I3 - R_w_x * dt
        F[3:6, 12:15] = -1.0 * I3 * dt
        F[6:9, 3:6] = -0.5 * Rd @ R_a_0_x * dt + -0.5 * Rr @ R_a_1_x @ (I3 - R_w_x * dt) * dt
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (Rd + Rr) * dt
        F[6:9, 12:15] = -0.5 * Rr @ R_a_1_x * dt * -dt
        F[9:12, 9:12] = I3
        F[12:15, 12:15] = I3
        V = np.zeros((15, 18))
        V[0:3, 0:3] = 0.25 * Rd * dt * dt
        V[0:3, 3:6] = 0.25 * -Rr @ R_a_1_x * dt * dt * 0.5 * dt
        V[0:3, 6:9] = 0.25 * Rr * dt * dt
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I3 * dt
        V[3:6, 9:12] = 0.5 * I3 * dt
        V[6:9, 0:3] = 0.5 * Rd * dt
        V[6:9, 3:6] = 0.5 * -Rr @ R_a_1_x * dt * 0.5 * dt
        V[6:9, 6:9] = 0.5 * Rr * dt
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I3 * dt
        V[12:15, 15:18] = I3 * dt
        self.jacobian = F @ self.jacobian
        self.covariance = F @ self.covariance @ F.T + V @ self.noise @ V.T
        # :148-157
        self.delta_p = rp
        self.delta_q = q_normalized(rq)
        self.delta_v = rv
        self.sum_dt += dt
        self.acc_0 = _acc_1.copy()
        self.gyr_0 = _gyr_1.copy()

    def evaluate(self, Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj) -> np.ndarray:  # :160-186
        G = self.cfg.G
        J = self.jacobian
        dp_dba = J[O_P:O_P + 3, O_BA:O_BA + 3]
        dp_dbg = J[O_P:O_P + 3, O_BG:O_BG + 3]
        dq_dbg = J[O_R:O_R + 3, O_BG:O_BG + 3]
        dv_dba = J[O_V:O_V + 3, O_BA:O_BA + 3]
        dv_dbg = J[O_V:O_V + 3, O_BG:O_BG + 3]
        dba = Bai - self.linearized_ba
        dbg = Bgi - self.linearized_bg
        cq = q_mul(self.delta_q, deltaQ(dq_dbg @ dbg))
        cv = self.delta_v + dv_dba @ dba + dv_dbg @ dbg
        cp = self.delta_p + dp_dba @ dba + dp_dbg @ dbg
        s = self.sum_dt
        Qi_inv = q_inv(Qi)
        r = np.zeros(15)
        r[O_P:O_P + 3] = q_rot(Qi_inv, 0.5 * G * s * s + Pj - Pi - Vi * s) - cp
        r[O_R:O_R + 3] = 2.0 * q_mul(q_inv(cq), q_mul(Qi_inv, Qj))[1:4]
        r[O_V:O_V + 3] = q_rot(Qi_inv, G * s + Vj - Vi) - cv
        r[O_BA:O_BA + 3] = Baj - Bai
        r[O_BG:O_BG + 3] = Bgj - Bgi
        return r

    # --- flat state (the C-ABI 'pre-integration' record) ---------------------------------
    def pack(self) -> np.ndarray:
        """[delta_p3, delta_q(xyzw)4, delta_v3, lin_ba3, lin_bg3, sum_dt1, J 225 col-major, P 225 col-major] = 467"""
        q = self.delta_q
        return np.concatenate([
            self.delta_p, [q[1], q[2], q[3], q[0]], self.delta_v,
            self.linearized_ba, self.linearized_bg, [self.sum_dt],
            self.jacobian.flatten(order="F"), self.covariance.flatten(order="F")])


# ----------------------------------------------------------------------------------------------
# include/factor/imu_factor.h
# ----------------------------------------------------------------------------------------------
class IMUFactor:
    def __init__(self, pre_integration: IntegrationBase):
        self.pre_integration = pre_integration
        self.imu_i = -1
        self.imu_j = -1
        self.jacobians: List[np.ndarray] = [None] * 4
        self.sqrt_info = None
        self.residual = None

    def setIndex(self, i, j):
        self.imu_i, self.imu_j = i, j

    def _common(self, PSi, VBi, PSj, VBj):
        pre = self.pre_integration
        Pi, Qi = np.asarray(PSi[0:3], dtype=np.float64), quat_from_pose(PSi)
        Vi, Bai, Bgi = (np.asarray(VBi[0:3], dtype=np.float64), np.asarray(VBi[3:6], dtype=np.float64),
                        np.asarray(VBi[6:9], dtype=np.float64))
        Pj, Qj = np.asarray(PSj[0:3], dtype=np.float64), quat_from_pose(PSj)
        Vj, Baj, Bgj = (np.asarray(VBj[0:3], dtype=np.float64), np.asarray(VBj[3:6], dtype=np.float64),
                        np.asarray(VBj[6:9], dtype=np.float64))
        residual = pre.evaluate(Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj)
        # imu_factor.h:44 / :181 : LLT(cov.inverse()).matrixL().transpose()
        sqrt_info = llt_upper(partial_piv_inverse(pre.covariance))
        G = pre.cfg.G
        s = pre.sum_dt
        J = pre.jacobian
        dp_dba = J[O_P:O_P + 3, O_BA:O_BA + 3]
        dp_dbg = J[O_P:O_P + 3, O_BG:O_BG + 3]
        dq_dbg = J[O_R:O_R + 3, O_BG:O_BG + 3]
        dv_dba = J[O_V:O_V + 3, O_BA:O_BA + 3]
        dv_dbg = J[O_V:O_V + 3, O_BG:O_BG + 3]
        Qi_inv = q_inv(Qi)
        Ri_inv = q_to_R(Qi_inv)
        cq = q_mul(pre.delta_q, deltaQ(dq_dbg @ (Bgi - pre.linearized_bg)))
        j0 = np.zeros((15, 6))
        j0[O_P:O_P + 3, O_P:O_P + 3] = -Ri_inv
        j0[O_P:O_P + 3, O_R:O_R + 3] = skew(q_rot(Qi_inv, 0.5 * G * s * s + Pj - Pi - Vi * s))
        j0[O_R:O_R + 3, O_R:O_R + 3] = -(Qleft(q_mul(q_inv(Qj), Qi)) @ Qright(cq))[1:4, 1:4]
        j0[O_V:O_V + 3, O_R:O_R + 3] = skew(q_rot(Qi_inv, G * s + Vj - Vi))
        j1 = np.zeros((15, 9))
        j1[O_P:O_P + 3, 0:3] = -Ri_inv * s
        j1[O_P:O_P + 3, 3:6] = -dp_dba
        j1[O_P:O_P + 3, 6:9] = -dp_dbg
        # Q9: uses the *uncorrected* delta_q (imu_factor.h:105,227)
        j1[O_R:O_R + 3, 6:9] = -Qleft(q_mul(q_mul(q_inv(Qj), Qi), pre.delta_q))[1:4, 1:4] @ dq_dbg
        j1[O_V:O_V + 3, 0:3] = -Ri_inv
        j1[O_V:O_V + 3, 3:6] = -dv_dba
        j1[O_V:O_V + 3, 6:9] = -dv_dbg
        j1[O_BA:O_BA + 3, 3:6] = -np.eye(3)
        j1[O_BG:O_BG + 3, 6:9] = -np.eye(3)
        j2 = np.zeros((15, 6))
        j2[O_P:O_P + 3, O_P:O_P + 3] = Ri_inv
        j2[O_R:O_R + 3, O_R:O_R + 3] = Qleft(q_mul(q_mul(q_inv(cq), Qi_inv), Qj))[1:4, 1:4]
        j3 = np.zeros((15, 9))
        j3[O_V:O_V + 3, 0:3] = Ri_inv
        j3[O_BA:O_BA + 3, 3:6] = np.eye(3)
        j3[O_BG:O_BG + 3, 6:9] = np.eye(3)
        return residual, sqrt_info, [j0, j1, j2, j3]

    def Evaluate(self, PSi, VBi, PSj, VBj):
        """Tangent twin  imu_factor.h:161-265 : unweighted J (15x6,15x9,15x6,15x9), separate sqrt_info."""
        self.residual, self.sqrt_info, self.jacobians = self._common(PSi, VBi, PSj, VBj)

    def EvaluateCeres(self, parameters, want=(True, True, True, True)):
        """ceres contract  imu_factor.h:23-159 : weighted residual(15), row-major 15x7/15x9 (7th col 0)."""
        r, s, js = self._common(*parameters)
        res = s @ r
        out = []
        for k, (j, w) in enumerate(zip(js, want)):
            if not w:
                out.append(None)
                continue
            if k in (0, 2):
                full = np.zeros((15, 7))
                full[:, :6] = j
            else:
                full = j.copy()
            out.append(s @ full)
        return res, out


# ----------------------------------------------------------------------------------------------
# src/factor/projection_factor.cpp
# ----------------------------------------------------------------------------------------------
class ProjectionFactor:
    def __init__(self, pts_i, pts_j, sqrt_info):
        self.pts_i = np.array(pts_i, dtype=np.float64)
        self.pts_j = np.array(pts_j, dtype=np.float64)
        self.sqrt_info = np.array(sqrt_info, dtype=np.float64)   # static member in the reference
        self.imu_i = self.imu_j = self.feature_idx = -1
        self.jacobians: List[np.ndarray] = [None] * 4
        self.residual = None

    def setIndex(self, i, j, f):
        self.imu_i, self.imu_j, self.feature_idx = i, j, f

    def _common(self, PSi, PSj, PSic, inv_dep):
        Pi, Qi = np.asarray(PSi[0:3], dtype=np.float64), quat_from_pose(PSi)
        Pj, Qj = np.asarray(PSj[0:3], dtype=np.float64), quat_from_pose(PSj)
        tic, qic = np.asarray(PSic[0:3], dtype=np.float64), quat_from_pose(PSic)
        inv_dep_i = float(inv_dep)
        pts_camera_i = self.pts_i / inv_dep_i
        pts_imu_i = q_rot(qic, pts_camera_i) + tic
        pts_w = q_rot(Qi, pts_imu_i) + Pi
        pts_imu_j = q_rot(q_inv(Qj), pts_w - Pj)
        pts_camera_j = q_rot(q_inv(qic), pts_imu_j - tic)
        dep_j = pts_camera_j[2]
        residual = (pts_camera_j / dep_j)[0:2] - self.pts_j[0:2]
        Ri, Rj, ric = q_to_R(Qi), q_to_R(Qj), q_to_R(qic)
        reduce = np.array([[1.0 / dep_j, 0, -pts_camera_j[0] / (dep_j * dep_j)],
                           [0, 1.0 / dep_j, -pts_camera_j[1] / (dep_j * dep_j)]])
        jaco_i = np.zeros((3, 6))
        jaco_i[:, 0:3] = ric.T @ Rj.T
        jaco_i[:, 3:6] = ric.T @ Rj.T @ Ri @ -skew(pts_imu_i)
        jaco_j = np.zeros((3, 6))
        jaco_j[:, 0:3] = ric.T @ -Rj.T
        jaco_j[:, 3:6] = ric.T @ skew(pts_imu_j)
        jaco_ex = np.zeros((3, 6))
        jaco_ex[:, 0:3] = ric.T @ (Rj.T @ Ri - np.eye(3))
        tmp_r = ric.T @ Rj.T @ Ri @ ric
        jaco_ex[:, 3:6] = (-tmp_r @ skew(pts_camera_i) + skew(tmp_r @ pts_camera_i) +
                           skew(ric.T @ (Rj.T @ (Ri @ tic + Pi - Pj) - tic)))
        jfeat = ric.T @ Rj.T @ Ri @ ric @ self.pts_i * -1.0 / (inv_dep_i * inv_dep_i)
        return residual, reduce, jaco_i, jaco_j, jaco_ex, jfeat

    def EvaluateOnlyJacobians(self, PSi, PSj, PSic, inv_dep):
        """projection_factor.cpp:124-196 : unweighted 2x6,2x6,2x6,2x1."""
        r, reduce, ji, jj, jex, jf = self._common(PSi, PSj, PSic, inv_dep)
        self.residual = r
        self.jacobians = [reduce @ ji, reduce @ jj, reduce @ jex, (reduce @ jf).reshape(2, 1)]

    def EvaluateCeres(self, parameters, want=(True, True, True, True)):
        """projection_factor.cpp:24-122 : weighted, row-major 2x7,2x7,2x7,2x1."""
        r, reduce, ji, jj, jex, jf = self._common(parameters[0], parameters[1], parameters[2], parameters[3][0])
        res = self.sqrt_info @ r
        red = self.sqrt_info @ reduce
        out = []
        for k, (j, w) in enumerate(zip((ji, jj, jex), want[:3])):
            if not w:
                out.append(None)
                continue
            full = np.zeros((2, 7))
            full[:, :6] = red @ j
            out.append(full)
        out.append((red @ jf).reshape(2, 1) if want[3] else None)
        return res, out

    def EvaluateResidual(self, PSi, PSj, PSic, inv_dep) -> float:
        """projection_factor.h:15-39"""
        r = self._common(PSi, PSj, PSic, inv_dep)[0]
        return float(r @ r)


class ProjectionTdFactor(ProjectionFactor):
    """VINS-Mono vins_estimator/src/factor/projection_td_factor.cpp (ABSENT from /root/reference, SURVEY.md
    section 0; restated from the published algorithm for BASELINE configs[3] -- parity unpinned):
    observations are shifted by the time offset td along their image velocity,
        pts_td = pts - (td - td_obs + TR / ROW * row) * velocity ,  velocity.z = 0 ,
    then the ProjectionFactor chain; 5th parameter block td (1) with
        jacobian_td = reduce * ric^T Rj^T Ri ric * velocity_i / inv_dep * -1 + sqrt_info * velocity_j.head(2)."""

    def __init__(self, pts_i, pts_j, velocity_i, velocity_j, td_i, td_j, row_i, row_j, sqrt_info, tr_over_row=0.0):
        super().__init__(pts_i, pts_j, sqrt_info)
        self.obs_i = np.array(pts_i, dtype=np.float64)
        self.obs_j = np.array(pts_j, dtype=np.float64)
        self.velocity_i = np.array([velocity_i[0], velocity_i[1], 0.0])
        self.velocity_j = np.array([velocity_j[0], velocity_j[1], 0.0])
        self.td_i, self.td_j, self.row_i, self.row_j = float(td_i), float(td_j), float(row_i), float(row_j)
        self.tr_over_row = float(tr_over_row)

    def EvaluateCeres(self, parameters, want=(True, True, True, True, True)):
        td = float(parameters[4][0])
        self.pts_i = self.obs_i - (td - self.td_i + self.tr_over_row * self.row_i) * self.velocity_i
        self.pts_j = self.obs_j - (td - self.td_j + self.tr_over_row * self.row_j) * self.velocity_j
        r, reduce, ji, jj, jex, jf = self._common(parameters[0], parameters[1], parameters[2], parameters[3][0])
        res = self.sqrt_info @ r
        red = self.sqrt_info @ reduce
        out = []
        for j, w in zip((ji, jj, jex), want[:3]):
            if not w:
                out.append(None)
                continue
            full = np.zeros((2, 7))
            full[:, :6] = red @ j
            out.append(full)
        out.append((red @ jf).reshape(2, 1) if want[3] else None)
        if want[4]:
            Ri, Rj = q_to_R(quat_from_pose(parameters[0])), q_to_R(quat_from_pose(parameters[1]))
            ric = q_to_R(quat_from_pose(parameters[2]))
            inv_dep = float(parameters[3][0])
            jtd = red @ ric.T @ Rj.T @ Ri @ ric @ self.velocity_i / inv_dep * -1.0 + self.sqrt_info @ self.velocity_j[0:2]
            out.append(jtd.reshape(2, 1))
        else:
            out.append(None)
        return res, out


# ----------------------------------------------------------------------------------------------
# include/factor/relative_pose_factor.h
# ----------------------------------------------------------------------------------------------
class RelativePoseFactor:
    def __init__(self, delta_t, delta_R):
        self.delta_t = np.array(delta_t, dtype=np.float64)
        self.delta_R = np.array(delta_R, dtype=np.float64)
        self.sqrt_info = np.zeros((0, 0))
        self.imu_i = self.imu_j = -1
        self.jacobians: List[np.ndarray] = [None, None]
        self.residual = np.zeros(6)

    def setIndex(self, i, j):
        self.imu_i, self.imu_j = i, j

    def shift(self):                                          # :126-129
        self.imu_i -= 1
        self.imu_j -= 1

    def _common(self, PSi, PSj):
        Pi, Qi = np.asarray(PSi[0:3], dtype=np.float64), quat_from_pose(PSi)
        Pj, Qj = np.asarray(PSj[0:3], dtype=np.float64), quat_from_pose(PSj)
        Rj, Ri = q_to_R(Qj), q_to_R(Qi)
        tij = q_rot(q_inv(Qi), Pj - Pi)
        res_t = self.delta_t - tij
        res_R = SO3.from_matrix(self.delta_R @ Rj.T @ Ri)
        lg = res_R.log()
        r = np.concatenate([res_t, lg])
        J = right_jacobian_inv_SO3(lg)
        ji = np.zeros((6, 6))
        ji[0:3, 0:3] = Ri.T
        ji[0:3, 3:6] = -skew(tij)
        ji[3:6, 3:6] = J
        jj = np.zeros((6, 6))
        jj[0:3, 0:3] = -Ri.T
        jj[3:6, 3:6] = -J @ Ri.T @ Rj
        return r, ji, jj

    def EvaluateOnlyJacobians(self, PSi, PSj):                # :72-101
        self.residual, ji, jj = self._common(PSi, PSj)
        self.jacobians = [ji, jj]

    def EvaluateCeres(self, parameters, want=(True, True)):   # :27-70
        r, ji, jj = self._common(parameters[0], parameters[1])
        res = self.sqrt_info @ r
        out = []
        for j, w in zip((ji, jj), want):
            if not w:
                out.append(None)
                continue
            full = np.zeros((6, 7))
            full[:, :6] = j
            out.append(self.sqrt_info @ full)
        return res, out

    def update(self, ti, Ri, tj, Rj, PSi, PSj):               # :103-117
        Pi, Qi = np.asarray(PSi[0:3], dtype=np.float64), quat_from_pose(PSi)
        Pj, Qj = np.asarray(PSj[0:3], dtype=np.float64), quat_from_pose(PSj)
        d_tj = Pj - tj
        d_ti = Pi - ti
        d_Rj = SO3.from_matrix(q_to_R(q_inv(Qj)) @ Rj)
        d_Ri = SO3.from_matrix(q_to_R(q_inv(Qi)) @ Ri)
        self.delta_t = self.delta_t + Ri.T @ d_tj - Ri.T @ d_ti + skew(self.delta_t) @ d_Ri.log()
        Ji = -q_to_R(q_mul(q_inv(Qj), Qi))
        self.delta_R = self.delta_R @ SO3.exp(Ji @ d_Ri.log()).matrix()
        self.delta_R = self.delta_R @ SO3.exp(d_Rj.log()).matrix()


# ----------------------------------------------------------------------------------------------
# include/factor/se3_prior_factor.h
# ----------------------------------------------------------------------------------------------
class SE3PriorFactor:
    def __init__(self, t_new, q_new=None, R_new=None):
        """ctor takes (Vector3d, Quaterniond); member R is a Matrix3d  (:12, :136)."""
        self.t = np.array(t_new, dtype=np.float64)
        self.R = q_to_R(q_new) if R_new is None else np.array(R_new, dtype=np.float64)
        self.sqrt_info = np.zeros((0, 0))
        self.index = -1
        self.jacobians: List[np.ndarray] = [None]
        self.residual = np.zeros(6)

    def setIndex(self, i):
        self.index = i

    def _common(self, PSi):
        Pi, Qi = np.asarray(PSi[0:3], dtype=np.float64), quat_from_pose(PSi)
        ri = SO3.from_quat(Qi)
        rp = SO3.from_matrix(self.R)
        res_r = rp.inverse() * ri
        lg = res_r.log()
        r = np.concatenate([Pi - self.t, lg])
        J = np.zeros((6, 6))
        J[0:3, 0:3] = np.eye(3)
        J[3:6, 3:6] = right_jacobian_inv_SO3(lg)
        return r, J

    def EvaluateOnlyJacobians(self, PSi):                     # :53-71
        self.residual, J = self._common(PSi)
        self.jacobians = [J]

    def EvaluateCeres(self, parameters, want=(True,)):        # :21-51
        r, J = self._common(parameters[0])
        res = self.sqrt_info @ r
        if not want[0]:
            return res, [None]
        full = np.zeros((6, 7))
        full[:, :6] = J
        return res, [self.sqrt_info @ full]

    def update(self, Pi, Ri, PSi):                            # :73-81
        T0 = np.asarray(Pi, dtype=np.float64)
        T1 = np.asarray(PSi[0:3], dtype=np.float64)
        Qi = quat_from_pose(PSi)
        R0, R1 = SO3.from_matrix(Ri), SO3.from_quat(Qi)
        delta_R = (R1.inverse() * R0).log()
        self.t = self.t + (T1 - T0)
        self.R = self.R @ SO3.exp(delta_R).matrix()


# ----------------------------------------------------------------------------------------------
# include/factor/linear9_factor.h
# ----------------------------------------------------------------------------------------------
class Linear9Factor:
    def __init__(self, VB):
        self.VB = np.array(VB, dtype=np.float64)
        self.sqrt_info = np.zeros((0, 0))
        self.index = -1
        self.jacobians: List[np.ndarray] = [None]
        self.residual = np.zeros(9)

    def setIndex(self, i):
        self.index = i

    def EvaluateOnlyJacobians(self, VBi):                     # :46-59
        self.residual = np.asarray(VBi[0:9], dtype=np.float64) - self.VB
        self.jacobians = [np.eye(9)]

    def EvaluateCeres(self, parameters, want=(True,)):        # :20-44
        r = np.asarray(parameters[0][0:9], dtype=np.float64) - self.VB
        res = self.sqrt_info @ r
        return res, [self.sqrt_info @ np.eye(9) if want[0] else None]

    def update(self, Vi, Bai, Bgi, speedBias):                # :61-69
        VB0 = np.concatenate([Vi, Bai, Bgi])
        VB1 = np.asarray(speedBias[0:9], dtype=np.float64)
        self.VB = self.VB + (VB1 - VB0)


# ----------------------------------------------------------------------------------------------
# include/factor/rollpitch_factor.h , yaw_factor.h
# ----------------------------------------------------------------------------------------------
class RollPitchFactor:
    def __init__(self, q=None, R=None):
        """ctor takes a Quaterniond; member R is Matrix3d (:14, :131)."""
        self.R = q_to_R(q) if R is None else np.array(R, dtype=np.float64)
        self.sqrt_info = np.zeros((0, 0))
        self.index = -1
        self.jacobians: List[np.ndarray] = [None]
        self.residual = np.zeros(2)

    def setIndex(self, i):
        self.index = i

    def shift(self):
        self.index -= 1

    def _common(self, PSi):
        Qi = quat_from_pose(PSi)
        Ri = SO3.from_quat(Qi)
        Rmeas = SO3.from_matrix(self.R)
        nZ = -1.0 * np.array([0.0, 0.0, 1.0])
        res = Rmeas * (Ri.inverse() * nZ)
        J = np.zeros((3, 6))
        J[:, 3:6] = skew(res) @ Rmeas.matrix()
        return res[0:2], J[0:2, :]

    def EvaluateOnlyJacobians(self, PSi):                     # :59-76
        self.residual, J = self._common(PSi)
        self.jacobians = [J]

    def EvaluateCeres(self, parameters, want=(True,)):        # :26-57
        r, J = self._common(parameters[0])
        res = self.sqrt_info @ r
        if not want[0]:
            return res, [None]
        full = np.zeros((2, 7))
        full[:, :6] = J
        return res, [self.sqrt_info @ full]

    def update(self, Rs, Qs):                                 # :78-83
        Qi = quat_from_pose(Qs)
        R0, R1 = SO3.from_matrix(Rs), SO3.from_quat(Qi)
        delta_R = (R1.inverse() * R0).log()
        self.R = self.R @ SO3.exp(delta_R).matrix()


class YawFactor:
    def __init__(self, q):
        self.yaw_meas = q_rot(q_inv(q), np.array([1.0, 0.0, 0.0]))   # yaw_factor.h:17
        self.sqrt_info = np.zeros((0, 0))
        self.index = -1
        self.jacobians: List[np.ndarray] = [None]
        self.residual = np.zeros(1)

    def _common(self, PSi):
        Qi = quat_from_pose(PSi)
        Ri = SO3.from_quat(Qi)
        res = Ri * self.yaw_meas
        J = np.zeros((3, 6))
        J[:, 3:6] = -Ri.matrix() @ skew(self.yaw_meas)
        return res[1:2], J[1:2, :]

    def EvaluateOnlyJacobians(self, PSi):                     # :51-65
        self.residual, J = self._common(PSi)
        self.jacobians = [J]

    def EvaluateCeres(self, parameters, want=(True,)):        # :23-49
        r, J = self._common(parameters[0])
        res = self.sqrt_info @ r
        if not want[0]:
            return res, [None]
        full = np.zeros((1, 7))
        full[:, :6] = J
        return res, [self.sqrt_info @ full]


# ----------------------------------------------------------------------------------------------
# src/factor/pose_local_parameterization.cpp
# ----------------------------------------------------------------------------------------------
def pose_plus(x, delta) -> np.ndarray:
    """PoseLocalParameterization::Plus  :3-19"""
    p = np.asarray(x[0:3], dtype=np.float64) + np.asarray(delta[0:3], dtype=np.float64)
    q = q_normalized(q_mul(quat_from_pose(x), deltaQ(delta[3:6])))
    return np.array([p[0], p[1], p[2], q[1], q[2], q[3], q[0]])


def pose_compute_jacobian() -> np.ndarray:
    """PoseLocalParameterization::ComputeJacobian :20-27 : [I6; 0] (7x6 row-major)."""
    j = np.zeros((7, 6))
    j[:6, :6] = np.eye(6)
    return j


# ----------------------------------------------------------------------------------------------
# include/factor/pose_graph_factors.h
# ----------------------------------------------------------------------------------------------
@dataclass
class CombinedFactors:
    relativePoseFactor: RelativePoseFactor = None
    rollPitchFactor: Optional[RollPitchFactor] = None
    vio_index: int = -1
    length: int = 0
    pg_index: int = 0
    covRel: np.ndarray = field(default_factory=lambda: np.zeros((6, 6)))
    covAbs: Optional[np.ndarray] = None
    distance: float = 0.0
    ts: float = 0.0
    Ri: np.ndarray = field(default_factory=lambda: np.eye(3))
    ti: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def __post_init__(self):
        if self.relativePoseFactor is None:                   # :19-25
            self.relativePoseFactor = RelativePoseFactor(np.zeros(3), np.eye(3))

    def __add__(self, other: "CombinedFactors") -> "CombinedFactors":   # :27-51
        R0, t0 = self.relativePoseFactor.delta_R, self.relativePoseFactor.delta_t
        R1, t1 = other.relativePoseFactor.delta_R, other.relativePoseFactor.delta_t
        s1 = other.relativePoseFactor.sqrt_info
        covRel1 = partial_piv_inverse(s1.T @ s1)
        Adj = se3_adj(R0, t0)
        self.covRel = self.covRel + Adj @ covRel1 @ Adj.T
        self.rollPitchFactor = other.rollPitchFactor
        # T0*T1 with Sophus SE3 (rotation stored as unit quaternion)
        q0, q1 = R_to_q(R0), R_to_q(R1)
        q01 = (SO3(q0) * SO3(q1)).q
        t01 = q_rot(q0, t1) + t0
        self.relativePoseFactor = RelativePoseFactor(t01, q_to_R(q01))
        self.relativePoseFactor.sqrt_info = llt_upper(partial_piv_inverse(self.covRel))
        self.distance = float(np.linalg.norm(t01))
        self.length += 1
        if self.vio_index == -1:
            self.ti = other.ti
            self.Ri = other.Ri
            self.vio_index = other.vio_index
            self.ts = other.ts
        return self


def double2vector_rot_diff(Rs0, pose0) -> np.ndarray:
    """Estimator::double2vector  src/estimator.cpp:520-546 : the yaw re-anchoring rotation `rot_diff`
    (Rs[0] = orientation of frame 0 before the solve, pose0 = para_Pose[0] after it).  The caller
    applies it to the prior factors as :549-550 does:  vioVBPrior->VB.tail<3>() (Q17: the gyro-bias
    slot) and vioPosePriorEdge->R are pre-multiplied by it."""
    Rs0 = np.asarray(Rs0, dtype=np.float64)
    R00 = q_to_R(quat_from_pose(pose0))
    origin_R0 = R2ypr(Rs0)
    origin_R00 = R2ypr(R00)
    y_diff = origin_R0[0] - origin_R00[0]
    rot_diff = ypr2R(np.array([y_diff, 0.0, 0.0]))
    if abs(abs(origin_R0[1]) - 90) < 1.0 or abs(abs(origin_R00[1]) - 90) < 1.0:
        rot_diff = Rs0 @ R00.T
    return rot_diff


# ----------------------------------------------------------------------------------------------
# Index maps ("OrderMap") -- the bit-exact contract
# ----------------------------------------------------------------------------------------------
def order_map_init(V: int):
    """src/estimator.cpp:747-758 : T0..T_{V-1}@6i, VB_{V-1}@6V, VB_0..VB_{V-2}@6V+9+9i.
    Returns dict name -> (offset, dim)."""
    m = {}
    idx = 0
    for i in range(V):
        m[("pose", i)] = (idx, 6)
        idx += 6
    m[("sb", V - 1)] = (idx, 9)
    idx += 9
    for i in range(V - 1):
        m[("sb", i)] = (idx, 9)
        idx += 9
    return m


def order_map_forward(L: int):
    """src/estimator.cpp:1153-1162 : T1@0, T0@6, landmark k @12+k."""
    m = {}
    idx = 0
    for i in (1, 0):
        m[("pose", i)] = (idx, 6)
        idx += 6
    for k in range(L):
        m[("feat", k)] = (idx, 1)
        idx += 1
    return m


def order_map_backward(V: int):
    """src/estimator.cpp:1358-1366 : T_V@0, VB_V@6, T_{V-1}@15, VB_{V-1}@21."""
    m = {}
    idx = 0
    for i in (V, V - 1):
        m[("pose", i)] = (idx, 6)
        idx += 6
        m[("sb", i)] = (idx, 9)
        idx += 9
    return m


def _accumulate(Lamda, blocks, jacs, info):
    """The block loop  src/estimator.cpp:787-802 / 1183-1201 / 1222-1237 / 1396-1411.
    blocks: list of (offset, dim) or None (block not in OrderMap -> skipped)."""
    n = len(blocks)
    for j in range(n):
        if blocks[j] is None:
            continue
        oj, dj = blocks[j]
        JtW = jacs[j].T @ info
        for k in range(j, n):
            if blocks[k] is None:
                continue
            ok, dk = blocks[k]
            H = JtW @ jacs[k]
            Lamda[oj:oj + dj, ok:ok + dk] += H
            if j != k:
                Lamda[ok:ok + dk, oj:oj + dj] += H.T


def _truncated_eig(Lamda_prior, alpha):
    """src/estimator.cpp:920-940 / 1311-1331 / 1479-1497 : keep eigenvalues strictly > ALPHA."""
    w, Vec = self_adjoint_eig(Lamda_prior)
    vset = [i for i in range(w.size) if w[i] > alpha]
    U = Vec[:, vset]
    D = np.diag(w[vset])
    return U, D, len(vset), w


def _recover(Ji, U, Dinv):
    """per-factor info recovery :947-950 etc.:  cov=(JU)Dinv(JU)^T ; Omega=cov.inverse() ;
    sqrt_info = LLT(Omega).matrixL().transpose()."""
    JU = Ji @ U
    covi = JU @ Dinv @ JU.T
    Omega = partial_piv_inverse(covi)
    return covi, Omega, llt_upper(Omega)


# ----------------------------------------------------------------------------------------------
# MargForward   src/estimator.cpp:1149-1352
# ----------------------------------------------------------------------------------------------
@dataclass
class ForwardInput:
    pose0: np.ndarray            # para_Pose[0]  (7)
    pose1: np.ndarray            # para_Pose[1]  (7)
    ex_pose: np.ndarray          # para_Ex_Pose[0] (7)
    inv_dep: np.ndarray          # para_Feature[MargPointIdx[k]]  (L)
    pts_i: np.ndarray            # forwardProjectiontoSparsify[k]->pts_i  (L,3)
    pts_j: np.ndarray            # ...->pts_j (L,3)
    prior_t: np.ndarray          # vioPosePriorEdge->t (3)
    prior_R: np.ndarray          # vioPosePriorEdge->R (3x3)
    prior_sqrt_info: np.ndarray  # 6x6
    rel_dt: np.ndarray           # vioRelativePoseEdges[1]->delta_t
    rel_dR: np.ndarray           # ...->delta_R
    rel_sqrt_info: np.ndarray    # 6x6
    rp_valid: bool = False       # !vioRollPitchEdges.empty() && [0]->index==0
    rp_sqrt_info: Optional[np.ndarray] = None   # 2x2


@dataclass
class ForwardOutput:
    Lamda: np.ndarray            # (12+L)^2 (diagnostic)
    Lamda_prior: np.ndarray      # 6x6
    qr_rank: int
    used_eig_path: bool
    eig_rank: int
    se3_t: np.ndarray            # forwardPosePriorEdgeToAdd->t
    se3_R: np.ndarray
    se3_sqrt_info: np.ndarray
    pg_dt: np.ndarray            # CombinedFactors->relativePoseFactor->delta_t
    pg_dR: np.ndarray
    pg_sqrt_info: np.ndarray
    pg_covRel: np.ndarray
    pg_distance: float
    pg_covAbs: Optional[np.ndarray]
    kld: float = float("nan")    # :1333-1345 (diagnostic)


def marg_forward(inp: ForwardInput, cfg: Config, structured: bool = False) -> ForwardOutput:
    """Estimator::MargForward.  structured=False is the literal dense algorithm (full-pivot LU of
    the (L+6)^2 block); structured=True eliminates the diagonal landmark block first
    (mathematically identical) -- used only to cross-check the GPU design."""
    L = int(inp.inv_dep.shape[0])
    n = L + 12
    om = order_map_forward(L)
    Lamda = np.zeros((n, n))
    info_p = cfg.proj_sqrt_info.T @ cfg.proj_sqrt_info                      # :1173
    for k in range(L):                                                      # :1168-1202
        f = ProjectionFactor(inp.pts_i[k], inp.pts_j[k], cfg.proj_sqrt_info)
        f.EvaluateOnlyJacobians(inp.pose0, inp.pose1, inp.ex_pose, inp.inv_dep[k])
        blocks = [om[("pose", 0)], om[("pose", 1)], None, om[("feat", k)]]  # ex pose not in OrderMap
        _accumulate(Lamda, blocks, f.jacobians, info_p)
    prior = SE3PriorFactor(inp.prior_t, R_new=inp.prior_R)                  # :1203-1211
    prior.sqrt_info = inp.prior_sqrt_info
    prior.EvaluateOnlyJacobians(inp.pose0)
    o, d = om[("pose", 0)]
    Lamda[o:o + d, o:o + d] += prior.jacobians[0].T @ (prior.sqrt_info.T @ prior.sqrt_info) @ prior.jacobians[0]
    rel = RelativePoseFactor(inp.rel_dt, inp.rel_dR)                        # :1212-1238
    rel.sqrt_info = inp.rel_sqrt_info
    rel.EvaluateOnlyJacobians(inp.pose0, inp.pose1)
    _accumulate(Lamda, [om[("pose", 0)], om[("pose", 1)]], rel.jacobians, rel.sqrt_info.T @ rel.sqrt_info)

    Lamda_rr = Lamda[0:6, 0:6].copy()
    Lamda_mm = Lamda[6:, 6:].copy()
    Lamda_rp = Lamda[0:12, 0:12].copy()                                     # :1243
    Psi, Psj = np.asarray(inp.pose0[0:3], float), np.asarray(inp.pose1[0:3], float)
    Qi, Qj = quat_from_pose(inp.pose0), quat_from_pose(inp.pose1)
    tij = q_rot(q_inv(Qi), Psj - Psi)
    Rij = q_to_R(q_mul(q_inv(Qi), Qj))
    pg = RelativePoseFactor(tij, Rij)
    pg.EvaluateOnlyJacobians(inp.pose0, inp.pose1)
    J = np.zeros((6, 12))
    J[:, 0:6] = pg.jacobians[0]       # Q1: [J_T0 | J_T1] applied to Lamda ordered [T1,T0]
    J[:, 6:12] = pg.jacobians[1]
    Jpinv = pseudo_inverse(J, 1e-8)
    rpOmega = Jpinv.T @ Lamda_rp @ Jpinv
    rpCov = partial_piv_inverse(rpOmega)
    pg.sqrt_info = llt_upper(rpOmega)
    covAbs = None
    if inp.rp_valid:
        covAbs = partial_piv_inverse(inp.rp_sqrt_info.T @ inp.rp_sqrt_info)

    Lamda_rm = Lamda[0:6, 6:].copy()
    if not structured:
        Lamda_mm_inv = full_piv_lu_solve_identity(Lamda_mm)                 # :1286
        Lamda_prior = Lamda_rr - Lamda_rm @ Lamda_mm_inv @ Lamda_rm.T       # :1288
    else:
        dl = np.diag(Lamda)[12:].copy()
        Bp = Lamda[0:12, 12:]
        S = Lamda[0:12, 0:12] - (Bp / dl) @ Bp.T
        Lamda_prior = S[0:6, 0:6] - S[0:6, 6:12] @ np.linalg.solve(S[6:12, 6:12], S[6:12, 0:6])

    se3 = SE3PriorFactor(Psj, q_new=Qj)                                     # :1291-1297
    se3.EvaluateOnlyJacobians(inp.pose1)
    Jr = se3.jacobians[0]
    rank, solve = full_piv_householder_qr(Lamda_prior, ESTIMATOR_EPS)       # :1304-1305
    used_eig = False
    eig_rank = 6
    if rank == 6:
        cov = solve(np.eye(6))
        covi = Jr @ cov @ Jr.T
    else:
        used_eig = True
        U, D, eig_rank, _ = _truncated_eig(Lamda_prior, cfg.alpha)
        Dinv = partial_piv_inverse(D) if eig_rank > 0 else np.zeros((0, 0))
        covi = Jr @ U @ Dinv @ (Jr @ U).T
    X = partial_piv_inverse(covi)                                           # :1330-1332
    kld = float("nan")
    if rank == 6:                                                           # :1333-1345 (computed, never used: Q7)
        phi = Jr.T @ X @ Jr
        cov = solve(np.eye(6))
        with np.errstate(all="ignore"):
            kld = float(0.5 * (np.trace(phi @ cov) - np.log(np.linalg.det(phi)) - np.log(np.linalg.det(cov)) - 6))
    se3.sqrt_info = llt_upper(X)                                            # :1349
    return ForwardOutput(Lamda, Lamda_prior, rank, used_eig, eig_rank, se3.t, se3.R, se3.sqrt_info,
                         pg.delta_t, pg.delta_R, pg.sqrt_info, rpCov, float(np.linalg.norm(pg.delta_t)), covAbs, kld)


# ----------------------------------------------------------------------------------------------
# MargBackward   src/estimator.cpp:1354-1539
# ----------------------------------------------------------------------------------------------
@dataclass
class BackwardInput:
    pose_i: np.ndarray           # para_Pose[V-1]
    sb_i: np.ndarray             # para_SpeedBias[V-1]
    pose_j: np.ndarray           # para_Pose[V]
    sb_j: np.ndarray             # para_SpeedBias[V]
    vb_prior: np.ndarray         # vioVBPrior->VB (9)  (only enters the residual; J = I9)
    vb_sqrt_info: np.ndarray     # 9x9
    pre: IntegrationBase         # backwardIMUtoSparsify->pre_integration


@dataclass
class BackwardOutput:
    Lamda: np.ndarray            # 30x30
    Lamda_prior: np.ndarray      # 21x21
    eigvals: np.ndarray
    rank: int
    rel_dt: np.ndarray           # backwardRelativePoseEdgeToAdd
    rel_dR: np.ndarray
    rel_sqrt_info: np.ndarray    # 6x6
    vb: np.ndarray               # backwardVBEdgeToAdd->VB
    vb_sqrt_info: np.ndarray     # 9x9
    rp_R: np.ndarray             # rollPitchFactor->R
    rp_sqrt_info: np.ndarray     # 2x2
    abs_info: np.ndarray         # 3x3 (computed, discarded: Q6)
    yaw_info: np.ndarray         # 1x1 (computed, discarded: Q6)
    kld: float


def marg_backward(inp: BackwardInput, cfg: Config) -> BackwardOutput:
    V = cfg.vo_size
    om = order_map_backward(V)
    Lamda = np.zeros((30, 30))
    o, d = om[("sb", V - 1)]                                                # :1372-1380
    Lamda[o:o + d, o:o + d] += np.eye(9).T @ (inp.vb_sqrt_info.T @ inp.vb_sqrt_info) @ np.eye(9)
    f = IMUFactor(inp.pre)                                                  # :1382-1412
    f.Evaluate(inp.pose_i, inp.sb_i, inp.pose_j, inp.sb_j)
    omegaI = f.sqrt_info.T @ f.sqrt_info
    _accumulate(Lamda, [om[("pose", V - 1)], om[("sb", V - 1)], om[("pose", V)], om[("sb", V)]], f.jacobians, omegaI)
    Lamda_rr = Lamda[0:21, 0:21]
    Lamda_mm = Lamda[21:30, 21:30]
    Lamda_rm = Lamda[0:21, 21:30]
    Lamda_mm_inv = full_piv_lu_solve_identity(Lamda_mm)                     # :1417
    Lamda_prior = Lamda_rr - Lamda_rm @ Lamda_mm_inv @ Lamda_rm.T           # :1419

    Psi, Psj = np.asarray(inp.pose_i[0:3], float), np.asarray(inp.pose_j[0:3], float)
    Qi, Qj = quat_from_pose(inp.pose_i), quat_from_pose(inp.pose_j)
    tij = q_rot(q_inv(Qi), Psj - Psi)
    Rij = q_to_R(q_mul(q_inv(Qi), Qj))
    rel = RelativePoseFactor(tij, Rij)                                      # :1435-1436
    rel.EvaluateOnlyJacobians(inp.pose_i, inp.pose_j)
    vbf = Linear9Factor(np.asarray(inp.sb_j[0:9], float))                   # :1439-1444
    vbf.EvaluateOnlyJacobians(inp.sb_j)
    rpf = RollPitchFactor(q=Qi)                                             # :1447-1449
    rpf.EvaluateOnlyJacobians(inp.pose_i)
    yf = YawFactor(Qi)                                                      # :1451-1452
    yf.EvaluateOnlyJacobians(inp.pose_i)
    Jr = np.zeros((21, 21))                                                 # :1456-1464
    Jr[0:6, 15:21] += rel.jacobians[0]
    Jr[0:6, 0:6] += rel.jacobians[1]
    Jr[6:15, 6:15] += vbf.jacobians[0]
    Jr[15:17, 15:21] += rpf.jacobians[0]
    Jr[17:20, 15:18] += np.eye(3)
    Jr[20:21, 15:21] += yf.jacobians[0]
    U, D, rank, w = _truncated_eig(Lamda_prior, cfg.alpha)                  # :1479-1497
    Dinv = partial_piv_inverse(D)
    infos = []
    _, O_rp, s_rp = _recover(Jr[0:6], U, Dinv)                              # :1500-1503
    infos.append((O_rp, 0))
    _, O_vb, s_vb = _recover(Jr[6:15], U, Dinv)                             # :1505-1508
    infos.append((O_vb, 6))
    _, O_gv, s_gv = _recover(Jr[15:17], U, Dinv)                            # :1510-1516
    infos.append((O_gv, 15))
    _, O_abs, _ = _recover(Jr[17:20], U, Dinv)                              # :1518
    infos.append((O_abs, 17))
    _, O_yaw, _ = _recover(Jr[20:21], U, Dinv)                              # :1519
    infos.append((O_yaw, 20))
    rel.sqrt_info, vbf.sqrt_info, rpf.sqrt_info = s_rp, s_vb, s_gv
    rpf.setIndex(V - 1)
    kld = _kld(Jr, U, D, Dinv, infos, 21, 21)                               # :1522-1534
    return BackwardOutput(Lamda, Lamda_prior, w, rank, rel.delta_t, rel.delta_R, s_rp, vbf.VB, s_vb,
                          rpf.R, s_gv, O_abs, O_yaw, kld)


def _kld(Jr, U, D, Dinv, infos, xdim, ndim) -> float:
    """KLD diagnostic (computed and never used by the reference: Q7)."""
    X = np.zeros((xdim, xdim))
    for Om, off in infos:
        X[off:off + Om.shape[0], off:off + Om.shape[0]] += Om
    A = (Jr @ U).T @ X @ Jr @ U
    with np.errstate(all="ignore"):
        a = np.trace(A @ Dinv)
        sa, la = np.linalg.slogdet(A)
        sd, ld = np.linalg.slogdet(Dinv)
        return float(0.5 * (a - la - ld - ndim))


# ----------------------------------------------------------------------------------------------
# initFactorGraph sparsification tail   src/estimator.cpp:745-1001
# ----------------------------------------------------------------------------------------------
@dataclass
class InitInput:
    poses: np.ndarray            # para_Pose[0..V-1]       (V,7)
    sbs: np.ndarray              # para_SpeedBias[0..V-1]  (V,9)
    pres: List[IntegrationBase]  # pre_integrations[1..V-1]  (V-1 of them; pres[i] links i -> i+1)


@dataclass
class InitOutput:
    Lamda: np.ndarray
    Lamda_prior: np.ndarray
    eigvals: np.ndarray
    rank: int
    rel_dt: np.ndarray           # (V-1,3)   vioRelativePoseEdges[1..V-1]
    rel_dR: np.ndarray           # (V-1,3,3)
    rel_sqrt_info: np.ndarray    # (V-1,6,6)
    se3_t: np.ndarray
    se3_R: np.ndarray
    se3_sqrt_info: np.ndarray
    vb: np.ndarray
    vb_sqrt_info: np.ndarray
    kld: float


def init_sparsify(inp: InitInput, cfg: Config) -> InitOutput:
    V = cfg.vo_size
    om = order_map_init(V)
    n = 15 * V
    Lamda = np.zeros((n, n))
    for i in range(V - 1):                                                  # :774-803
        j = i + 1
        f = IMUFactor(inp.pres[i])
        f.setIndex(i, j)
        f.Evaluate(inp.poses[i], inp.sbs[i], inp.poses[j], inp.sbs[j])
        omegaI = f.sqrt_info.T @ f.sqrt_info
        _accumulate(Lamda, [om[("pose", i)], om[("sb", i)], om[("pose", j)], om[("sb", j)]], f.jacobians, omegaI)
    r_dim = 6 * V + 9
    m_dim = 9 * (V - 1)
    Lamda_rr = Lamda[0:r_dim, 0:r_dim]
    Lamda_mm = Lamda[r_dim:, r_dim:]
    Lamda_rm = Lamda[0:r_dim, r_dim:]
    Lamda_mm_inv = full_piv_lu_solve_identity(Lamda_mm)                     # :814
    Lamda_prior = Lamda_rr - Lamda_rm @ Lamda_mm_inv @ Lamda_rm.T           # :816
    rels = []
    Jr = np.zeros((r_dim, r_dim))
    rows = 0
    for i in range(V - 1):                                                  # :822-837, :879-896
        j = i + 1
        Psi, Psj = np.asarray(inp.poses[i][0:3], float), np.asarray(inp.poses[j][0:3], float)
        Qi, Qj = quat_from_pose(inp.poses[i]), quat_from_pose(inp.poses[j])
        tij = q_rot(q_inv(Qi), Psj - Psi)
        Rij = q_to_R(q_mul(q_inv(Qi), Qj))
        rf = RelativePoseFactor(tij, Rij)
        rf.setIndex(i, j)
        rf.EvaluateOnlyJacobians(inp.poses[i], inp.poses[j])
        rels.append(rf)
        Jr[rows:rows + 6, om[("pose", i)][0]:om[("pose", i)][0] + 6] += rf.jacobians[0]
        Jr[rows:rows + 6, om[("pose", j)][0]:om[("pose", j)][0] + 6] += rf.jacobians[1]
        rows += 6
    se3 = SE3PriorFactor(np.asarray(inp.poses[0][0:3], float), q_new=quat_from_pose(inp.poses[0]))  # :840-847
    se3.EvaluateOnlyJacobians(inp.poses[0])
    se3.setIndex(0)
    Jr[rows:rows + 6, 0:6] += se3.jacobians[0]                              # :898-905
    rows += 6
    vbf = Linear9Factor(np.asarray(inp.sbs[V - 1][0:9], float))             # :850-858
    vbf.setIndex(V - 1)
    vbf.EvaluateOnlyJacobians(inp.sbs[V - 1])
    o = om[("sb", V - 1)][0]
    Jr[rows:rows + 9, o:o + 9] += vbf.jacobians[0]                          # :906-913
    rows += 9
    U, D, rank, w = _truncated_eig(Lamda_prior, cfg.alpha)                  # :920-940
    Dinv = partial_piv_inverse(D)
    infos = []
    hdim = 0
    for rf in rels:                                                         # :944-952
        _, Om, s = _recover(Jr[hdim:hdim + 6], U, Dinv)
        infos.append((Om, hdim))
        rf.sqrt_info = s
        hdim += 6
    _, Om, s = _recover(Jr[hdim:hdim + 6], U, Dinv)                         # :954-962
    infos.append((Om, hdim))
    se3.sqrt_info = s
    hdim += 6
    _, Om, s = _recover(Jr[hdim:hdim + 9], U, Dinv)                         # :964-972
    infos.append((Om, hdim))
    vbf.sqrt_info = s
    hdim += 9
    kld = _kld(Jr, U, D, Dinv, infos, hdim, r_dim)                          # :974-988
    return InitOutput(Lamda, Lamda_prior, w, rank,
                      np.array([r.delta_t for r in rels]), np.array([r.delta_R for r in rels]),
                      np.array([r.sqrt_info for r in rels]), se3.t, se3.R, se3.sqrt_info, vbf.VB, vbf.sqrt_info, kld)


# ----------------------------------------------------------------------------------------------
# VINS-Mono MarginalizationInfo::marginalize (HKUST-Aerial-Robotics/VINS-Mono,
# vins_estimator/src/factor/marginalization_factor.cpp) -- NOT under /root/reference: IS-VINS deleted
# the class.  Restated from the published algorithm as the checker of the generic
# facade (north_star's MarginalizationInfo API); parity unpinned.
# ----------------------------------------------------------------------------------------------
def vins_mono_marginalize(factors, pos: int, m: int, eps: float = 1e-8):
    """factors: list of (residual (k,), [(tangent position, J (k x local_size)), ...]) already evaluated
    and loss-corrected (ResidualBlockInfo::Evaluate).  Returns dict(A, b, A_red, b_red,
    linearized_jacobians, linearized_residuals, rank).  Literal: dense A, JOINT eigen-thresholded
    pseudo-inverse of Amm = (A_mm + A_mm^T)/2, eigen-decomposition of the reduced system."""
    A = np.zeros((pos, pos))
    b = np.zeros(pos)
    for r, blocks in factors:                               # ThreadsConstructA
        for i, (pi, Ji) in enumerate(blocks):
            for j, (pj, Jj) in enumerate(blocks):
                if j < i:
                    continue
                blk = Ji.T @ Jj
                A[pi:pi + Ji.shape[1], pj:pj + Jj.shape[1]] += blk
                if j != i:
                    A[pj:pj + Jj.shape[1], pi:pi + Ji.shape[1]] += blk.T
            b[pi:pi + Ji.shape[1]] += Ji.T @ r
    out = vins_mono_schur_eig(A, b, m, eps)
    out.update({"A": A, "b": b})
    return out


def vins_mono_schur_eig(A, b, m: int, eps: float = 1e-8):
    """The back half of VINS-Mono's MarginalizationInfo::marginalize on assembled normal equations: joint
    eigen-thresholded pseudo-inverse of Amm, Schur complement, eigen-decomposition of the reduced system."""
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    w, V = np.linalg.eigh(Amm)
    winv = np.where(w > eps, 1.0 / np.where(w > eps, w, 1.0), 0.0)
    Amm_inv = (V * winv) @ V.T
    bmm, Amr, Arm, Arr, brr = b[:m], A[:m, m:], A[m:, :m], A[m:, m:], b[m:]
    A_red = Arr - Arm @ Amm_inv @ Amr
    b_red = brr - Arm @ Amm_inv @ bmm
    w2, V2 = np.linalg.eigh(0.5 * (A_red + A_red.T))
    S = np.where(w2 > eps, w2, 0.0)
    S_inv = np.where(w2 > eps, 1.0 / np.where(w2 > eps, w2, 1.0), 0.0)
    lin_J = np.sqrt(S)[:, None] * V2.T
    lin_r = np.sqrt(S_inv) * (V2.T @ b_red)
    return {"A_red": A_red, "b_red": b_red, "linearized_jacobians": lin_J,
            "linearized_residuals": lin_r, "rank": int(np.sum(w2 > eps)), "min_eig_Amm": float(w.min())}


class MarginalizationFactor:
    """VINS-Mono marginalization_factor.cpp `MarginalizationFactor::Evaluate` (published algorithm; IS-VINS
    deleted the class): the prior of the previous marginalization as a residual block.
    keep: list of (global size, idx = keep_block_idx - m); x0: list of the kept blocks at the linearization
    point (keep_block_data); lin_J (n x n), lin_r (n)."""

    def __init__(self, lin_J, lin_r, keep, x0):
        self.J, self.r0, self.keep = np.asarray(lin_J, float), np.asarray(lin_r, float), list(keep)
        self.x0 = [np.asarray(v, float).copy() for v in x0]

    def EvaluateCeres(self, params):
        n = self.J.shape[0]
        dx = np.zeros(n)
        for (size, idx), x, x0 in zip(self.keep, params, self.x0):
            x = np.asarray(x, float)
            if size != 7:
                dx[idx:idx + size] = x - x0
            else:
                dx[idx:idx + 3] = x[0:3] - x0[0:3]
                q0 = np.array([x0[6], x0[3], x0[4], x0[5]])          # (w, x, y, z)
                q = np.array([x[6], x[3], x[4], x[5]])
                d = q_mul(q_inv(q0), q)
                dx[idx + 3:idx + 6] = 2.0 * d[1:4] if d[0] >= 0 else 2.0 * -d[1:4]
        r = self.r0 + self.J @ dx
        js = []
        for size, idx in self.keep:
            local = 6 if size == 7 else size
            j = np.zeros((n, size))
            j[:, :local] = self.J[:, idx:idx + local]
            js.append(j)
        return r, js


def schur_complement_longdouble(A, b, m: int):
    """Adjudicator for the generic marginalization: A_rr - A_rm A_mm^-1 A_mr and b_r - A_rm A_mm^-1 b_m in
    80-bit extended precision (np.longdouble, eps ~ 1e-19) by a Cholesky solve; valid when A_mm is positive
    definite (then the eigen-thresholded pseudo-inverse is the inverse).  Returned as float64."""
    Al = np.asarray(A, dtype=np.longdouble)
    bl = np.asarray(b, dtype=np.longdouble)
    Amm = (Al[:m, :m] + Al[:m, :m].T) / np.longdouble(2)
    L = np.zeros((m, m), dtype=np.longdouble)
    for j in range(m):
        d = Amm[j, j] - L[j, :j] @ L[j, :j]
        L[j, j] = np.sqrt(d)
        if j + 1 < m:
            L[j + 1:, j] = (Amm[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    rhs = np.concatenate([Al[:m, m:], bl[:m, None]], axis=1)
    Y = np.zeros_like(rhs)
    for i in range(m):                       # L Y = rhs
        Y[i] = (rhs[i] - L[i, :i] @ Y[:i]) / L[i, i]
    S = Al[m:, m:] - Y[:, :-1].T @ Y[:, :-1]
    s = bl[m:] - Y[:, :-1].T @ Y[:, -1]
    return np.asarray(S, dtype=np.float64), np.asarray(s, dtype=np.float64)
