"""CPU tests of the oracle itself (no GPU): the reference ships no tests or golden vectors for
this path, so the oracle is pinned by (1) the reference's own finite-difference
recipes, (2) algebraic identities, (3) an independent C restatement, (4) committed fixtures."""
import os

import numpy as np
import pytest

from is_vins_b200.batch import pack_events
from oracle import isv_oracle as O
from oracle import ref_c, sim
from tests.helpers import compare_outputs, load_batch, oracle_outputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def chain():
    return sim.make_chain(sim.seed_for(1, 5), L=60, rounds=2)


def _fd(fun, x_pose_list, eps):
    """Forward difference over the tangent of a list of pose blocks (q <- q * deltaQ(d)), the recipe
    of ProjectionFactor::check (/root/reference/src/factor/projection_factor.cpp:250-298)."""
    r0 = fun(x_pose_list)
    cols = []
    for b, x in enumerate(x_pose_list):
        dim = 6 if len(x) == 7 else len(x)
        for k in range(dim):
            xs = [np.array(v, float) for v in x_pose_list]
            d = np.zeros(dim)
            d[k] = eps
            xs[b] = O.pose_plus(x, d) if len(x) == 7 else x + d
            cols.append((fun(xs) - r0) / eps)
    return np.stack(cols, axis=1)


def test_projection_jacobian_fd(chain):
    f_in = chain.events[0].fwd_in
    for k in (0, 7, 31):
        f = O.ProjectionFactor(f_in.pts_i[k], f_in.pts_j[k], chain.cfg.proj_sqrt_info)
        lam = np.array([f_in.inv_dep[k]])
        f.EvaluateOnlyJacobians(f_in.pose0, f_in.pose1, f_in.ex_pose, lam[0])
        J = np.concatenate(f.jacobians, axis=1)

        def res(xs):
            g = O.ProjectionFactor(f_in.pts_i[k], f_in.pts_j[k], chain.cfg.proj_sqrt_info)
            g.EvaluateOnlyJacobians(xs[0], xs[1], xs[2], xs[3][0])
            return g.residual
        Jfd = _fd(res, [f_in.pose0, f_in.pose1, f_in.ex_pose, lam], 1e-6)
        assert np.allclose(J, Jfd, rtol=0, atol=2e-5 * max(1.0, np.abs(J).max()))
        # ceres twin = sqrt_info * tangent twin, 7th column zero, row-major blocks
        r, Jc = f.EvaluateCeres([f_in.pose0, f_in.pose1, f_in.ex_pose, lam])
        assert np.allclose(r, chain.cfg.proj_sqrt_info @ f.residual)
        assert np.allclose(Jc[0][:, :6], chain.cfg.proj_sqrt_info @ f.jacobians[0]) and np.all(Jc[0][:, 6] == 0)


def test_recovered_factor_jacobians_fd(chain):
    ev = chain.events[1]
    f_in, b_in = ev.fwd_in, ev.bwd_in
    rel = O.RelativePoseFactor(f_in.rel_dt, f_in.rel_dR)
    rel.EvaluateOnlyJacobians(f_in.pose0, f_in.pose1)

    def res_rel(xs):
        g = O.RelativePoseFactor(f_in.rel_dt, f_in.rel_dR)
        g.EvaluateOnlyJacobians(xs[0], xs[1])
        return g.residual
    assert np.allclose(np.concatenate(rel.jacobians, 1), _fd(res_rel, [f_in.pose0, f_in.pose1], 1e-7), atol=1e-5)
    se3 = O.SE3PriorFactor(f_in.prior_t, R_new=f_in.prior_R)
    se3.EvaluateOnlyJacobians(f_in.pose0)

    def res_se3(xs):
        g = O.SE3PriorFactor(f_in.prior_t, R_new=f_in.prior_R)
        g.EvaluateOnlyJacobians(xs[0])
        return g.residual
    assert np.allclose(se3.jacobians[0], _fd(res_se3, [f_in.pose0], 1e-7), atol=1e-5)
    rp = O.RollPitchFactor(q=O.quat_from_pose(b_in.pose_i))
    rp.EvaluateOnlyJacobians(b_in.pose_j)

    def res_rp(xs):
        g = O.RollPitchFactor(q=O.quat_from_pose(b_in.pose_i))
        g.EvaluateOnlyJacobians(xs[0])
        return g.residual
    assert np.allclose(rp.jacobians[0], _fd(res_rp, [b_in.pose_j], 1e-7), atol=1e-5)


def test_imu_jacobian_fd(chain):
    b = chain.events[0].bwd_in
    f = O.IMUFactor(b.pre)
    f.Evaluate(b.pose_i, b.sb_i, b.pose_j, b.sb_j)
    J = np.concatenate(f.jacobians, axis=1)

    def res(xs):
        g = O.IMUFactor(b.pre)
        g.Evaluate(xs[0], xs[1], xs[2], xs[3])
        return g.residual
    Jfd = _fd(res, [b.pose_i, b.sb_i, b.pose_j, b.sb_j], 1e-7)
    # Q9 (uncorrected delta_q in d r_R / d bg) makes that one block a first-order approximation
    mask = np.ones_like(J, bool)
    mask[3:6, 12:15] = False
    assert np.allclose(J[mask], Jfd[mask], atol=5e-4 * max(1.0, np.abs(J).max()))
    # sqrt_info^T sqrt_info == covariance^-1
    assert np.allclose(f.sqrt_info.T @ f.sqrt_info @ b.pre.covariance, np.eye(15), atol=1e-6)


def test_preintegration_repropagate_is_deterministic(chain):
    b = chain.events[0].bwd_in
    pre = b.pre
    p0 = pre.pack().copy()
    pre.repropagate(pre.linearized_ba, pre.linearized_bg)
    assert np.array_equal(p0, pre.pack())
    assert abs(pre.sum_dt - 0.05) < 1e-12


def test_schur_is_block_of_inverse_and_structured_equals_literal(chain):
    for ev in chain.events:
        fo = ev.fwd_out
        cov = np.linalg.inv(fo.Lamda)
        assert np.allclose(np.linalg.inv(cov[0:6, 0:6]), fo.Lamda_prior, rtol=1e-7, atol=1e-7 * np.abs(fo.Lamda_prior).max())
        fs = O.marg_forward(ev.fwd_in, chain.cfg, structured=True)
        assert np.linalg.norm(fs.se3_sqrt_info - fo.se3_sqrt_info) <= 1e-12 * np.linalg.norm(fo.se3_sqrt_info)
        bo = ev.bwd_out
        covb = np.linalg.pinv(bo.Lamda, rcond=1e-14)
        assert bo.rank == 15 and fo.qr_rank == 6 and not fo.used_eig_path


def test_marginal_preservation_and_kld(chain):
    """per-factor marginal preservation Omega_i^-1 = J_i (U D^-1 U^T) J_i^T and KLD >= 0."""
    for ev in chain.events:
        bo = ev.bwd_out
        U, D, rank, w = O._truncated_eig(bo.Lamda_prior, chain.cfg.alpha)
        Sigma = U @ np.linalg.inv(D) @ U.T
        omega = bo.vb_sqrt_info.T @ bo.vb_sqrt_info
        assert np.allclose(np.linalg.inv(omega), Sigma[6:15, 6:15], rtol=1e-6, atol=1e-12)
        # the reference subtracts the full dimension (21) although A, D are rank x rank (:1532): a
        # harmless quirk of an unused diagnostic (Q7); with the true dimension the KLD is >= 0
        assert bo.kld + 0.5 * (21 - bo.rank) >= -1e-6
        assert np.all(np.tril(bo.vb_sqrt_info, -1) == 0) and np.all(np.diag(bo.vb_sqrt_info) > 0)
    assert chain.init_out.rank == 42 and chain.init_out.kld + 0.5 * (57 - 42) >= -1e-6


def test_order_maps_bit_exact():
    assert O.order_map_forward(3) == {("pose", 1): (0, 6), ("pose", 0): (6, 6), ("feat", 0): (12, 1),
                                      ("feat", 1): (13, 1), ("feat", 2): (14, 1)}
    assert O.order_map_backward(8) == {("pose", 8): (0, 6), ("sb", 8): (6, 9), ("pose", 7): (15, 6), ("sb", 7): (21, 9)}
    m = O.order_map_init(8)
    assert m[("pose", 0)] == (0, 6) and m[("pose", 7)] == (42, 6) and m[("sb", 7)] == (48, 9)
    assert m[("sb", 0)] == (57, 9) and m[("sb", 6)] == (111, 9)


@pytest.mark.parametrize("name", ["cfg1_L150_ragged.npz", "cfg2_L1000_literal.npz"])
def test_oracle_reproduces_golden_fixtures(name):
    """Regression anchor: re-running the generator's seeds reproduces the committed vectors, and
    the independent C restatement (oracle/isv_ref.c) agrees with them."""
    batch, ref, z = load_batch(os.path.join(GOLD, name))
    if name.startswith("cfg1"):
        ev = sim.make_chain(int(z["seeds"][0]), L=150, rounds=3).events
        ev += sim.make_chain(int(z["seeds"][1]), L=[0, 1, 31, 32, 33, 80], rounds=6, max_gap=3).events
        again = oracle_outputs(ev)
        assert max(compare_outputs(again, ref).values()) <= 1e-12
        b2 = pack_events(ev)
        assert np.array_equal(b2.lm_obs, batch.lm_obs) and np.array_equal(b2.preint, batch.preint)
    out_c = ref_c.marg_window_batch(batch, 3, 0, structured=False)
    assert max(compare_outputs(out_c, ref).values()) <= 1e-9
    assert np.array_equal(out_c.rank, ref.rank)
    out_s = ref_c.marg_window_batch(batch, 3, 0, structured=True)
    assert max(compare_outputs(out_s, ref).values()) <= 1e-9


def test_bench_fixtures_are_clean():
    for L in (150, 1000):
        batch, ref, _ = load_batch(os.path.join(GOLD, f"bench_windows_L{L}.npz"))
        assert batch.n == 8 and np.all(np.diff(batch.lm_offset) == L)
        assert np.all(ref.rank == np.array([6, 15])) and np.all(np.isfinite(ref.se3))
        out_s = ref_c.marg_window_batch(batch, 3, 0, structured=True)
        assert max(compare_outputs(out_s, ref).values()) <= 1e-9


# ------------------------------------------------------------------------------------------------
# The oracle's Eigen stand-ins against INDEPENDENT implementations of the same algorithm class (LAPACK):
# VERDICT r1 "what's missing" 2.  Eigen itself is absent from the image, LAPACK (through SciPy) is not.
# ------------------------------------------------------------------------------------------------
def _spd(rng, n, cond):
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    return (q * np.logspace(0, np.log10(cond), n)) @ q.T


@pytest.mark.parametrize("n,cond", [(6, 1e2), (9, 1e6), (15, 1e8), (156, 1e5), (406, 1e7)])
def test_full_piv_lu_matches_lapack_complete_pivoting(n, cond):
    """FullPivLU + solve(Identity) (src/estimator.cpp:814,1286,1417) vs LAPACK dgetc2 / dgesc2 (LU with complete
    pivoting): same algorithm class written by someone else.  Same pivot sequence, same factors, same inverse."""
    import scipy.linalg.lapack as la
    rng = np.random.default_rng(n)
    A = _spd(rng, n, cond) + 1e-3 * rng.normal(size=(n, n))      # not exactly symmetric, like Lamda_mm after round-off
    X = O.full_piv_lu_solve_identity(A)
    lu, ipiv, jpiv, info = la.dgetc2(A)
    assert info == 0
    Y = np.empty((n, n))
    for c in range(n):
        e = np.zeros(n)
        e[c] = 1.0
        x, scale = la.dgesc2(lu, e, ipiv, jpiv)
        Y[:, c] = x / scale
    tol = 1e-13 * cond
    assert np.linalg.norm(X - Y) <= tol * np.linalg.norm(Y)
    assert np.linalg.norm(A @ X - np.eye(n)) <= 1e-14 * cond * n


def test_full_piv_lu_rank_deficient_zero_fills_like_eigen():
    """rank < n: Eigen's solve() drops the dependent unknowns (zero-fill).  Checked against the minimum-residual
    property on the range: A X A = A for the oracle's X when the right-hand sides lie in range(A)."""
    rng = np.random.default_rng(3)
    B = rng.normal(size=(8, 5))
    A = B @ B.T                                                   # rank 5
    X = O.full_piv_lu_solve_identity(A)
    assert np.linalg.matrix_rank(X) <= 5
    assert np.linalg.norm(A @ X @ A - A) <= 1e-9 * np.linalg.norm(A)


@pytest.mark.parametrize("n,rank", [(6, 6), (6, 3), (6, 5), (12, 12), (15, 9)])
def test_full_piv_householder_qr_matches_lapack_pivoted_qr(n, rank):
    """FullPivHouseholderQR with setThreshold(1e-16) (src/estimator.cpp:1304-1309) vs LAPACK dgeqp3 (column-pivoted
    Householder QR, scipy.linalg.qr(pivoting=True)): rank decision, |det| and the solve."""
    import scipy.linalg as sla
    rng = np.random.default_rng(10 * n + rank)
    B = rng.normal(size=(n, rank))
    A = B @ B.T if rank < n else _spd(rng, n, 1e6)
    r, solve = O.full_piv_householder_qr(A, 1e-16)
    Q, R, P = sla.qr(A, pivoting=True)
    d = np.abs(np.diag(R))
    if rank == n:
        assert r == n
        X = solve(np.eye(n))
        assert np.linalg.norm(X - np.linalg.inv(A)) <= 1e-9 * np.linalg.norm(X)
        Y = np.zeros((n, n))
        Y[P, :] = sla.solve_triangular(R, Q.T)
        assert np.linalg.norm(X - Y) <= 1e-9 * np.linalg.norm(Y)
    else:
        # exact rank-deficiency leaves pivots at rounding level: with the reference's 1e-16 threshold (Q5) Eigen counts
        # every pivot above 1e-16 * maxpivot, which is NOT the numerical rank; the early exit (corner <= eps * size *
        # biggest) is what stops it.  LAPACK's pivots show the same gap.
        assert int(np.sum(d > 1e-10 * d[0])) == rank
        assert rank <= r <= n
        r_tight, _ = O.full_piv_householder_qr(A, 1e-10)
        assert r_tight == rank


@pytest.mark.parametrize("n,cond", [(2, 10.0), (6, 1e4), (9, 1e8), (15, 1e10)])
def test_llt_upper_matches_lapack_potrf(n, cond):
    import scipy.linalg.lapack as la
    rng = np.random.default_rng(n)
    M = _spd(rng, n, cond)
    U = O.llt_upper(M)
    c, info = la.dpotrf(M, lower=0)
    assert info == 0
    assert np.linalg.norm(U - np.triu(c)) <= 1e-13 * cond ** 0.5 * np.linalg.norm(U)
    assert np.all(np.tril(U, -1) == 0) and np.all(np.diag(U) > 0)
    # reads only the lower triangle (Eigen LLT's contract): garbage above the diagonal must not matter
    M2 = M.copy()
    M2[np.triu_indices(n, 1)] = 123.0
    assert np.array_equal(O.llt_upper(M2), U)
    # not SPD -> NaN (Eigen: NumericalIssue + garbage); the CUDA path raises ISV_W_NOT_SPD
    M3 = M.copy()
    M3[n - 1, n - 1] = -1.0
    assert np.isnan(O.llt_upper(M3)).any()
