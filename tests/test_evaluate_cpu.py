"""CPU: the oracle's ceres-contract twins (the checker of tests/test_evaluate_gpu.py) against finite
differences taken through PoseLocalParameterization::Plus -- the recipe of the reference's own
`check()` functions (src/factor/projection_factor.cpp:197-299, relative_pose_factor.h:153-186) --
and the host mirror's packing / tiling logic."""
import numpy as np

from is_vins_b200 import FactorProblem, capi
from oracle import isv_oracle as O
from oracle import sim


def _fd_blocks(fn, params, tangent_dims, eps=1e-6):
    """d residual / d tangent for every block, central differences; pose blocks move through Plus()."""
    out = []
    for b, td in enumerate(tangent_dims):
        cols = []
        for k in range(td):
            d = np.zeros(td)
            d[k] = eps
            pp, pm = [np.array(x, float) for x in params], [np.array(x, float) for x in params]
            if td == 6:
                pp[b], pm[b] = O.pose_plus(params[b], d), O.pose_plus(params[b], -d)
            else:
                pp[b], pm[b] = pp[b] + d, pm[b] - d
            cols.append((fn(pp) - fn(pm)) / (2 * eps))
        out.append(np.array(cols).T)
    return out


def test_ceres_twins_match_finite_differences():
    p = sim.make_problem(sim.seed_for(6, 9), n_features=12)
    s = p.cfg.proj_sqrt_info
    k = 3
    i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
    pf = O.ProjectionFactor(p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0]), s)
    params = [p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1]]
    r, js = pf.EvaluateCeres(params)
    fd = _fd_blocks(lambda q: pf.EvaluateCeres(q, want=(False,) * 4)[0], params, (6, 6, 6, 1))
    for a, b in zip(js, fd):
        assert np.allclose(np.asarray(a)[:, :b.shape[1]], b, rtol=2e-5, atol=2e-4 * np.abs(b).max())
    imu = O.IMUFactor(p.imu_pre[4])
    params = [p.poses[4], p.sbs[4], p.poses[5], p.sbs[5]]
    r, js = imu.EvaluateCeres(params)
    fd = _fd_blocks(lambda q: imu.EvaluateCeres(q, want=(False,) * 4)[0], params, (6, 9, 6, 9))
    for bi, (a, b) in enumerate(zip(js, fd)):
        a = a[:, :b.shape[1]]
        if bi == 1:   # Q9: d r_R / d bg_i uses the uncorrected delta_q -- not the true derivative
            a, b = np.delete(a, slice(3, 6), 0), np.delete(b, slice(3, 6), 0)
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4 * np.abs(b).max())
    for fac, params, dims in ((p.rel[2], [p.poses[2], p.poses[3]], (6, 6)), (p.se3[0], [p.poses[0]], (6,)),
                              (p.vb[0], [p.sbs[p.vb[0].index]], (9,)), (p.rp[0], [p.poses[p.rp[0].index]], (6,)),
                              (p.yaw[0], [p.poses[p.yaw[0].index]], (6,))):
        r, js = fac.EvaluateCeres(params)
        fd = _fd_blocks(lambda q: fac.EvaluateCeres(q, want=(False,) * len(dims))[0], params, dims)
        for a, b in zip(js, fd):
            assert np.allclose(a[:, :b.shape[1]], b, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(b).max()))


def test_cauchy_corrector_is_sqrt_rho_prime():
    r = np.array([3.0, 4.0])
    rr, (j,) = sim.cauchy_correct(r, [np.eye(2)], 1.0)
    sc = 1.0 / np.sqrt(1.0 + 25.0)
    assert np.allclose(rr, r * sc) and np.allclose(j, np.eye(2) * sc)
    # 0.5 * rho(s) linearised: gradient J^T r of the corrected problem equals rho'(s) J^T r
    assert np.allclose(j.T @ rr, (1.0 / 26.0) * r)


def test_factor_problem_pack_and_tile():
    p = sim.make_problem(sim.seed_for(6, 3), n_features=9)
    fp = FactorProblem.from_factors(p)
    P = fp.proj_idx.shape[1]
    assert fp.proj_idx.dtype == np.int32 and fp.proj_obs.shape == (5, P)
    assert fp.imu_preint.shape == (p.cfg.all_buf_size - 1, capi.PREINT_REC)
    assert fp.rel_rec.shape == (p.cfg.vo_size - 1, capi.REL_REC) and fp.yaw_rec.shape == (1, capi.YAW_REC)
    # record layout: column-major matrices
    assert np.array_equal(fp.rel_rec[1, 3:12].reshape(3, 3).T, p.rel[1].delta_R)
    assert np.array_equal(fp.rel_rec[1, 12:].reshape(6, 6).T, p.rel[1].sqrt_info)
    t = fp.tile(3)
    assert t.proj_idx.shape == (4, 3 * P) and t.pose.shape[0] == 3 * fp.pose.shape[0]
    assert np.array_equal(t.proj_idx[:, 2 * P:] - t.proj_idx[:, :P],
                          np.array([[2 * 18], [2 * 18], [2], [2 * len(fp.feature)]]) * np.ones((1, P), np.int64))
    assert np.array_equal(t.imu_idx[-1], fp.imu_idx[-1] + 2 * 18)
    assert np.array_equal(t.vb_idx, fp.vb_idx[0] + 18 * np.arange(3))


def test_projection_td_twin_matches_finite_differences():
    """ProjectionTdFactor (VINS-Mono, absent from the reference): analytic blocks incl. d r / d td vs FD."""
    p = sim.make_problem(sim.seed_for(6, 11), n_features=6)
    rng = np.random.default_rng(5)
    k = 2
    i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
    pf = O.ProjectionTdFactor(p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0]),
                              rng.normal(0, 0.3, 2), rng.normal(0, 0.3, 2), 0.004, -0.003, 57.0, -120.0,
                              p.cfg.proj_sqrt_info, tr_over_row=0.03 / 480)
    params = [p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1], np.array([0.011])]
    r, js = pf.EvaluateCeres(params)
    fd = _fd_blocks(lambda q: pf.EvaluateCeres(q, want=(False,) * 5)[0], params, (6, 6, 6, 1, 1))
    for a, b in zip(js, fd):
        assert np.allclose(np.asarray(a)[:, :b.shape[1]], b, rtol=2e-5, atol=2e-4 * np.abs(b).max())


def test_oracle_marginalization_factor_fd_and_linearization_point():
    """VINS-Mono MarginalizationFactor restated (oracle/isv_oracle.py): residual = r0 at the linearization
    point, central-difference Jacobians in the PoseLocalParameterization tangent, q and -q are the same point."""
    from oracle import isv_oracle as O
    rng = np.random.default_rng(3)
    keep = [(7, 0), (9, 6), (7, 15), (1, 21)]
    n = 22
    J, r0 = rng.normal(size=(n, n)), rng.normal(size=n)
    q = lambda: (lambda v: v / np.linalg.norm(v))(rng.normal(size=4))
    x0 = [np.concatenate([rng.normal(size=3), q()]), rng.normal(size=9), np.concatenate([rng.normal(size=3), q()]),
          rng.normal(size=1)]
    f = O.MarginalizationFactor(J, r0, keep, x0)
    r, js = f.EvaluateCeres(x0)
    assert np.allclose(r, r0, atol=1e-14)
    x = [v.copy() for v in x0]
    x[1] += 0.01 * rng.normal(size=9)
    x[2][0:3] += 0.01
    r1, js = f.EvaluateCeres(x)
    xm = [v.copy() for v in x]
    xm[2][3:7] *= -1.0
    assert np.allclose(f.EvaluateCeres(xm)[0], r1, atol=1e-13)
    h = 1e-6
    for b, (size, idx) in enumerate(keep):
        local = 6 if size == 7 else size
        assert js[b].shape == (n, size) and (size != 7 or np.all(js[b][:, 6] == 0))
        for c in range(local):
            d = np.zeros(local)
            d[c] = h
            xp, xn = [v.copy() for v in x], [v.copy() for v in x]
            if size == 7:
                xp[b], xn[b] = O.pose_plus(x[b], d), O.pose_plus(x[b], -d)
            else:
                xp[b], xn[b] = x[b] + d, x[b] - d
            fd = (f.EvaluateCeres(xp)[0] - f.EvaluateCeres(xn)[0]) / (2 * h)
            assert np.allclose(fd, js[b][:, c], atol=1e-6), (b, c)
