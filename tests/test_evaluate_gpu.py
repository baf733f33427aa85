"""-m gpu: batched ceres-`Evaluate` kernels (isv_eval_*_batch) vs the oracle's EvaluateCeres for every
factor type problemSolve() owns (src/estimator.cpp:1004-1146).  Tolerance: 1e-9 relative per block
(north_star); the row-major 2x7 / 6x7 / 15x7 layout with a zero 7th column is checked exactly."""
import numpy as np
import pytest

from is_vins_b200 import DeviceProblem, FactorProblem, capi, eval_problem
from oracle import isv_oracle as O
from oracle import sim
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _check(host, ref, n_copies=1):
    worst = 0.0
    P = len(ref["proj"])
    for c in range(n_copies):
        for k, (r, js) in enumerate(ref["proj"]):
            kk = c * P + k
            worst = max(worst, rel_err(host["proj_res"][kk], r))
            for name, j in zip(("proj_ji", "proj_jj", "proj_je", "proj_jf"), js):
                if name in host:
                    worst = max(worst, rel_err(host[name][kk], np.asarray(j).ravel()))
                    if name != "proj_jf":
                        assert host[name][kk][6] == 0.0 and host[name][kk][13] == 0.0
    for k, (r, js) in enumerate(ref["imu"]):
        worst = max(worst, rel_err(host["imu_res"][k], r))
        rec = host["imu_jac"][k]
        for off, w, j in ((0, 7, js[0]), (105, 9, js[1]), (240, 7, js[2]), (345, 9, js[3])):
            worst = max(worst, rel_err(rec[off:off + 15 * w], j.ravel()))
    for name, width in (("rel", 84), ("se3", 42), ("vb", 81), ("rp", 14), ("yaw", 7)):
        for k, (r, js) in enumerate(ref[name]):
            worst = max(worst, rel_err(host[name + "_res"][k], r))
            flat = np.concatenate([j.ravel() for j in js])
            assert flat.size == width
            worst = max(worst, rel_err(host[name + "_jac"][k], flat))
    return worst


@pytest.mark.parametrize("cauchy_a", [0.0, 1.0])
def test_evaluate_matches_oracle(backend, cauchy_a):
    p = sim.make_problem(sim.seed_for(6, 0), n_features=70)
    fp = FactorProblem.from_factors(p)
    dp = DeviceProblem(fp, "cuda:0")
    eval_problem(backend, dp, cauchy_a)
    backend.synchronize()
    assert int(dp.status.item()) == 0
    ref = sim.eval_problem_oracle(p, cauchy_a)
    worst = _check(dp.host(), ref)
    assert worst <= TOL, worst


def test_evaluate_ragged_tiled_and_null_blocks(backend):
    """P not a multiple of 32 or 128, three concatenated windows (index offsets), ex-pose block
    constant (jacobians[2] == nullptr) -> must not be written and the rest must be unchanged."""
    p = sim.make_problem(sim.seed_for(6, 1), n_features=37, max_track=5)
    fp = FactorProblem.from_factors(p).tile(3)
    dp = DeviceProblem(fp, "cuda:0", want_ex_jac=False)
    eval_problem(backend, dp, fused_call=False)   # the three per-class entry points
    backend.synchronize()
    assert int(dp.status.item()) == 0
    host = dp.host()
    ref = sim.eval_problem_oracle(p)
    assert "proj_je" not in host
    assert _check(host, ref, n_copies=3) <= TOL
    # residual-only call (ceres' jacobians == nullptr)
    dq = DeviceProblem(fp, "cuda:0", want_jac=False)
    eval_problem(backend, dq)
    backend.synchronize()
    hq = dq.host()
    for k in ("proj_res", "imu_res", "rel_res", "se3_res", "vb_res", "rp_res", "yaw_res"):
        assert np.array_equal(hq[k], host[k])


def test_evaluate_flags_bad_index(backend):
    p = sim.make_problem(sim.seed_for(6, 2), n_features=10)
    fp = FactorProblem.from_factors(p)
    fp.proj_idx[3, 5] = len(fp.feature) + 3
    dp = DeviceProblem(fp, "cuda:0")
    eval_problem(backend, dp)
    backend.synchronize()
    assert int(dp.status.item()) & capi.W_BAD_INDEX


def test_projection_td_factor_matches_oracle(backend):
    """ProjectionTdFactor through isv_eval_projection_batch (td_obs != NULL): 5 blocks incl. jac_td."""
    import ctypes as C
    import torch
    from oracle import isv_oracle as O
    p = sim.make_problem(sim.seed_for(6, 12), n_features=45)
    rng = np.random.default_rng(12)
    P = p.proj_idx.shape[1]
    tdo = np.zeros((8, P))
    tdo[0:4] = rng.normal(0, 0.4, (4, P))
    tdo[4:6] = rng.normal(0, 0.005, (2, P))
    tdo[6:8] = rng.uniform(-240, 240, (2, P))
    td = np.array([0.013, -0.02])
    td_idx = rng.integers(0, 2, P).astype(np.int32)
    tr = 0.033 / 480
    dev = "cuda:0"
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
    d = {"pose": t(p.poses), "sb": t(p.sbs), "ex": t(p.ex), "feat": t(p.feat), "idx": t(p.proj_idx, torch.int32),
         "obs": t(p.proj_obs), "tdo": t(tdo), "td": t(td), "tdi": t(td_idx, torch.int32)}
    o = {k: torch.zeros((P, w), dtype=torch.float64, device=dev) for k, w in
         (("res", 2), ("ji", 14), ("jj", 14), ("je", 14), ("jf", 2), ("jt", 2))}
    st = torch.zeros((1,), dtype=torch.int32, device=dev)
    pb = capi.isv_param_blocks(len(p.poses), len(p.sbs), 1, len(p.feat), d["pose"].data_ptr(), d["sb"].data_ptr(),
                               d["ex"].data_ptr(), d["feat"].data_ptr())
    pf = capi.isv_proj_factors(P, P, d["idx"].data_ptr(), d["obs"].data_ptr(), 1.0, d["tdo"].data_ptr(),
                               d["td"].data_ptr(), d["tdi"].data_ptr(), 2, 0, tr)
    po = capi.isv_proj_eval(*[o[k].data_ptr() for k in ("res", "ji", "jj", "je", "jf", "jt")])
    capi.check(backend.lib.isv_eval_projection_batch(backend.h, C.byref(pb), C.byref(pf), C.byref(po),
                                                     C.c_void_p(st.data_ptr())), "isv_eval_projection_batch")
    backend.synchronize()
    assert int(st.item()) == 0
    h = {k: v.cpu().numpy() for k, v in o.items()}
    worst = 0.0
    for k in range(P):
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        fac = O.ProjectionTdFactor(p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0]), tdo[0:2, k],
                                   tdo[2:4, k], tdo[4, k], tdo[5, k], tdo[6, k], tdo[7, k], p.cfg.proj_sqrt_info, tr)
        r, js = fac.EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1], td[td_idx[k]:td_idx[k] + 1]])
        r, js = sim.cauchy_correct(r, js, 1.0)
        worst = max(worst, rel_err(h["res"][k], r))
        for name, jb in zip(("ji", "jj", "je", "jf", "jt"), js):
            worst = max(worst, rel_err(h[name][k], np.asarray(jb).ravel()))
    assert worst <= TOL, worst


@pytest.mark.parametrize("name", ["problem_F60.npz", "problem_F300_host0.npz"])
def test_committed_problem_fixtures(backend, name):
    """tests/golden/problem_*.npz: the oracle's Evaluate outputs committed with the inputs."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    z = np.load(path)
    dp = DeviceProblem(FactorProblem.load(path), "cuda:0")
    eval_problem(backend, dp, 1.0)
    backend.synchronize()
    assert int(dp.status.item()) == 0
    h = dp.host()
    n = z["ref_proj_res"].shape[0]
    for k in ("proj_res", "proj_ji", "proj_jj", "proj_je", "proj_jf"):
        assert rel_err(h[k][:n], z["ref_" + k]) <= TOL, k
    # the IMU factor carries no loss function (problemSolve passes NULL): cauchy_a must not touch it
    assert rel_err(h["imu_res"], z["ref_imu_res"]) <= TOL
    assert rel_err(h["imu_jac"], z["ref_imu_jac"]) <= TOL


def test_pose_local_parameterization_plus_matches_oracle(backend):
    """PoseLocalParameterization::Plus (src/factor/pose_local_parameterization.cpp:3-19), batched: p + dp,
    normalize(q * deltaQ(dtheta)) with the unnormalised small-angle deltaQ of utility.h:11-24 -- including large steps
    (where deltaQ is far from a rotation) and in-place operation.  Tolerance 1e-15 (one normalisation)."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(11)
    n = 1000
    x = np.zeros((n, 7))
    x[:, 0:3] = rng.normal(0, 3.0, (n, 3))
    q = rng.normal(size=(n, 4))
    x[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    delta = rng.normal(0, 0.01, (n, 6))
    delta[::7] *= 100.0                                     # large steps
    delta[1] = 0.0                                          # zero step: x unchanged (up to the normalisation)
    ref = np.array([O.pose_plus(x[k], delta[k]) for k in range(n)])
    dx, dd = torch.from_numpy(x).cuda(), torch.from_numpy(delta).cuda()
    out = torch.empty_like(dx)
    capi.check(backend.lib.isv_pose_plus_batch(backend.h, n, C.c_void_p(dx.data_ptr()), C.c_void_p(dd.data_ptr()),
                                               C.c_void_p(out.data_ptr())), "isv_pose_plus_batch")
    backend.synchronize()
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-15 * max(1.0, np.abs(ref).max())
    assert np.abs(np.linalg.norm(got[:, 3:7], axis=1) - 1.0).max() <= 4e-16
    assert np.abs(got[1] - x[1]).max() <= 2e-16
    capi.check(backend.lib.isv_pose_plus_batch(backend.h, n, C.c_void_p(dx.data_ptr()), C.c_void_p(dd.data_ptr()),
                                               C.c_void_p(dx.data_ptr())), "isv_pose_plus_batch")   # in place
    backend.synchronize()
    assert np.array_equal(dx.cpu().numpy(), got)
