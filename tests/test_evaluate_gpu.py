"""-m gpu: batched ceres-`Evaluate` kernels (isv_eval_*_batch) vs the oracle's EvaluateCeres for every
factor type problemSolve() owns (src/estimator.cpp:1004-1146).  Tolerance: 1e-9 relative per block
(north_star); the row-major 2x7 / 6x7 / 15x7 layout with a zero 7th column is checked exactly."""
import numpy as np
import pytest

from is_vins_b200 import DeviceProblem, FactorProblem, capi, eval_problem
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
