"""-m gpu: the `MarginalizationInfo` facade (north_star API: addResidualBlockInfo / preMarginalize /
marginalize / getParameterBlocks) on the generic GPU engine vs the oracle's literal restatement of
VINS-Mono's MarginalizationInfo::marginalize (dense A, joint eigen pseudo-inverse of A_mm, eigh of the
reduced system).  The oldest frame's pose + speed-bias and every feature hosted in it are marginalized,
VINS-Mono style.  Parity unpinned by the reference (IS-VINS deleted the class).
Tolerance 1e-9 relative on the reduced system, on J^T J and on J^T r (row order / sign of
`linearized_jacobians` are conventions of the eigensolver, the prior cost |J dx + r|^2 is not)."""
import numpy as np
import pytest

from is_vins_b200 import MarginalizationInfo, ResidualBlockInfo
from is_vins_b200.marginalization import LOCAL_SIZE
from oracle import isv_oracle as O
from oracle import sim
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _build(backend, p, cauchy_a, ex_constant):
    const = [("ex_pose", 0)] if ex_constant else []
    mi = MarginalizationInfo(backend, eps=1e-8, cauchy_a=cauchy_a, constant=const)
    oracle_factors = []      # (kind-specific oracle evaluation, parameter keys)
    # IMU factor 0 -> 1, drop pose 0 and speed-bias 0
    keys = [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)]
    mi.addResidualBlockInfo(ResidualBlockInfo("imu", keys, drop_set=[0, 1], preint=p.imu_pre[0].pack()))
    r, js = O.IMUFactor(p.imu_pre[0]).EvaluateCeres([p.poses[0], p.sbs[0], p.poses[1], p.sbs[1]])
    oracle_factors.append((r, js, keys))
    # every projection factor hosted in frame 0, drop the host pose and the feature
    s = p.cfg.proj_sqrt_info
    for k in range(p.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        if i != 0:
            continue
        keys = [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)]
        pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
        mi.addResidualBlockInfo(ResidualBlockInfo("projection", keys, drop_set=[0, 3], pts_i=pts_i, pts_j=pts_j))
        r, js = O.ProjectionFactor(pts_i, pts_j, s).EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1]])
        oracle_factors.append(sim.cauchy_correct(r, js, cauchy_a) + (keys,))
    # the prior factors touching frame 0
    se3, rel = p.se3[0], p.rel[0]
    keys = [("pose", 0)]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", keys, drop_set=[0], t=se3.t, R=se3.R, sqrt_info=se3.sqrt_info))
    oracle_factors.append(sim.cauchy_correct(*se3.EvaluateCeres([p.poses[0]]), cauchy_a) + (keys,))
    keys = [("pose", 0), ("pose", 1)]
    mi.addResidualBlockInfo(ResidualBlockInfo("rel", keys, drop_set=[0], delta_t=rel.delta_t, delta_R=rel.delta_R,
                                              sqrt_info=rel.sqrt_info))
    oracle_factors.append(sim.cauchy_correct(*rel.EvaluateCeres([p.poses[0], p.poses[1]]), cauchy_a) + (keys,))
    return mi, oracle_factors, set(const)


@pytest.mark.parametrize("cauchy_a,ex_constant,nfeat", [(1.0, True, 120), (0.0, False, 61)])
def test_marginalization_info_matches_vins_mono_oracle(backend, cauchy_a, ex_constant, nfeat):
    p = sim.make_problem(sim.seed_for(9, int(ex_constant)), n_features=nfeat, max_track=9, host0=0.6)
    mi, ofac, const = _build(backend, p, cauchy_a, ex_constant)
    mi.preMarginalize({"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat})
    mi.marginalize()
    assert mi.status == 0, hex(mi.status)
    idx = mi.parameter_block_idx
    # block index mapping: dense marginalized blocks first, then the scalar ones, then the kept ones
    assert idx[("pose", 0)] == 0 and idx[("speed_bias", 0)] == 6
    n_diag = sum(1 for k in idx if k[0] == "feature")
    assert n_diag > 8 and mi.m == 15 + n_diag
    assert sorted(v for k, v in idx.items() if k[0] == "feature") == list(range(15, 15 + n_diag))
    assert min(v for k, v in idx.items() if v >= mi.m) == mi.m and mi.pos == mi.m + mi.n
    # the oracle on the same ordering
    facs = []
    for r, js, keys in ofac:
        blocks = [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k not in const]
        facs.append((r, blocks))
    ref = O.vins_mono_marginalize(facs, mi.pos, mi.m, eps=1e-8)
    assert ref["min_eig_Amm"] > 1e-8          # block-wise == joint pseudo-inverse
    # The information-form Schur complement loses cond(A_mm) * eps digits in ANY FP64 implementation
    # (VINS-Mono's included), so the literal FP64 oracle is not a 1e-9 yardstick here.  Adjudicate against
    # an 80-bit extended-precision Schur complement: the CUDA result must meet 1e-9 or be at least as
    # accurate as the literal FP64 algorithm.
    S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], mi.m)
    e_ref, e_gpu = rel_err(ref["A_red"], S_hp), rel_err(mi.A_red, S_hp)
    print(f"A_red vs 80-bit truth: literal FP64 oracle {e_ref:.2e}, CUDA {e_gpu:.2e}")
    assert e_gpu <= max(1e-9, 2.0 * e_ref), (e_gpu, e_ref)
    assert rel_err(mi.b_red, s_hp) <= max(1e-9, 2.0 * rel_err(ref["b_red"], s_hp))
    # `eigenvalue > eps` (eps = 1e-8) is decided by rounding noise for the directions the reduced system
    # does not constrain (noise ~ cond * eps_mach * lam_max >> eps): the rank is only defined above that floor
    lam_ref = np.linalg.eigvalsh(0.5 * (ref["A_red"] + ref["A_red"].T))
    floor = max(1e-8, 8.0 * max(e_ref, e_gpu) * lam_ref.max())
    lam_gpu = np.linalg.norm(mi.linearized_jacobians, axis=1) ** 2
    well = int(np.sum(lam_ref > floor))
    assert int(np.sum(lam_gpu > floor)) == well and well <= mi.rank <= mi.n and well <= ref["rank"]
    J, r = mi.linearized_jacobians, mi.linearized_residuals
    Jr, rr = ref["linearized_jacobians"], ref["linearized_residuals"]
    # the eigen stage itself, on the CUDA path's own reduced system: J^T J = A_red, J^T r = b_red (full rank)
    assert rel_err(J.T @ J, mi.A_red) <= 1e-9
    if mi.rank == mi.n:
        assert rel_err(J.T @ r, mi.b_red) <= 1e-9
    tol = max(1e-9, 4.0 * e_ref)
    assert rel_err(J.T @ J, Jr.T @ Jr) <= tol
    assert rel_err(J.T @ r, Jr.T @ rr) <= tol
    # the same eigenvalues in the same (ascending) row order; dropped rows are exactly zero
    assert np.allclose(np.sort(lam_gpu)[-well:], np.sort(lam_ref)[-well:], rtol=1e-6, atol=floor)
    assert np.all(np.diff(lam_gpu[lam_gpu > 0]) >= 0)           # ascending, zero rows first
    assert np.count_nonzero(lam_gpu) == mi.rank
    keep = mi.getParameterBlocks()
    assert keep[0][2] == 0 and all(b[2] > a[2] for a, b in zip(keep, keep[1:]))
    assert sum(LOCAL_SIZE[k[0][0]] for k in keep) == mi.n


def test_reduced_camera_system_of_a_whole_window(backend):
    """SURVEY 8f rank 1, second half: the reduced camera system DENSE_SCHUR forms inside problemSolve()
    (src/estimator.cpp:1004-1124) built on the GPU: every factor of the 18-frame window, every feature
    eliminated (diagonal block, DMMA Schur product), no dense block, no eigen-decomposition."""
    p = sim.make_problem(sim.seed_for(9, 7), n_features=90, max_track=8)
    const = {("ex_pose", 0)}
    mi = MarginalizationInfo(backend, eps=1e-8, cauchy_a=1.0, constant=list(const))
    ofac = []
    s = p.cfg.proj_sqrt_info
    for k in range(p.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        keys = [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)]
        pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
        mi.addResidualBlockInfo(ResidualBlockInfo("projection", keys, drop_set=[3], pts_i=pts_i, pts_j=pts_j))
        r, js = O.ProjectionFactor(pts_i, pts_j, s).EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1]])
        ofac.append(sim.cauchy_correct(r, js, 1.0) + (keys,))
    for k, (i, j) in enumerate(p.imu_idx):
        keys = [("pose", int(i)), ("speed_bias", int(i)), ("pose", int(j)), ("speed_bias", int(j))]
        mi.addResidualBlockInfo(ResidualBlockInfo("imu", keys, preint=p.imu_pre[k].pack()))
        ofac.append(O.IMUFactor(p.imu_pre[k]).EvaluateCeres([p.poses[i], p.sbs[i], p.poses[j], p.sbs[j]]) + (keys,))
    for rf in p.rel:
        keys = [("pose", rf.imu_i), ("pose", rf.imu_j)]
        mi.addResidualBlockInfo(ResidualBlockInfo("rel", keys, delta_t=rf.delta_t, delta_R=rf.delta_R, sqrt_info=rf.sqrt_info))
        ofac.append(sim.cauchy_correct(*rf.EvaluateCeres([p.poses[rf.imu_i], p.poses[rf.imu_j]]), 1.0) + (keys,))
    sf, vf = p.se3[0], p.vb[0]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", [("pose", sf.index)], t=sf.t, R=sf.R, sqrt_info=sf.sqrt_info))
    ofac.append(sim.cauchy_correct(*sf.EvaluateCeres([p.poses[sf.index]]), 1.0) + ([("pose", sf.index)],))
    mi.addResidualBlockInfo(ResidualBlockInfo("vb", [("speed_bias", vf.index)], VB=vf.VB, sqrt_info=vf.sqrt_info))
    ofac.append(sim.cauchy_correct(*vf.EvaluateCeres([p.sbs[vf.index]]), 1.0) + ([("speed_bias", vf.index)],))
    for rp in p.rp:
        mi.addResidualBlockInfo(ResidualBlockInfo("rp", [("pose", rp.index)], R=rp.R, sqrt_info=rp.sqrt_info))
        ofac.append(sim.cauchy_correct(*rp.EvaluateCeres([p.poses[rp.index]]), 1.0) + ([("pose", rp.index)],))
    mi.preMarginalize({"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat})
    mi.marginalize(schur_only=True)
    assert mi.status == 0, hex(mi.status)
    idx = mi.parameter_block_idx
    n_feat = sum(1 for k in idx if k[0] == "feature")
    assert mi.m == n_feat and mi.n == 18 * 15 and sorted(v for k, v in idx.items() if k[0] == "feature") == list(range(n_feat))
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k not in const])
            for r, js, keys in ofac]
    ref = O.vins_mono_marginalize(facs, mi.pos, mi.m, eps=1e-8)
    S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], mi.m)
    e_ref, e_gpu = rel_err(ref["A_red"], S_hp), rel_err(mi.A_red, S_hp)
    print(f"reduced camera system vs 80-bit truth: literal FP64 {e_ref:.2e}, CUDA {e_gpu:.2e}")
    assert e_gpu <= max(1e-9, 2.0 * e_ref)
    assert rel_err(mi.b_red, s_hp) <= max(1e-9, 2.0 * rel_err(ref["b_red"], s_hp))
    assert np.array_equal(mi.A_red != 0, mi.A_red.T != 0)      # symmetric fill pattern


def test_marginalization_with_projection_td_factors(backend):
    """BASELINE configs[3] (online td estimation): VINS-Mono style marginalization of the oldest frame where
    every visual factor is a ProjectionTdFactor and the time offset para_Td is a kept 1-d block.  Neither the
    class nor the factor exists in the reference: oracle = the restated VINS-Mono formulas."""
    p = sim.make_problem(sim.seed_for(9, 3), n_features=100, max_track=9, host0=0.6)
    rng = np.random.default_rng(93)
    tr = 0.033 / 480
    td = np.array([0.007])
    const = {("ex_pose", 0)}
    mi = MarginalizationInfo(backend, eps=1e-8, cauchy_a=1.0, constant=list(const), tr_over_row=tr)
    ofac = []
    keys = [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)]
    mi.addResidualBlockInfo(ResidualBlockInfo("imu", keys, drop_set=[0, 1], preint=p.imu_pre[0].pack()))
    ofac.append(O.IMUFactor(p.imu_pre[0]).EvaluateCeres([p.poses[0], p.sbs[0], p.poses[1], p.sbs[1]]) + (keys,))
    s = p.cfg.proj_sqrt_info
    for k in range(p.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        if i != 0:
            continue
        keys = [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f), ("td", 0)]
        pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
        m = dict(velocity_i=rng.normal(0, 0.3, 2), velocity_j=rng.normal(0, 0.3, 2), td_i=rng.normal(0, 0.004),
                 td_j=rng.normal(0, 0.004), row_i=rng.uniform(-240, 240), row_j=rng.uniform(-240, 240))
        mi.addResidualBlockInfo(ResidualBlockInfo("projection_td", keys, drop_set=[0, 3], pts_i=pts_i, pts_j=pts_j, **m))
        fac = O.ProjectionTdFactor(pts_i, pts_j, m["velocity_i"], m["velocity_j"], m["td_i"], m["td_j"], m["row_i"],
                                   m["row_j"], s, tr)
        r, js = fac.EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1], td])
        ofac.append(sim.cauchy_correct(r, js, 1.0) + (keys,))
    se3 = p.se3[0]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", [("pose", 0)], drop_set=[0], t=se3.t, R=se3.R, sqrt_info=se3.sqrt_info))
    ofac.append(sim.cauchy_correct(*se3.EvaluateCeres([p.poses[0]]), 1.0) + ([("pose", 0)],))
    mi.preMarginalize({"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat, "td": td})
    mi.marginalize()
    assert mi.status == 0, hex(mi.status)
    idx = mi.parameter_block_idx
    assert idx[("td", 0)] >= mi.m                      # the time offset is kept, with its own column
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k not in const])
            for r, js, keys in ofac]
    ref = O.vins_mono_marginalize(facs, mi.pos, mi.m, eps=1e-8)
    S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], mi.m)
    e_ref, e_gpu = rel_err(ref["A_red"], S_hp), rel_err(mi.A_red, S_hp)
    print(f"td marginalization vs 80-bit truth: literal FP64 {e_ref:.2e}, CUDA {e_gpu:.2e}")
    assert e_gpu <= max(1e-9, 2.0 * e_ref)
    assert rel_err(mi.b_red, s_hp) <= max(1e-9, 2.0 * rel_err(ref["b_red"], s_hp))
    t0 = idx[("td", 0)] - mi.m
    assert mi.A_red[t0, t0] > 0 and np.count_nonzero(mi.A_red[t0]) > 6      # td couples with the kept poses
    assert rel_err(mi.linearized_jacobians.T @ mi.linearized_jacobians, mi.A_red) <= 1e-9


def test_marginalization_factor_chains_two_rounds(backend):
    """VINS-Mono's frame-to-frame use of the class: round 1 marginalizes the oldest frame; the window slides
    (addr_shift); round 2 marginalizes the new oldest frame with the round-1 prior as a `MarginalizationFactor`
    residual block (marginalization_factor.cpp `MarginalizationFactor::Evaluate`, published algorithm -- IS-VINS
    deleted it, parity unpinned).  Checked: the factor's Evaluate (residual = r0 + J dx incl. the w < 0 sign
    branch, per-block Jacobians) against the oracle, then the round-2 reduced system against the 80-bit truth."""
    p = sim.make_problem(sim.seed_for(9, 13), n_features=160, max_track=9, host0=0.4)
    const = {("ex_pose", 0)}
    mi, _, _ = _build(backend, p, 1.0, True)
    para1 = {"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat}
    mi.preMarginalize(para1)
    mi.marginalize()
    assert mi.status == 0
    keep1 = mi.getParameterBlocks()
    # ---- the window slides: frame i -> i - 1; the next solve moves the estimates a little ----------------
    shift = {k: ((k[0], k[1] - 1) if k[0] in ("pose", "speed_bias") else k) for k, _, _ in keep1}
    keys2 = mi.getParameterBlocks(addr_shift=shift)
    assert len(keys2) == len(keep1) and ("pose", 0) in keys2 and ("speed_bias", 0) in keys2
    rng = np.random.default_rng(5)
    poses2 = p.poses[1:].copy()
    poses2[:, 0:3] += rng.normal(0, 0.01, poses2[:, 0:3].shape)
    for q in poses2:
        dq = O.q_mul(O.quat_from_pose(q), np.concatenate([[1.0], rng.normal(0, 0.002, 3)]))
        dq /= np.linalg.norm(dq)
        q[3:6], q[6] = dq[1:4], dq[0]
    poses2[3, 3:7] *= -1.0                      # same rotation, other sign: the `w < 0` branch of Evaluate
    sbs2 = p.sbs[1:] + rng.normal(0, 0.003, p.sbs[1:].shape)
    feat2 = p.feat * (1.0 + rng.normal(0, 0.01, p.feat.shape))
    para2 = {"pose": poses2, "speed_bias": sbs2, "ex_pose": p.ex, "feature": feat2}
    mi2 = MarginalizationInfo(backend, eps=1e-8, cauchy_a=1.0, constant=list(const))
    ofac = []
    # the prior: never dropped itself; the blocks of the new oldest frame are in its drop set
    drop = [c for c, k in enumerate(keys2) if k in (("pose", 0), ("speed_bias", 0))]
    mi2.addResidualBlockInfo(ResidualBlockInfo("marginalization", keys2, drop_set=drop, prior=mi))
    ofa = O.MarginalizationFactor(mi.linearized_jacobians, mi.linearized_residuals,
                                  [(size, idx) for _, size, idx in keep1], [mi.keep_block_data[k] for k, _, _ in keep1])
    val = lambda k: {"pose": poses2, "speed_bias": sbs2, "ex_pose": p.ex}[k[0]][k[1]]
    r_pr, j_pr = ofa.EvaluateCeres([val(k) for k in keys2])
    ofac.append((r_pr, j_pr, keys2))
    keys = [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)]
    mi2.addResidualBlockInfo(ResidualBlockInfo("imu", keys, drop_set=[0, 1], preint=p.imu_pre[1].pack()))
    ofac.append(O.IMUFactor(p.imu_pre[1]).EvaluateCeres([poses2[0], sbs2[0], poses2[1], sbs2[1]]) + (keys,))
    s = p.cfg.proj_sqrt_info
    nproj = 0
    for k in range(p.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        if i != 1:
            continue
        nproj += 1
        keys = [("pose", 0), ("pose", j - 1), ("ex_pose", e), ("feature", f)]
        pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
        mi2.addResidualBlockInfo(ResidualBlockInfo("projection", keys, drop_set=[0, 3], pts_i=pts_i, pts_j=pts_j))
        r, js = O.ProjectionFactor(pts_i, pts_j, s).EvaluateCeres([poses2[0], poses2[j - 1], p.ex[e], feat2[f:f + 1]])
        ofac.append(sim.cauchy_correct(r, js, 1.0) + (keys,))
    assert nproj > 10
    mi2.preMarginalize(para2)
    # ---- MarginalizationFactor::Evaluate ---------------------------------------------------------------
    assert rel_err(mi2.prior_residuals, r_pr) <= 1e-12
    for jg, jo in zip(mi2.prior_jacobians, j_pr):
        assert np.array_equal(jg, jo)
    flipped = ofa.keep[list(keys2).index(("pose", 3))][1]
    unflipped = [val(k) * (np.array([1, 1, 1, -1, -1, -1, -1.0]) if k == ("pose", 3) else 1.0) for k in keys2]
    assert rel_err(ofa.EvaluateCeres(unflipped)[0], r_pr) <= 1e-12        # q and -q: the same point
    mi2.marginalize()
    assert mi2.status == 0, hex(mi2.status)
    idx = mi2.parameter_block_idx
    assert idx[("pose", 0)] == 0 and idx[("speed_bias", 0)] == 6
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k not in const])
            for r, js, keys in ofac]
    ref = O.vins_mono_marginalize(facs, mi2.pos, mi2.m, eps=1e-8)
    assert ref["min_eig_Amm"] > 1e-8
    S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], mi2.m)
    e_ref, e_gpu = rel_err(ref["A_red"], S_hp), rel_err(mi2.A_red, S_hp)
    print(f"round-2 A_red vs 80-bit truth: literal FP64 {e_ref:.2e}, CUDA {e_gpu:.2e}; flipped pose column {flipped}")
    assert e_gpu <= max(1e-9, 2.0 * e_ref)
    assert rel_err(mi2.b_red, s_hp) <= max(1e-9, 2.0 * rel_err(ref["b_red"], s_hp))
    J, r = mi2.linearized_jacobians, mi2.linearized_residuals
    assert rel_err(J.T @ J, mi2.A_red) <= 1e-9
    tol = max(1e-9, 4.0 * e_ref)
    assert rel_err(J.T @ J, ref["linearized_jacobians"].T @ ref["linearized_jacobians"]) <= tol
    # the kept set of round 2 still carries every block the round-1 prior touched, except the dropped frame
    kept2 = {k for k, _, _ in mi2.getParameterBlocks()}
    assert kept2 >= (set(keys2) - {("pose", 0), ("speed_bias", 0)} - const)


def test_config4_window20_td_marginalization(backend):
    """BASELINE configs[3] variant (b): WINDOW_SIZE = 20, ~2000 live features, online td estimation,
    VINS-Mono style MARGIN_OLD through the facade: the steady-state previous prior over the whole window
    (MarginalizationFactor), the IMU factor 0 -> 1 and ~1750 ProjectionTdFactors hosted in the oldest frame;
    extrinsics estimated, para_Td kept => n_keep = 20 * 15 + 6 + 1 = 307.  Inputs = the committed fixture
    tests/golden/problem_W20_F2000_td.npz (checked here against its generator).  Parity unpinned by the reference."""
    import os
    from is_vins_b200 import FactorProblem, PriorState, add_margin_old_blocks
    path = os.path.join(os.path.dirname(__file__), "golden", "problem_W20_F2000_td.npz")
    fp, z = FactorProblem.load(path), np.load(path)
    p = sim.w20_problem()
    assert np.array_equal(fp.pose, p.poses) and np.array_equal(fp.proj_obs, p.proj_obs) and np.array_equal(fp.feature, p.feat)
    tdx = sim.make_td_observations(p, sim.W20_SEED + 1)
    keys, Jp, r0, x0 = sim.make_window_prior(p, sim.W20_SEED + 2)
    assert np.array_equal(z["td_obs"], tdx["td_obs"]) and np.array_equal(z["prior_J"], Jp) and np.array_equal(z["prior_r0"], r0)
    td, tr = z["td"], float(z["tr_over_row"][0])
    prior = PriorState(keys, z["prior_J"], z["prior_r0"], x0)
    mi = MarginalizationInfo(backend, eps=1e-8, cauchy_a=1.0, tr_over_row=tr)       # ESTIMATE_EXTRINSIC: no constant block
    blocks = add_margin_old_blocks(mi, fp, z["td_obs"], prior)
    para = {"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat, "td": td}
    mi.preMarginalize(para)
    mi.marginalize()
    assert mi.status == 0, hex(mi.status)
    idx = mi.parameter_block_idx
    L0 = sum(1 for k in idx if k[0] == "feature")
    n_fac = sum(1 for b in blocks if b.kind == "projection_td")
    assert mi.n == 20 * 15 + 6 + 1 and mi.m == 15 + L0 and 200 <= L0 <= 300 and 1400 <= n_fac <= 2200
    assert idx[("pose", 0)] == 0 and idx[("speed_bias", 0)] == 6 and idx[("td", 0)] >= mi.m and idx[("ex_pose", 0)] >= mi.m
    # the oracle's restatement of the same residual blocks, in the block order the facade chose
    s = p.cfg.proj_sqrt_info
    ofac = []
    for b in blocks:
        val = [np.atleast_1d(np.asarray(para[k[0]][k[1]], float)) for k in b.parameter_blocks]
        M = b.members
        if b.kind == "marginalization":
            mf = O.MarginalizationFactor(Jp, r0, [(size, i) for _, size, i in prior.getParameterBlocks()], x0)
            r, js = mf.EvaluateCeres(val)
            assert rel_err(mi.prior_residuals, r) <= 1e-12
        elif b.kind == "imu":
            r, js = O.IMUFactor(p.imu_pre[0]).EvaluateCeres(val)
        else:
            fac = O.ProjectionTdFactor(M["pts_i"], M["pts_j"], M["velocity_i"], M["velocity_j"], M["td_i"], M["td_j"],
                                       M["row_i"], M["row_j"], s, tr)
            r, js = sim.cauchy_correct(*fac.EvaluateCeres(val), 1.0)
        ofac.append((r, js, b.parameter_blocks))
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys_, js)]) for r, js, keys_ in ofac]
    ref = O.vins_mono_marginalize(facs, mi.pos, mi.m, eps=1e-8)
    assert ref["min_eig_Amm"] > 1e-8
    S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], mi.m)
    e_ref, e_gpu = rel_err(ref["A_red"], S_hp), rel_err(mi.A_red, S_hp)
    print(f"configs[3] (b) n_keep={mi.n} m={mi.m} factors={n_fac}: A_red vs 80-bit truth: literal FP64 {e_ref:.2e}, CUDA {e_gpu:.2e}")
    assert e_gpu <= max(1e-9, 2.0 * e_ref)
    assert rel_err(mi.b_red, s_hp) <= max(1e-9, 2.0 * rel_err(ref["b_red"], s_hp))
    J, r = mi.linearized_jacobians, mi.linearized_residuals
    assert mi.rank == mi.n                                   # the previous prior constrains every kept direction
    assert rel_err(J.T @ J, mi.A_red) <= 1e-9 and rel_err(J.T @ r, mi.b_red) <= 1e-9
    lam = np.linalg.norm(J, axis=1) ** 2
    assert np.all(np.diff(lam) >= 0)
    assert np.allclose(lam, np.linalg.eigvalsh(0.5 * (S_hp + S_hp.T)), rtol=1e-7)


def test_rank_deficient_dense_block_takes_the_eigen_path_and_is_flagged(backend):
    """The dense marginalized block normally gets its inverse on a fast path (Gauss-Jordan + a provable eigenvalue
    bound).  When the block is rank deficient -- here pose 0 is dropped with only two visual factors (4 residual rows
    for 6 tangent dimensions) constraining it -- the literal VINS-Mono step must run instead: eigen-decomposition,
    eigenvalues <= eps zeroed (pseudo-inverse), ISV_W_RANK_DEFICIENT raised.  eps = 1 keeps the threshold above the
    rounding noise of the null directions so that the oracle's `eigh` and the CUDA Jacobi agree on what is dropped."""
    p = sim.make_problem(sim.seed_for(9, 31), n_features=40, max_track=6, host0=0.5)
    hosted, seen = [], set()
    for k in range(p.proj_idx.shape[1]):                    # the first observation of two different features of frame 0
        if int(p.proj_idx[0, k]) == 0 and int(p.proj_idx[3, k]) not in seen and len(hosted) < 2:
            seen.add(int(p.proj_idx[3, k]))
            hosted.append(k)
    assert len(hosted) == 2
    const = {("ex_pose", 0)}
    mi = MarginalizationInfo(backend, eps=1.0, cauchy_a=0.0, constant=list(const))
    ofac = []
    s = p.cfg.proj_sqrt_info
    for k in hosted:
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        keys = [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)]
        pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
        mi.addResidualBlockInfo(ResidualBlockInfo("projection", keys, drop_set=[0], pts_i=pts_i, pts_j=pts_j))   # features kept
        ofac.append(O.ProjectionFactor(pts_i, pts_j, s).EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1]]) + (keys,))
    # a prior on the kept pose(s) so that the reduced system is not itself degenerate
    rel = p.rel[1]
    keys = [("pose", rel.imu_i), ("pose", rel.imu_j)]
    mi.addResidualBlockInfo(ResidualBlockInfo("rel", keys, delta_t=rel.delta_t, delta_R=rel.delta_R, sqrt_info=rel.sqrt_info))
    ofac.append(rel.EvaluateCeres([p.poses[rel.imu_i], p.poses[rel.imu_j]]) + (keys,))
    mi.preMarginalize({"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat})
    mi.marginalize()
    assert mi.status & 0x02, hex(mi.status)                 # ISV_W_RANK_DEFICIENT
    assert mi.status & ~0x02 == 0, hex(mi.status)
    idx = mi.parameter_block_idx
    assert mi.m == 6 and idx[("pose", 0)] == 0
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k not in const])
            for r, js, keys in ofac]
    ref = O.vins_mono_marginalize(facs, mi.pos, mi.m, eps=1.0)
    lam_mm = np.linalg.eigvalsh(0.5 * (ref["A"][:6, :6] + ref["A"][:6, :6].T))
    assert int(np.sum(lam_mm > 1.0)) == 4 and lam_mm[1] < 1e-3 * lam_mm[2]     # rank 4, clear gap
    assert rel_err(mi.A_red, ref["A_red"]) <= 1e-9 and rel_err(mi.b_red, ref["b_red"]) <= 1e-9
    assert np.all(np.isfinite(mi.linearized_jacobians))
    assert rel_err(mi.linearized_jacobians.T @ mi.linearized_jacobians, ref["linearized_jacobians"].T @ ref["linearized_jacobians"]) <= 1e-9


def test_more_dense_marginalized_columns_than_kept_columns_in_a_batch(backend):
    """m_dense = 15 (pose 0 + speed-bias 0) against a single kept pose (n = 6).  The Schur kernel's
    per-problem scratch T = A_rm pinv is n x m_dense: with a stride of n * n the CTAs of a batch overwrote each other's
    slice (and the last one wrote past the allocation).  Every problem of a 5-problem batch must reproduce the
    single-problem result, which must match the VINS-Mono oracle."""
    import ctypes as C

    import torch

    from is_vins_b200 import capi
    p = sim.make_problem(sim.seed_for(9, 77), n_features=8, max_track=4)
    mi = MarginalizationInfo(backend, eps=1e-8, cauchy_a=0.0)
    se3, rel, vb = p.se3[0], p.rel[0], p.vb[0]
    ofac = []
    keys = [("pose", 0)]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", keys, drop_set=[0], t=se3.t, R=se3.R, sqrt_info=se3.sqrt_info))
    ofac.append(se3.EvaluateCeres([p.poses[0]]) + (keys,))
    keys = [("pose", 0), ("pose", 1)]
    mi.addResidualBlockInfo(ResidualBlockInfo("rel", keys, drop_set=[0], delta_t=rel.delta_t, delta_R=rel.delta_R,
                                              sqrt_info=rel.sqrt_info))
    ofac.append(rel.EvaluateCeres([p.poses[0], p.poses[1]]) + (keys,))
    keys = [("speed_bias", 0)]
    mi.addResidualBlockInfo(ResidualBlockInfo("vb", keys, drop_set=[0], VB=vb.VB, sqrt_info=vb.sqrt_info))
    ofac.append(vb.EvaluateCeres([p.sbs[0]]) + (keys,))
    mi.preMarginalize({"pose": p.poses, "speed_bias": p.sbs, "ex_pose": p.ex, "feature": p.feat})
    mi.marginalize(keep_tables=True)
    assert mi.status == 0 and mi.m == 15 and mi.n == 6
    idx = mi.parameter_block_idx
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js)]) for r, js, keys in ofac]
    ref = O.vins_mono_marginalize(facs, mi.pos, mi.m, eps=1e-8)
    assert rel_err(mi.A_red, ref["A_red"]) <= 1e-9 and rel_err(mi.b_red, ref["b_red"]) <= 1e-9
    # the same problem five times in one call
    gi, tabs = mi._gi, mi._tables
    NP, nf, pos, n = 5, gi.n_factors, gi.pos, mi.n
    fa = np.frombuffer(tabs["factors_bytes"], dtype=np.dtype([("res", "<i8"), ("nres", "<i4"), ("nb", "<i4"),
                                                               ("fb", "<i4"), ("prob", "<i4")])).copy()
    big = np.tile(fa, NP)
    big["prob"] = np.repeat(np.arange(NP, dtype=np.int32), nf)
    dev = "cuda:0"
    d_f = torch.from_numpy(big.view(np.uint8)).to(dev)
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device=dev)
    o = {"A": z(NP, pos, pos), "b": z(NP, pos), "A_red": z(NP, n, n), "b_red": z(NP, n), "J": z(NP, n, n), "r": z(NP, n),
         "rank": torch.zeros((NP,), dtype=torch.int32, device=dev), "status": torch.zeros((NP,), dtype=torch.int32, device=dev)}
    gi.n_problems, gi.n_factors, gi.factors = NP, NP * nf, d_f.data_ptr()
    go = type(mi._go)(o["A"].data_ptr(), o["b"].data_ptr(), o["A_red"].data_ptr(), o["b_red"].data_ptr(), o["J"].data_ptr(),
                      o["r"].data_ptr(), o["rank"].data_ptr(), o["status"].data_ptr())
    capi.check(backend.lib.isv_marginalize_generic(backend.h, C.byref(gi), C.byref(go)), "isv_marginalize_generic")
    backend.synchronize()
    assert int(torch.count_nonzero(o["status"]).item()) == 0
    for q in range(NP):
        A = o["A_red"][q].T.cpu().numpy()
        assert rel_err(A, mi.A_red) <= 1e-12, q
        assert rel_err(o["b_red"][q].cpu().numpy(), mi.b_red) <= 1e-12, q
        J = o["J"][q].T.cpu().numpy()
        assert rel_err(J.T @ J, ref["A_red"]) <= 1e-9, q
