"""-m gpu: randomised parity sweep of the window kernels (through the C ABI) against the C oracle on thousands of
DISTINCT windows, plus the targeted edge cases the reference's formulas branch on (VERDICT r1 "next round" 3):

  * >= 1024 distinct windows per shape: L ~ 150, L ~ 1000 and a wide ragged shape L ~ U{0..200} (empty windows and
    the 31/32/33/63/64/65 lane-tile edges included); structured C oracle on all, the literal dense algorithm of the
    reference on a subset; the error histogram is printed;
  * residual rotation of SE3PriorFactor / RelativePoseFactor on both sides of Sophus' near-pi cut
    (include/utility/sophus_utils.hpp:224-231: |phi| < pi - 1e-5) and of the small-angle cut (|phi|^2 <= 1e-10);
  * an eigenvalue of MargBackward's Lamda_prior within 1e-6 (relative) of ALPHA on both sides: the strict `> ALPHA`
    cut (src/estimator.cpp:1482) must keep / drop exactly that eigenvalue;
  * vioRollPitchEdges[0] valid / invalid (covAbs, src/estimator.cpp:1265-1271);
  * L = 0 with a healthy prior.

Tolerance 1e-9 relative Frobenius per recovered-factor block (north_star), ranks exact, status 0."""
import math

import numpy as np
import pytest

import bench
from is_vins_b200 import DeviceBatch, MargBackend, capi
from is_vins_b200.batch import outputs_rel_diff
from oracle import isv_oracle as O
from oracle import ref_c

TOL = 1e-9
pytestmark = pytest.mark.gpu
FIELDS = ("se3", "pg", "rel", "vb", "rp", "rank", "status")


def _run_gpu(backend, batch, which=capi.RUN_BOTH):
    db = DeviceBatch(batch, "cuda:0")
    backend.marg_window_batch(db, which)
    backend.synchronize()
    return db.outputs()


def _histogram(err):
    edges = [0, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, np.inf]
    h, _ = np.histogram(err, bins=edges)
    return " ".join(f"<{e:g}:{c}" for e, c in zip(edges[1:], h))


@pytest.mark.parametrize("L,ragged,n,seed", [(150, 0.25, 1536, 101), (1000, 0.25, 1024, 102), (100, 1.0, 2048, 103)])
def test_randomised_sweep_matches_c_oracle(backend, L, ragged, n, seed):
    counts = None
    if ragged == 1.0:
        # uniform in 0..2L, with the empty window and the lane-tile edges planted explicitly
        counts = np.random.default_rng(seed).integers(0, 2 * L + 1, n)
        edges = [0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 2 * L]
        counts[:len(edges)] = edges
    batch = bench.make_batch(L, n, seed, ragged=ragged, counts=counts)
    counts = np.diff(batch.lm_offset)
    out = _run_gpu(backend, batch)
    ref = ref_c.marg_window_batch(batch, 3, 0, True)
    err = outputs_rel_diff(out, ref, 3)
    worst = int(np.argmax(err))
    print(f"\nsweep L~{L} ragged {ragged}: {n} distinct windows, worst rel err {err.max():.3e} at window {worst} "
          f"(L = {counts[worst]}), median {np.median(err):.2e}; histogram {_histogram(err)}")
    assert np.all(np.isfinite(err)) and err.max() <= TOL, (worst, err.max())
    assert np.array_equal(out.rank, ref.rank)
    assert not out.status.any() and not ref.status.any()
    # the literal dense algorithm of the reference (FullPivLU of the (L+6)^2 block) on a subset
    idx = np.linspace(0, n - 1, 24 if L >= 1000 else 96).round().astype(np.int64)
    lit = ref_c.marg_window_batch(batch.take(idx), 3, 0, False)
    sub = type(out)(*[getattr(out, f)[idx] for f in FIELDS])
    el = outputs_rel_diff(sub, lit, 3)
    print(f"   literal subset ({len(idx)} windows): worst {el.max():.3e}")
    assert el.max() <= TOL and np.array_equal(sub.rank, lit.rank)


def test_fused_event_kernel_matches_c_oracle_and_batch_kernels(backend):
    """The one-launch fused kernel (one CTA per window: marg_event_fused_kernel, the route of isv_marg_event and of small
    device batches) on 148 distinct ragged windows -- incl. the empty window, fewer landmarks than the seven landmark warps,
    the lane-tile edges and two L ~ 1000 windows -- against the C oracle and against the warp-per-window batch kernels:
    same device functions, only the summation order of the landmark Gram differs."""
    n, seed = 148, 104
    counts = np.random.default_rng(seed).integers(0, 301, n)
    edges = [0, 1, 2, 6, 7, 8, 31, 32, 33, 63, 64, 65, 127, 128, 129, 223, 224, 225, 896, 897, 1000, 1250]
    counts[:len(edges)] = edges
    batch = bench.make_batch(150, n, seed, ragged=1.0, counts=counts)
    backend.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 148)
    l0 = backend.launch_count
    out = _run_gpu(backend, batch)
    assert backend.launch_count - l0 == 1
    backend.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 0)
    try:
        l0 = backend.launch_count
        outb = _run_gpu(backend, batch)
        assert backend.launch_count - l0 == 7   # 2 zero-fills, 2 x factor Jacobians, landmark phase, tail, backward
    finally:
        backend.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 148)
    ref = ref_c.marg_window_batch(batch, 3, 0, True)
    err = outputs_rel_diff(out, ref, 3)
    worst = int(np.argmax(err))
    print(f"\nfused kernel: {n} windows, worst rel err vs C oracle {err.max():.3e} at window {worst} (L = {counts[worst]}), "
          f"median {np.median(err):.2e}; vs batch kernels {outputs_rel_diff(out, outb, 3).max():.3e}")
    assert np.all(np.isfinite(err)) and err.max() <= TOL
    assert np.array_equal(out.rank, ref.rank) and not out.status.any()
    assert outputs_rel_diff(out, outb, 3).max() <= 1e-11
    # discrete outcomes agree exactly on both routes
    assert np.array_equal(out.rank, outb.rank) and np.array_equal(out.status, outb.status)
    # status bits travel through shared memory in the fused kernel: a non-unit quaternion must still be flagged
    bad = batch.take(np.arange(4))
    bad.pose_fwd = bad.pose_fwd.copy()
    bad.pose_fwd[2, 0, 3:7] *= 1.001
    ob = _run_gpu(backend, bad)
    assert ob.status[2] & 0x08          # ISV_W_NONUNIT_QUAT
    assert not ob.status[[0, 1, 3]].any()


@pytest.mark.parametrize("L", [0, 5, 150, 4000])
def test_event_call_all_routes_and_staging_limits(backend, L):
    """isv_marg_event on its three routes against the C oracle: L = 0 (no landmark block), 5 (fewer landmarks than landmark
    warps), 150, and 4000 (the event no longer fits the kernel's shared-memory staging area: only the records are staged by
    the bulk copy, the landmarks are read in place)."""
    b = bench.make_batch(max(L, 1), 1, 900 + L, ragged=0.0, counts=np.array([L]))
    ref = ref_c.marg_window_batch(b, 3, 0, True)
    o = b.lm_obs
    fwd = (b.pose_fwd[0, 0], b.pose_fwd[0, 1], b.ex_pose, o[5], np.ascontiguousarray(o[0:3].T),
           np.ascontiguousarray(np.vstack([o[3:5], np.ones((1, o.shape[1]))]).T), b.prior_se3[0], b.prior_rel[0], b.prior_rp[0])
    bwd = (b.pose_bwd[0, 0], b.sb_bwd[0, 0], b.pose_bwd[0, 1], b.sb_bwd[0, 1], b.prior_vb[0], b.preint[0])
    try:
        for mode in (0, 1, 2, 0):
            backend.set_tuning(capi.TUNE_EVENT_MODE, mode)
            (se3, pg, rkf, stf), (rel, vb, rp, rkb, stb) = backend.marg_event(fwd, bwd)
            out = type(ref)(se3[None], pg[None], rel[None], vb[None], rp[None], np.array([[rkf, rkb]], np.int32), np.array([stf], np.int32))
            err = outputs_rel_diff(out, ref, 3).max()
            assert err <= TOL, (L, mode, err)
            assert (rkf, rkb) == (int(ref.rank[0, 0]), int(ref.rank[0, 1])) and stf == 0 and stb == 0, (L, mode)
    finally:
        backend.set_tuning(capi.TUNE_EVENT_MODE, 0)


def _rot(axis, angle):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    return O.SO3.exp(axis * angle).matrix()


def _fwd_input(b, w):
    a, c = int(b.lm_offset[w]), int(b.lm_offset[w + 1])
    rp_valid = b.prior_rp is not None and b.prior_rp[w, 0] != 0.0
    return O.ForwardInput(b.pose_fwd[w, 0], b.pose_fwd[w, 1], b.ex_pose, b.lm_obs[5, a:c].copy(),
                          np.ascontiguousarray(b.lm_obs[0:3, a:c].T), np.ascontiguousarray(np.vstack([b.lm_obs[3:5, a:c], np.ones(c - a)]).T),
                          b.prior_se3[w, 0:3], b.prior_se3[w, 3:12].reshape(3, 3).T, b.prior_se3[w, 12:48].reshape(6, 6).T,
                          b.prior_rel[w, 0:3], b.prior_rel[w, 3:12].reshape(3, 3).T, b.prior_rel[w, 12:48].reshape(6, 6).T,
                          bool(rp_valid), b.prior_rp[w, 1:5].reshape(2, 2).T if rp_valid else None)


# angles: both sides of the near-pi cut (pi - 1e-5) and of the small-angle cut (1e-5), and far inside each branch
ANGLES = [math.pi - 1e-3, math.pi - 2e-5, math.pi - 5e-6, math.pi - 1e-7, 2e-5, 5e-6, 1e-8, 0.0, 1.0, 2.5]


def test_residual_rotation_branches_of_the_prior_factors(backend):
    """Windows whose SE3 prior / relative-pose prior disagree with the estimate by a rotation of a chosen angle: every
    branch of SO3::log and rightJacobianInvSO3 is taken.  Checked against BOTH restatements (C and NumPy)."""
    base = bench.make_batch(60, len(ANGLES) * 2, 301)
    rng = np.random.default_rng(5)
    for k, ang in enumerate(ANGLES):
        axis = rng.normal(size=3)
        E = _rot(axis, ang)
        # SE3PriorFactor: residual rotation = R_prior^T R_0  ->  R_prior = R_0 E^T           (window 2k)
        w = 2 * k
        R0 = O.q_to_R(O.quat_from_pose(base.pose_fwd[w, 0]))
        base.prior_se3[w, 3:12] = (R0 @ E.T).flatten(order="F")
        # RelativePoseFactor: residual rotation = delta_R R_1^T R_0  ->  delta_R = E R_0^T R_1   (window 2k+1)
        w = 2 * k + 1
        R0 = O.q_to_R(O.quat_from_pose(base.pose_fwd[w, 0]))
        R1 = O.q_to_R(O.quat_from_pose(base.pose_fwd[w, 1]))
        base.prior_rel[w, 3:12] = (E @ R0.T @ R1).flatten(order="F")
    out = _run_gpu(backend, base, capi.RUN_FORWARD)
    ref = ref_c.marg_window_batch(base, 1, 0, True)
    err = outputs_rel_diff(out, ref, 1)
    # at pi - 1e-7 the logarithm's own conditioning (d angle / d R ~ 1 / sin) amplifies the 1e-16 input rounding
    # to ~1e-9: that single angle gets 1e-7
    tol = np.array([1e-7 if abs(ANGLES[w // 2] - (math.pi - 1e-7)) < 1e-12 else TOL for w in range(base.n)])
    print("\nresidual-rotation sweep, rel err per angle (se3 prior, rel-pose prior):")
    for k, ang in enumerate(ANGLES):
        print(f"   angle {ang:.9f}: {err[2 * k]:.2e} {err[2 * k + 1]:.2e}")
    assert np.all(err <= tol), err
    assert np.array_equal(out.rank[:, 0], ref.rank[:, 0]) and not out.status.any()
    # and against the NumPy oracle on the near-pi / small-angle windows
    cfg = O.Config()
    for w in range(base.n):
        fo = O.marg_forward(_fwd_input(base, w), cfg, structured=True)
        e = np.linalg.norm(out.se3_sqrt_info(w) - fo.se3_sqrt_info) / np.linalg.norm(fo.se3_sqrt_info)
        e = max(e, np.linalg.norm(out.pg_sqrt_info(w) - fo.pg_sqrt_info) / np.linalg.norm(fo.pg_sqrt_info))
        assert e <= tol[w], (w, ANGLES[w // 2], e)


def test_alpha_next_to_an_eigenvalue_keeps_or_drops_exactly_it():
    """ALPHA = lambda_k (1 -/+ 1e-6): the strict `> ALPHA` cut (src/estimator.cpp:1482) must keep lambda_k on one side
    and drop it on the other; everything recovered from the truncated spectrum must match the oracle run at the same
    ALPHA."""
    batch = bench.make_batch(40, 6, 401)
    cfg_o = O.Config()
    for w in range(batch.n):
        bin_ = O.BackwardInput(batch.pose_bwd[w, 0], batch.sb_bwd[w, 0], batch.pose_bwd[w, 1], batch.sb_bwd[w, 1],
                               batch.prior_vb[w, 0:9], batch.prior_vb[w, 9:90].reshape(9, 9).T, _preint(batch, w, cfg_o))
        lam = np.sort(O.marg_backward(bin_, cfg_o).eigvals)[-15:]
        k = 3 + (w % 9)                                  # an eigenvalue inside the non-zero spectrum
        for side, keep in ((1.0 - 1e-6, 15 - k), (1.0 + 1e-6, 15 - k - 1)):
            alpha = float(lam[k] * side)
            cfg = capi.default_config()
            cfg.alpha = alpha
            be = MargBackend(0, cfg)
            one = batch.take([w])
            out = be.marg_window_batch_host(one, capi.RUN_BACKWARD)
            be.close()
            cfg_c = capi.default_config()
            cfg_c.alpha = alpha
            ref = ref_c.marg_window_batch(one, 2, 1, True, cfg_c)
            assert int(ref.rank[0, 1]) == keep, (w, side, ref.rank, keep)
            assert int(out.rank[0, 1]) == keep, (w, side, out.rank, keep)
            # the truncated pseudo-inverse is rank deficient: sqrt-info of the recovered factors is only defined where
            # the reference's own LLT succeeds; compare whatever the oracle reports as finite
            if np.all(np.isfinite(ref.rel)) and np.all(np.isfinite(ref.vb)) and not ref.status.any():
                assert outputs_rel_diff(out, ref, 2).max() <= 1e-7, (w, side)


def _preint(b, w, cfg):
    """IntegrationBase rebuilt from the raw IMU samples of window w (the record in `preint` came from the same call)"""
    init = b.imu_init[w]
    pre = O.IntegrationBase(init[0:3], init[3:6], init[6:9], init[9:12], cfg)
    for s in b.imu_raw[w]:
        pre.push_back(float(s[0]), s[1:4], s[4:7])
    assert np.allclose(pre.pack(), b.preint[w], rtol=1e-12, atol=1e-15)
    return pre


def test_rollpitch_edge_valid_and_invalid(backend):
    """covAbs of the pose-graph record = (s^T s)^-1 of vioRollPitchEdges[0] when its index is 0, else zero."""
    b = bench.make_batch(80, 64, 501)
    valid = b.prior_rp[:, 0] != 0
    # force both cases to be present in numbers
    s = np.array([[30.0, 4.0], [0.0, 25.0]])
    b.prior_rp[::2, 0] = 1.0
    b.prior_rp[::2, 1:5] = s.flatten(order="F") * np.linspace(0.8, 1.2, 32)[:, None]
    b.prior_rp[1::2] = 0.0
    out = _run_gpu(backend, b, capi.RUN_FORWARD)
    ref = ref_c.marg_window_batch(b, 1, 0, True)
    assert outputs_rel_diff(out, ref, 1).max() <= TOL and not out.status.any()
    assert np.all(out.pg[1::2, 85:89] == 0.0)
    for w in range(0, 64, 2):
        sw = b.prior_rp[w, 1:5].reshape(2, 2).T
        assert np.allclose(out.pg_covAbs(w), np.linalg.inv(sw.T @ sw), rtol=1e-12)
    del valid


def test_no_landmarks_with_a_healthy_prior(backend):
    """L = 0: MargPointIdx empty.  The prior + relative-pose information alone still give a full-rank Lamda_prior."""
    b = bench.make_batch(0, 256, 601, ragged=0.0)
    assert b.n_landmarks == 0
    out = _run_gpu(backend, b)
    ref = ref_c.marg_window_batch(b, 3, 0, False)            # literal algorithm: the dense Lamda is just 12 x 12
    err = outputs_rel_diff(out, ref, 3)
    print(f"\nL = 0: worst rel err {err.max():.2e}")
    assert err.max() <= TOL and np.array_equal(out.rank, ref.rank) and np.all(out.rank[:, 0] == 6)
    assert not out.status.any()


def test_full_size_ragged_batch_is_schedule_independent(backend):
    """BASELINE-size ragged batch (9472 distinct windows, L ~ 1000): bit-identical results from two runs and from a
    run on the reversed window order (a window's result may not depend on which warp / SM / neighbours it gets)."""
    b = bench.make_batch(1000, 9472, 701)
    out = _run_gpu(backend, b)
    out2 = _run_gpu(backend, b)
    assert not out.status.any()
    for f in FIELDS:
        assert np.array_equal(getattr(out, f), getattr(out2, f)), f
    rev = b.take(np.arange(b.n)[::-1])
    outr = _run_gpu(backend, rev)
    for f in FIELDS:
        assert np.array_equal(getattr(out, f), getattr(outr, f)[::-1]), f
    # size-independent properties: orthonormal rotations and upper-triangular positive square-root informations
    R = out.se3[:, 3:12].reshape(-1, 3, 3)
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-12
    S = out.se3[:, 12:48].reshape(-1, 6, 6).transpose(0, 2, 1)        # column-major records
    assert np.all(np.tril(S, -1) == 0) and np.all(np.diagonal(S, axis1=1, axis2=2) > 0)
    S9 = out.vb[:, 9:90].reshape(-1, 9, 9).transpose(0, 2, 1)
    assert np.all(np.tril(S9, -1) == 0) and np.all(np.diagonal(S9, axis1=1, axis2=2) > 0)


def test_raw_imu_and_z_one_inputs(backend):
    """ABI 2 inputs: raw IMU samples instead of the pre-integration record (preintegrate_kernel runs inside the batch
    call) and ISV_IN_PTS_I_Z_ONE (pts_i.z not read / not copied).  Forward results agree with the ABI 1 path to rounding,
    backward results match the oracle (whose record came from the NumPy IntegrationBase) at 1e-9; the host-pointer entry
    point gives the device path's bits and rejects a broken z promise."""
    b = bench.make_batch(150, 700, 801)
    ref = ref_c.marg_window_batch(b, 3, 0, True)
    out0 = _run_gpu(backend, b)
    db = DeviceBatch(b, "cuda:0", raw_imu=True, z_one=True)
    backend.marg_window_batch(db, capi.RUN_BOTH)
    backend.synchronize()
    out1 = db.outputs()
    # z == 1 folded into the arithmetic (F p = F[:,0] x + F[:,1] y + F[:,2]): same value, different rounding
    assert outputs_rel_diff(out1, out0, 1).max() <= 1e-12
    assert not out1.status.any() and np.array_equal(out1.rank, ref.rank)
    e_ref, e_01 = outputs_rel_diff(out1, ref, 3).max(), outputs_rel_diff(out1, out0, 2).max()
    print(f"\nraw-IMU path: vs oracle {e_ref:.2e}, vs record path {e_01:.2e}")
    assert e_ref <= TOL and e_01 <= 1e-10
    out2 = backend.marg_window_batch_host(b, capi.RUN_BOTH, raw_imu=True, z_one=True)
    for f in FIELDS:
        assert np.array_equal(getattr(out1, f), getattr(out2, f)), f
    out3 = backend.marg_window_batch_host(b, capi.RUN_BACKWARD, raw_imu=True)
    assert np.array_equal(out3.vb, out1.vb) and np.array_equal(out3.rel, out1.rel)
    # ABI 3: pts_i.x / pts_i.y handed over as the FP32 values they are in the reference (cv::Point2f): widened exactly on
    # load, so every bit of the result is the one the double inputs give -- on the device path and on the host path, with
    # the f64 components 0, 1 poisoned to prove they are not read
    from is_vins_b200.backend import xy_as_f32
    xyf = xy_as_f32(b.lm_obs)
    dbx = DeviceBatch(b, "cuda:0", raw_imu=True, z_one=True, xy_f32=True)
    dbx.t["lm_obs"][0:2].fill_(float("nan"))
    backend.marg_window_batch(dbx, capi.RUN_BOTH)
    backend.synchronize()
    outx = dbx.outputs()
    dby = DeviceBatch(b, "cuda:0", xy_f32=True)            # without the z promise: the other kernel instantiation
    backend.marg_window_batch(dby, capi.RUN_BOTH)
    backend.synchronize()
    outy = dby.outputs()
    keep = b.lm_obs[0:2].copy()
    b.lm_obs[0:2] = np.nan
    outh = backend.marg_window_batch_host(b, capi.RUN_BOTH, raw_imu=True, z_one=True, xy_f32=xyf)
    b.lm_obs[0:2] = keep
    for f in FIELDS:
        assert np.array_equal(getattr(outx, f), getattr(out1, f)), f
        assert np.array_equal(getattr(outh, f), getattr(out1, f)), f
        assert np.array_equal(getattr(outy, f), getattr(out0, f)), f
    # ABI 4: the prior records in and the recovered factors out WITHOUT their structural zeros (upper-triangular sqrt_info
    # blocks as 21 / 45 / 3 numbers, the symmetric covRel as 21): expanded / compacted on the device around the same kernels,
    # so after unpacking every bit is the one the full records give
    from is_vins_b200.batch import pack_tri_inputs, unpack_outputs
    tri = pack_tri_inputs(b)
    assert tri["prior_se3"].shape[1] == capi.SE3_TRI_REC and tri["prior_vb"].shape[1] == capi.VB_TRI_REC
    outp = backend.marg_window_batch_host(b, capi.RUN_BOTH, raw_imu=True, z_one=True, xy_f32=xyf, tri_in=tri, tri_out=True)
    assert outp.pg.shape[1] == capi.PG_TRI_REC and outp.rp.shape[1] == capi.RP_TRI_REC
    outu = unpack_outputs(outp)
    for f in FIELDS:
        if f == "pg":   # covRel = rpOmega^-1 is symmetric to rounding only: the packed form carries its upper triangle
            assert np.array_equal(outu.pg[:, :48], out1.pg[:, :48]) and np.array_equal(outu.pg[:, 84:], out1.pg[:, 84:])
            up = np.triu(np.ones((6, 6), bool)).flatten(order="F")
            assert np.array_equal(outu.pg[:, 48:84][:, up], out1.pg[:, 48:84][:, up])
            assert np.abs(outu.pg[:, 48:84] - out1.pg[:, 48:84]).max() <= 1e-12 * np.abs(out1.pg[:, 48:84]).max()
        else:
            assert np.array_equal(getattr(outu, f), getattr(out1, f)), f
    outq = unpack_outputs(backend.marg_window_batch_host(b, capi.RUN_BACKWARD, tri_in=tri, tri_out=True))   # record-IMU inputs
    assert np.array_equal(outq.rel, out0.rel) and np.array_equal(outq.vb, out0.vb) and np.array_equal(outq.rp, out0.rp)
    outf = unpack_outputs(backend.marg_window_batch_host(b, capi.RUN_FORWARD, tri_out=True))                # full records in
    assert np.array_equal(outf.se3, out0.se3) and np.array_equal(outf.pg[:, :48], out0.pg[:, :48])
    dbt = DeviceBatch(b, "cuda:0")
    dbt.flags = capi.IN_TRI_RECORDS                        # the packed forms belong to the host-pointer entry point only
    with pytest.raises(capi.IsvError):
        backend.marg_window_batch(dbt, capi.RUN_BOTH)
    with pytest.raises(ValueError):                        # values that are not floats are refused by the mirror
        bad = b.lm_obs.copy()
        bad[0, 3] += 1e-12
        xy_as_f32(bad)
    b.lm_obs[2, 0] = 2.0                                   # the promise is spot-checked on the host path
    with pytest.raises(capi.IsvError):
        backend.marg_window_batch_host(b, capi.RUN_BOTH, z_one=True)
