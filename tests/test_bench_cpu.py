"""CPU tests of bench.py's synthetic-window generator and of the comparison helpers it uses: every window of a
bench batch must be a different, valid MARGIN_OLD event (the old batch was 8 windows tiled)."""
import numpy as np

import bench
from is_vins_b200.batch import outputs_rel_diff
from oracle import ref_c


def test_bench_batch_windows_are_distinct_ragged_and_valid():
    L, n = 150, 1100                                  # > 2 x the 512-state pool: copies of a pool state must differ too
    b = bench.make_batch(L, n, 11)
    counts = np.diff(b.lm_offset)
    assert counts.min() >= round(0.75 * L) and counts.max() <= round(1.25 * L) and len(np.unique(counts)) > 40
    assert abs(counts.mean() - L) < 0.03 * L
    # unit quaternions (the structured kernels flag |q|^2 - 1 > 1e-9)
    for p in (b.pose_fwd, b.pose_bwd):
        assert np.abs(np.sum(p[..., 3:7] ** 2, axis=-1) - 1.0).max() < 1e-14
    # no two windows share a state or an observation
    assert len(np.unique(b.pose_fwd.reshape(n, -1), axis=0)) == n
    assert len(np.unique(b.prior_se3, axis=0)) == n and len(np.unique(b.prior_vb, axis=0)) == n
    assert len(np.unique(b.lm_obs[0:2].T, axis=0)) == b.n_landmarks and len(np.unique(b.lm_obs[5])) == b.n_landmarks
    # pts_i.x / pts_i.y are cv::Point2f values in the reference (feature_tracker_simple.h:55): FP32-representable here too,
    # which is what lets the e2e path ship them as floats (ABI 3) without changing a bit of the arithmetic
    assert np.array_equal(b.lm_obs[0:2].astype(np.float32).astype(np.float64), b.lm_obs[0:2])
    # landmarks are in front of both cameras and inside a sane field of view
    assert np.all(b.lm_obs[5] > 0.1) and np.all(b.lm_obs[5] < 1.1) and np.all(b.lm_obs[2] == 1.0)
    assert np.abs(b.lm_obs[3:5]).max() < 3.0
    # the C oracle accepts every window (status 0, full rank) and its two algorithms agree on a sample
    out = ref_c.marg_window_batch(b, 3, 0, True)
    assert not out.status.any() and np.all(out.rank == [6, 15])
    assert len(np.unique(out.se3, axis=0)) == n and len(np.unique(out.vb, axis=0)) == n
    idx = np.arange(0, n, 50)
    lit = ref_c.marg_window_batch(b.take(idx), 3, 0, False)
    sub = type(out)(*[getattr(out, f)[idx] for f in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
    assert outputs_rel_diff(sub, lit, 3).max() <= 1e-9
    # same seed -> same batch (ranks regenerate their shard deterministically)
    b2 = bench.make_batch(L, n, 11)
    assert np.array_equal(b.lm_obs, b2.lm_obs) and np.array_equal(b.prior_rel, b2.prior_rel)
    assert not np.array_equal(bench.make_batch(L, 64, 12).lm_obs[:, :64], b.lm_obs[:, :64])


def test_take_and_slice_agree():
    b = bench.make_batch(40, 20, 3)
    s, t = b.slice(5, 9), b.take([5, 6, 7, 8])
    for f in b.FIELDS:
        x, y = getattr(s, f), getattr(t, f)
        assert (x is None and y is None) or np.array_equal(x, y), f
    r = b.take([7, 2])
    a7, b7 = int(b.lm_offset[7]), int(b.lm_offset[8])
    assert np.array_equal(r.lm_obs[:, : b7 - a7], b.lm_obs[:, a7:b7]) and np.array_equal(r.pose_bwd[1], b.pose_bwd[2])


def test_reference_and_cuda_arms_describe_the_same_config():
    """The driver compares the two arms' `config` objects key by key."""
    assert bench.workload_config(1000, 9472) == bench.workload_config(1000, 9472)
    assert "configs[1]" in bench.workload_config(1000, 9472)["workload"]
    assert "configs[2]" in bench.workload_config(150, 9472, 2)["workload"]
