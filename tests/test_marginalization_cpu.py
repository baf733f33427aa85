"""CPU: the host-side bookkeeping of the `MarginalizationInfo` facade (block order = the bit-exact index contract of
the generic engine, drop sets, constant blocks, the td rule, prior hand-over) -- no device work."""
import numpy as np
import pytest

from is_vins_b200 import FactorProblem, MarginalizationInfo, PriorState, ResidualBlockInfo, add_margin_old_blocks
from is_vins_b200.marginalization import LOCAL_SIZE


def _problem(n_pose=5, n_feat=6):
    rng = np.random.default_rng(0)
    pidx, pobs = [], []
    for f in range(n_feat):
        host = 0 if f % 2 == 0 else 1
        for j in range(host + 1, min(host + 3, n_pose)):
            pidx.append((host, j, 0, f))
            pobs.append(rng.normal(size=5))
    z = lambda *s: np.zeros(s)
    return FactorProblem(z(n_pose, 7), z(n_pose, 9), z(1, 7), z(n_feat), np.array(pidx, np.int32).T.copy(),
                         np.array(pobs).T.copy(), np.array([(i, i + 1) for i in range(n_pose - 1)], np.int32),
                         z(n_pose - 1, 467))


def test_block_order_is_dense_then_scalars_then_kept_by_first_appearance():
    fp = _problem()
    mi = MarginalizationInfo(None, constant=[("ex_pose", 0)])
    blocks = add_margin_old_blocks(mi, fp)
    pos, m_dense, diag = mi.order_blocks()
    idx = mi.parameter_block_idx
    assert m_dense == 15 and idx[("pose", 0)] == 0 and idx[("speed_bias", 0)] == 6
    hosted = [f for f in range(6) if f % 2 == 0]                       # features hosted in frame 0, in factor order
    assert diag == [("feature", f) for f in hosted]
    assert [idx[("feature", f)] for f in hosted] == [15, 16, 17]
    assert ("ex_pose", 0) not in idx                                   # constant: no column
    assert not any(k == ("feature", 1) for k in idx)                   # hosted in frame 1: not part of this problem
    # kept blocks by first appearance: the IMU factor's pose 1 / speed-bias 1, then pose 2 from the first projection
    kept = sorted((v, k) for k, v in idx.items() if v >= mi.m)
    assert [k for _, k in kept] == [("pose", 1), ("speed_bias", 1), ("pose", 2)]
    assert mi.m == 18 and mi.n == 6 + 9 + 6 and pos == mi.m + mi.n
    assert [b.kind for b in blocks][:2] == ["imu", "projection"] and len(blocks) == 1 + 6
    keep = mi.getParameterBlocks()
    assert [k for k, _, _ in keep] == [("pose", 1), ("speed_bias", 1), ("pose", 2)]
    assert [i for _, _, i in keep] == [0, 6, 15] and [s for _, s, _ in keep] == [7, 9, 7]
    shift = {k: (k[0], k[1] - 1) for k, _, _ in keep}
    assert mi.getParameterBlocks(addr_shift=shift) == [("pose", 0), ("speed_bias", 0), ("pose", 1)]


def test_td_block_is_kept_and_cannot_be_dropped():
    fp = _problem()
    P = fp.proj_idx.shape[1]
    mi = MarginalizationInfo(None, tr_over_row=1e-4)                   # extrinsics estimated: ex_pose gets a column
    add_margin_old_blocks(mi, fp, td_obs=np.zeros((8, P)))
    pos, m_dense, diag = mi.order_blocks()
    idx = mi.parameter_block_idx
    assert idx[("td", 0)] >= mi.m and idx[("ex_pose", 0)] >= mi.m
    assert mi.n == 6 + 9 + 6 + 6 + 1
    with pytest.raises(ValueError):
        mi.addResidualBlockInfo(ResidualBlockInfo(
            "projection_td", [("pose", 0), ("pose", 1), ("ex_pose", 0), ("feature", 0), ("td", 0)], drop_set=[4],
            pts_i=np.ones(3), pts_j=np.ones(3), velocity_i=np.zeros(2), velocity_j=np.zeros(2), td_i=0.0, td_j=0.0,
            row_i=0.0, row_j=0.0))
    with pytest.raises(AssertionError):                                # wrong parameter-block families for the kind
        ResidualBlockInfo("imu", [("pose", 0), ("pose", 1), ("speed_bias", 0), ("speed_bias", 1)])


def test_prior_state_hands_over_the_kept_blocks():
    keys = [("pose", 0), ("speed_bias", 0), ("pose", 1), ("ex_pose", 0), ("td", 0)]
    n = sum(LOCAL_SIZE[k[0]] for k in keys)
    rng = np.random.default_rng(1)
    x0 = [rng.normal(size={"pose": 7, "speed_bias": 9, "ex_pose": 7, "td": 1}[k[0]]) for k in keys]
    prior = PriorState(keys, np.triu(rng.normal(size=(n, n))), rng.normal(size=n), x0)
    assert prior.n == n == 6 + 9 + 6 + 6 + 1
    assert [(k, s, i) for k, s, i in prior.getParameterBlocks()] == [
        (("pose", 0), 7, 0), (("speed_bias", 0), 9, 6), (("pose", 1), 7, 15), (("ex_pose", 0), 7, 21), (("td", 0), 1, 27)]
    fp = _problem()
    mi = MarginalizationInfo(None)
    blocks = add_margin_old_blocks(mi, fp, td_obs=np.zeros((8, fp.proj_idx.shape[1])), prior=prior)
    assert blocks[0].kind == "marginalization" and blocks[0].drop_set == [0, 1]       # the oldest frame's blocks
    mi.order_blocks()
    idx = mi.parameter_block_idx
    assert idx[("pose", 0)] == 0 and idx[("speed_bias", 0)] == 6
    # first appearance is now the prior's block order: pose 1, ex_pose, td come before the IMU factor's speed-bias 1
    kept = [k for _, k in sorted((v, k) for k, v in idx.items() if v >= mi.m)]
    assert kept[:4] == [("pose", 1), ("ex_pose", 0), ("td", 0), ("speed_bias", 1)]
    with pytest.raises(AssertionError):                                # blocks must match the prior's kept blocks
        ResidualBlockInfo("marginalization", keys[:-1], prior=prior)


def test_bad_drop_sets_are_rejected_before_any_state_changes():
    """drop_set indices are validated first; a scalar block the previous prior kept cannot be dropped into the
    diagonal block (its coupling through the prior would be ignored by the diagonal elimination)."""
    mi = MarginalizationInfo(None)
    with pytest.raises(ValueError):
        mi.addResidualBlockInfo(ResidualBlockInfo("se3", [("pose", 0)], drop_set=[1], t=np.zeros(3), R=np.eye(3),
                                                  sqrt_info=np.eye(6)))
    assert mi.factors == [] and mi.parameter_block_size == {}
    # a prior that kept feature 3, and a new round that wants to drop it as a scalar
    keys = [("pose", 0), ("feature", 3)]
    prior = PriorState(keys, np.eye(7), np.zeros(7), [np.zeros(7), np.zeros(1)])
    mi = MarginalizationInfo(None)
    mi.addResidualBlockInfo(ResidualBlockInfo("marginalization", keys, drop_set=[1], prior=prior))
    with pytest.raises(ValueError, match="diagonal"):
        mi.order_blocks()
