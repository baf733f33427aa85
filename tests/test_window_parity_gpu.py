"""-m gpu: the CUDA window kernels (through the C ABI) against the oracle on seeded windows.

Tolerance: north_star asks <= 1e-9 relative (FP64); stated here as relative Frobenius error per
recovered factor.  Discrete outcomes (ranks) are compared exactly.
"""
import numpy as np
import pytest

from is_vins_b200 import DeviceBatch, capi, pack_events
from oracle import sim
from tests.helpers import compare_event, expected_ranks

TOL = 1e-9
pytestmark = pytest.mark.gpu


def _events(config_id, batch_id, L, rounds):
    ch = sim.make_chain(sim.seed_for(config_id, batch_id), L=L, rounds=rounds)
    return ch.events


@pytest.mark.parametrize("L", [150, 1, 31, 32, 33, 1000])
def test_device_batch_matches_oracle(backend, L):
    events = _events(1, L % 7, L, 3)
    batch = pack_events(events)
    db = DeviceBatch(batch, "cuda:0")
    backend.marg_window_batch(db, capi.RUN_BOTH)
    backend.synchronize()
    out = db.outputs()
    for w, ev in enumerate(events):
        errs = compare_event(out, w, ev)
        worst = max(errs.values())
        assert worst <= TOL, (L, w, errs)
        rf, rb = expected_ranks(ev)
        assert (int(out.rank[w, 0]), int(out.rank[w, 1])) == (rf, rb)
        assert int(out.status[w]) == 0


def test_host_batch_and_single_window(backend):
    events = _events(1, 3, [150, 80, 200], 3)
    batch = pack_events(events)
    out = backend.marg_window_batch_host(batch, capi.RUN_BOTH)
    for w, ev in enumerate(events):
        assert max(compare_event(out, w, ev).values()) <= TOL
    # single-window wrappers = what Estimator::MargForward()/MargBackward() call
    ev = events[1]
    f, b = ev.fwd_in, ev.bwd_in
    se3, pg, rank, status = backend.marg_forward(f.pose0, f.pose1, f.ex_pose, f.inv_dep, f.pts_i, f.pts_j,
                                                 batch.prior_se3[1], batch.prior_rel[1], batch.prior_rp[1])
    assert status == 0 and rank == 6
    assert np.array_equal(se3, out.se3[1]) and np.array_equal(pg, out.pg[1])
    rel, vb, rp, rank, status = backend.marg_backward(b.pose_i, b.sb_i, b.pose_j, b.sb_j, batch.prior_vb[1],
                                                      batch.preint[1])
    assert status == 0 and rank == ev.bwd_out.rank
    assert np.array_equal(rel, out.rel[1]) and np.array_equal(vb, out.vb[1]) and np.array_equal(rp, out.rp[1])
