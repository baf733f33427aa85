"""-m gpu: initFactorGraph sparsification tail (src/estimator.cpp:745-1001) vs the oracle."""
import numpy as np
import pytest

from oracle import sim
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_init_sparsify_matches_oracle(backend):
    chains = [sim.make_chain(sim.seed_for(6, b), L=8, rounds=1) for b in range(3)]
    poses = np.array([c.init_in.poses for c in chains])
    sbs = np.array([c.init_in.sbs for c in chains])
    pre = np.array([[p.pack() for p in c.init_in.pres] for c in chains])
    out = backend.init_sparsify(poses, sbs, pre)
    for w, c in enumerate(chains):
        io = c.init_out
        assert int(out["rank"][w]) == io.rank == 42
        assert int(out["status"][w]) == 0
        for f in range(7):
            rec = out["rel"][w, f]
            assert rel_err(rec[0:3], io.rel_dt[f]) <= TOL
            assert rel_err(rec[3:12].reshape(3, 3).T, io.rel_dR[f]) <= TOL
            assert rel_err(rec[12:48].reshape(6, 6).T, io.rel_sqrt_info[f]) <= TOL, (w, f)
        assert rel_err(out["se3"][w, 0:3], io.se3_t) <= TOL
        assert rel_err(out["se3"][w, 3:12].reshape(3, 3).T, io.se3_R) <= TOL
        assert rel_err(out["se3"][w, 12:48].reshape(6, 6).T, io.se3_sqrt_info) <= TOL
        assert rel_err(out["vb"][w, 0:9], io.vb) <= TOL
        assert rel_err(out["vb"][w, 9:90].reshape(9, 9).T, io.vb_sqrt_info) <= TOL
