"""-m gpu: batched IMU pre-integration (integration_base.h:30-158) vs the oracle's IntegrationBase,
and MargBackward fed by the device-computed records (BASELINE configs[2]: pre-integration of the
backward interval included)."""
import numpy as np
import pytest

from is_vins_b200 import capi, pack_events
from oracle import sim
from tests.helpers import compare_event, rel_err

pytestmark = pytest.mark.gpu


def test_preintegration_matches_oracle_and_feeds_backward(backend):
    events = []
    for b in range(3):
        events += sim.make_chain(sim.seed_for(3, b), L=12, rounds=2, max_gap=1 + b).events
    K = max(e.raw_imu.shape[0] for e in events)
    raw = np.zeros((len(events), K, 7))
    kc = np.zeros(len(events), np.int32)
    init = np.zeros((len(events), 12))
    for w, e in enumerate(events):
        k = e.raw_imu.shape[0]
        raw[w, :k] = e.raw_imu
        kc[w] = k
        init[w] = np.concatenate([e.acc0, e.gyr0, e.bwd_in.pre.linearized_ba, e.bwd_in.pre.linearized_bg])
    rec = backend.preintegrate(raw, init, kc)
    for w, e in enumerate(events):
        ref = e.bwd_in.pre.pack()
        assert rel_err(rec[w, 0:17], ref[0:17]) <= 1e-13                 # delta_p/q/v, biases, sum_dt
        assert rel_err(rec[w, 17:242], ref[17:242]) <= 1e-12             # jacobian
        assert rel_err(rec[w, 242:467], ref[242:467]) <= 1e-11           # covariance
    batch = pack_events(events)
    batch.preint = rec
    out = backend.marg_window_batch_host(batch, capi.RUN_BACKWARD)
    for w, e in enumerate(events):
        assert max(compare_event(out, w, e, which=2).values()) <= 1e-9
        assert int(out.rank[w, 1]) == e.bwd_out.rank and int(out.status[w]) == 0
