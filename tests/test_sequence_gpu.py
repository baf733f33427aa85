"""-m gpu: the device-resident sequence state (isv_seq_*) driven for R frames against the oracle's own
estimator chain (oracle/sim.make_chain: update -> double2vector -> MargForward/MargBackward ->
slideWindow rotation, src/estimator.cpp:1133-1144, :520-550, :1149-1539, :1605-1638) and the
pose-graph accumulator (CombinedFactors::operator+, pose_graph_factors.h:27-51).  Per frame the device
receives only states and observations; priors never leave the GPU.  Tolerance 1e-9 relative."""
import numpy as np
import pytest

from is_vins_b200 import SequenceState, WindowOutputs, capi, pack_events
from oracle import isv_oracle as O
from oracle import sim
from tests.helpers import compare_event, expected_ranks, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9
ROUNDS = 10          # > Vo_SIZE so that roll-pitch edges reach index 0, are consumed and erased


def _cm(rec9):
    return np.asarray(rec9).reshape(3, 3).T


def _check_state(st, q, V, ref):
    worst = 0.0
    t, R, s = ref["se3"]
    worst = max(worst, rel_err(st["se3"][q, 0:3], t), rel_err(_cm(st["se3"][q, 3:12]), R),
                rel_err(st["se3"][q, 12:].reshape(6, 6).T, s))
    vb, s = ref["vb"]
    worst = max(worst, rel_err(st["vb"][q, 0:9], vb), rel_err(st["vb"][q, 9:].reshape(9, 9).T, s))
    for i, (dt, dR, s) in ref["rel"].items():
        rec = st["rel"][i, q]
        worst = max(worst, rel_err(rec[0:3], dt), rel_err(_cm(rec[3:12]), dR), rel_err(rec[12:].reshape(6, 6).T, s))
    valid = {i for i in range(V) if st["rp_valid"][i, q]}
    assert valid == set(ref["rp"].keys()), (valid, ref["rp"].keys())
    for i, (R, s) in ref["rp"].items():
        rec = st["rp"][i, q]
        worst = max(worst, rel_err(_cm(rec[0:9]), R), rel_err(rec[9:].reshape(2, 2).T, s))
    return worst


def test_sequences_stay_on_device(backend):
    chains = [sim.make_chain(sim.seed_for(7, b), L=[25 + 7 * b + r for r in range(ROUNDS)], rounds=ROUNDS,
                             with_yaw=True) for b in range(3)]
    n, V = len(chains), chains[0].cfg.vo_size
    seq = SequenceState(backend, n)
    rank = seq.init(np.array([c.init_in.poses for c in chains]), np.array([c.init_in.sbs for c in chains]),
                    np.array([[p.pack() for p in c.init_in.pres] for c in chains]))
    assert list(rank) == [c.init_out.rank for c in chains]
    # oracle-side pose-graph accumulators
    accs = [O.CombinedFactors() for _ in chains]
    pg_index = [0] * n
    count = [0] * n
    worst = 0.0
    n_emitted = 0
    for r in range(ROUNDS):
        evs = [c.events[r] for c in chains]
        seq.update(np.array([e.upd["old_P"] for e in evs]), np.array([e.upd["old_R"] for e in evs]),
                   np.array([e.upd["old_vb"] for e in evs]), np.array([e.upd["new_pose"] for e in evs]),
                   np.array([e.upd["new_sb"] for e in evs]))
        rot = seq.yaw(np.array([e.upd["old_R"][0] for e in evs]), np.array([e.upd["new_pose"][0] for e in evs]))
        for q, e in enumerate(evs):
            assert rel_err(rot[q], e.rot_diff) <= 1e-12
        batch = pack_events(evs)
        batch.prior_se3[:] = np.nan      # must not be read: priors come from the device state
        batch.prior_rel[:] = np.nan
        batch.prior_vb[:] = np.nan
        flag, kf = seq.marginalize(batch, ts=np.array([e.pg_meta["ts"] for e in evs]),
                                   Ri=np.array([e.pg_meta["Ri"] for e in evs]),
                                   ti=np.array([e.pg_meta["ti"] for e in evs]), pg_cut_distance=0.1)
        st = seq.export()
        out = WindowOutputs(st["last_se3"], st["last_pg"], st["last_rel"], st["last_vb"], st["last_rp"],
                            st["last_rank"], st["last_status"])
        for q, e in enumerate(evs):
            errs = compare_event(out, q, e)
            worst = max(worst, max(errs.values()))
            assert (int(out.rank[q, 0]), int(out.rank[q, 1])) == expected_ranks(e)
            assert int(out.status[q]) == 0
            worst = max(worst, _check_state(st, q, V, e.state_after))
            # CombinedFactors chain on the oracle side
            fo = e.fwd_out
            cur = O.CombinedFactors()
            cur.relativePoseFactor = O.RelativePoseFactor(fo.pg_dt, fo.pg_dR)
            cur.relativePoseFactor.sqrt_info = fo.pg_sqrt_info
            cur.covRel, cur.distance = fo.pg_covRel, fo.pg_distance
            if e.pg_meta["rp"] is not None:
                cur.rollPitchFactor = O.RollPitchFactor(R=e.pg_meta["rp"][0])
                cur.rollPitchFactor.sqrt_info = e.pg_meta["rp"][1]
            cur.vio_index, cur.ts, cur.Ri, cur.ti = count[q], e.pg_meta["ts"], e.pg_meta["Ri"], e.pg_meta["ti"]
            count[q] += 1
            acc = accs[q] + cur
            emit = acc.distance > 0.1
            assert bool(flag[q]) == emit
            rec = kf[q] if emit else st["acc"][q]
            rp = acc.relativePoseFactor
            worst = max(worst, rel_err(rec[0:3], rp.delta_t), rel_err(_cm(rec[3:12]), rp.delta_R),
                        rel_err(rec[12:48].reshape(6, 6).T, rp.sqrt_info),
                        rel_err(rec[capi.ACC_COVREL:capi.ACC_COVREL + 36].reshape(6, 6).T, acc.covRel),
                        rel_err(rec[capi.ACC_DISTANCE], acc.distance),
                        rel_err(_cm(rec[capi.ACC_RI:capi.ACC_RI + 9]), acc.Ri), rel_err(rec[capi.ACC_TI:capi.ACC_TI + 3], acc.ti))
            assert int(rec[capi.ACC_LENGTH]) == acc.length and int(rec[capi.ACC_VIO_INDEX]) == acc.vio_index
            assert int(rec[capi.ACC_PG_INDEX]) == pg_index[q] and rec[capi.ACC_TS] == acc.ts
            assert bool(rec[capi.ACC_RP_VALID]) == (acc.rollPitchFactor is not None)
            if acc.rollPitchFactor is not None:
                worst = max(worst, rel_err(_cm(rec[capi.ACC_RP:capi.ACC_RP + 9]), acc.rollPitchFactor.R))
            if emit:
                n_emitted += 1
                pg_index[q] += 1
                accs[q] = O.CombinedFactors(pg_index=pg_index[q])
                assert int(st["acc"][q][capi.ACC_VIO_INDEX]) == -1 and int(st["acc"][q][capi.ACC_PG_INDEX]) == pg_index[q]
            else:
                accs[q] = acc
    assert n_emitted > 0, "no keyframe cut exercised"
    assert worst <= TOL, worst
    # checkpoint / resume: a restored state continues bit-identically
    snap = seq.export()
    seq2 = SequenceState(backend, n)
    seq2.restore(snap)
    again = seq2.export()
    for k in ("rel", "se3", "vb", "rp", "rp_valid", "acc", "pg_count"):
        assert np.array_equal(again[k], snap[k]), k
    seq.close()
    seq2.close()
