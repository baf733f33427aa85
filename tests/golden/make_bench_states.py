"""Generates tests/golden/bench_states_512.npz: the non-landmark inputs of 512 genuinely different MARGIN_OLD
events (64 seeded chains x 8 rounds of the oracle's estimator chain, oracle/sim.py, L = 150 per round so the
priors carry a realistic amount of visual information).  bench.py draws every window's state from this pool and
generates the landmark observations itself (count, image positions, depths, both observations) from the
window's own poses, so no two windows of a bench batch are numerically alike.  Run from the repo root:

    python tests/golden/make_bench_states.py          (~2 min)

Seeds: sim.seed_for(7, chain) = 20267000 + chain.  These are INPUTS only (no expected outputs are stored: bench.py
checks its results against oracle/isv_ref.c at run time, tests/test_bench_cpu.py pins the generator).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is_vins_b200.batch import pack_events  # noqa: E402
from oracle import sim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_CHAINS, ROUNDS = 64, 8


def main():
    ev = []
    for c in range(N_CHAINS):
        ch = sim.make_chain(sim.seed_for(7, c), L=150, rounds=ROUNDS, structured=True, with_yaw=(c % 2 == 1))
        ev += ch.events
    b = pack_events(ev)
    assert b.imu_raw is not None
    d = {f: getattr(b, f) for f in b.FIELDS if f not in ("lm_offset", "lm_obs")}
    d["seeds"] = np.array([sim.seed_for(7, c) for c in range(N_CHAINS)])
    np.savez_compressed(os.path.join(HERE, "bench_states_512.npz"), **d)
    print("windows", b.n, "rp_valid", int(b.prior_rp[:, 0].sum()))


if __name__ == "__main__":
    main()
