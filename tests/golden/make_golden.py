"""Generates the committed fixtures under tests/golden/ from the NumPy oracle (oracle/isv_oracle.py)
driven by the synthetic chain generator (oracle/sim.py).  Run from the repo root:

    python tests/golden/make_golden.py

The reference has no golden vectors for this path: these are the ORACLE's outputs
on seeded inputs (seed = 20260000 + 1000*config_id + batch_id), i.e. regression anchors for the
restatement and the inputs of bench.py, not reference-pinned truth.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is_vins_b200.batch import pack_events  # noqa: E402
from oracle import sim  # noqa: E402
from tests.helpers import oracle_outputs, save_batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def chain_events(config_id, batch_id, L, rounds, **kw):
    return sim.make_chain(sim.seed_for(config_id, batch_id), L=L, rounds=rounds, **kw).events


def main():
    # cfg 1: MH_01-shaped window, L = 150 (+ ragged feature counts incl. the warp-tile edges 31/32/33, 1 and 0)
    ev = chain_events(1, 0, 150, 3) + chain_events(1, 1, [0, 1, 31, 32, 33, 80], 6, max_gap=3)
    save_batch(os.path.join(HERE, "cfg1_L150_ragged.npz"), pack_events(ev), oracle_outputs(ev),
               {"seeds": np.array([sim.seed_for(1, 0), sim.seed_for(1, 1)])})
    # cfg 2: L = 1000, literal dense oracle (full-pivot LU of the 1006^2 block) on 2 windows
    ev = chain_events(2, 0, 1000, 2)
    save_batch(os.path.join(HERE, "cfg2_L1000_literal.npz"), pack_events(ev), oracle_outputs(ev),
               {"seeds": np.array([sim.seed_for(2, 0)])})
    # bench inputs: 8 base windows per shape; expected outputs from the structured oracle path
    # (identical to the literal one to ~1e-15, see tests/test_oracle_cpu.py)
    # (cfg 4a: L = 2000 > NUM_OF_F is impossible in the reference, SURVEY.md 8d -- the runtime-L engine takes it)
    only = os.environ.get("ISV_GOLDEN_ONLY")
    for cfg_id, L in ((2, 1000), (1, 150), (4, 2000)):
        if only and str(L) != only:
            continue
        ev = []
        for b in range(2):
            ev += chain_events(cfg_id, 10 + b, L, 4, structured=True)
        save_batch(os.path.join(HERE, f"bench_windows_L{L}.npz"), pack_events(ev), oracle_outputs(ev),
                   {"seeds": np.array([sim.seed_for(cfg_id, 10), sim.seed_for(cfg_id, 11)])})


if __name__ == "__main__":
    main()
