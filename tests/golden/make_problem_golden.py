"""Generates tests/golden/problem_*.npz: whole problemSolve() factor lists (oracle/sim.make_problem) in the
C ABI's array layout (is_vins_b200.FactorProblem) plus the oracle's ceres-Evaluate outputs of the first
factors of every kind.  Inputs of tools/bench_eval.py / tools/bench_marg_generic.py (which must not import
oracle/) and regression anchors for tests/test_evaluate_gpu.py.  Run from the repo root:

    python tests/golden/make_problem_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from is_vins_b200 import FactorProblem  # noqa: E402
from is_vins_b200.evaluate import DeviceProblem  # noqa: E402
from oracle import sim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, seed, td=False, **kw):
    p = sim.w20_problem() if td else sim.make_problem(seed, **kw)
    fp = FactorProblem.from_factors(p)
    d = {k: getattr(fp, k) for k in DeviceProblem.IN}
    if td:    # ProjectionTdFactor members + a steady-state previous prior over the whole window
        t = sim.make_td_observations(p, seed + 1)
        keys, J, r0, x0 = sim.make_window_prior(p, seed + 2)
        d.update(td_obs=t["td_obs"], td=t["td"], tr_over_row=np.array([t["tr_over_row"]]),
                 prior_J=J, prior_r0=r0, prior_x0=np.concatenate([np.ravel(v) for v in x0]))
    ref = sim.eval_problem_oracle(p, 1.0)
    n = min(64, len(ref["proj"]))
    d["ref_proj_res"] = np.array([r for r, _ in ref["proj"][:n]])
    d["ref_proj_ji"] = np.array([js[0].ravel() for _, js in ref["proj"][:n]])
    d["ref_proj_jj"] = np.array([js[1].ravel() for _, js in ref["proj"][:n]])
    d["ref_proj_je"] = np.array([js[2].ravel() for _, js in ref["proj"][:n]])
    d["ref_proj_jf"] = np.array([js[3].ravel() for _, js in ref["proj"][:n]])
    d["ref_imu_res"] = np.array([r for r, _ in ref["imu"]])
    d["ref_imu_jac"] = np.array([np.concatenate([j.ravel() for j in js]) for _, js in ref["imu"]])
    d["seed"] = np.array([seed])
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, "P =", fp.proj_idx.shape[1])


def main():
    save("problem_F1000.npz", sim.seed_for(6, 100), n_features=1000)                       # bench_eval workload
    save("problem_F300_host0.npz", sim.seed_for(9, 100), n_features=300, max_track=9, host0=0.8)  # marginalization
    save("problem_F60.npz", sim.seed_for(6, 101), n_features=60)                           # small regression anchor
    save("problem_W20_F2000_td.npz", sim.W20_SEED, td=True)                                    # configs[3]: 21 frames, td


if __name__ == "__main__":
    main()
