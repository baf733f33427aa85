import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from is_vins_b200 import MargBackend, capi
from oracle import sim, isv_oracle as O
from tests.helpers import rel_err
from tests.test_linalg_gpu import _run
be = MargBackend(0)
evs = sim.make_chain(sim.seed_for(1, 11), L=150, rounds=4, structured=True).events
A = np.array([ev.bwd_out.Lamda_prior for ev in evs])
A = 0.5 * (A + np.transpose(A, (0, 2, 1)))
G, lam, info = _run(be, A)
for w, ev in enumerate(evs):
    bo = ev.bwd_out
    r = info[w, 0]
    Gb = G[w][:r]; lm = lam[w][:r]
    keep = lm > 0.1
    Sigma = (Gb[keep].T / lm[keep] ** 2) @ Gb[keep]
    rel = O.RelativePoseFactor(bo.rel_dt, bo.rel_dR); rel.EvaluateOnlyJacobians(ev.bwd_in.pose_i, ev.bwd_in.pose_j)
    J = np.zeros((6, 21)); J[:, 15:21] = rel.jacobians[0]; J[:, 0:6] = rel.jacobians[1]
    s = O.llt_upper(np.linalg.inv(J @ Sigma @ J.T))
    GG = Gb @ Gb.T
    off = np.abs(GG - np.diag(np.diag(GG)))
    relorth = (off / np.sqrt(np.outer(np.diag(GG), np.diag(GG)))).max()
    print(w, "rows", r, "sweeps", info[w, 1], "kept", keep.sum(), "err vs oracle %.2e" % rel_err(s, bo.rel_sqrt_info),
          "max rel non-orth %.2e" % relorth, "recon %.2e" % (np.linalg.norm(Gb.T @ Gb - A[w]) / np.linalg.norm(A[w])), "lam small", np.sort(lm)[:3])
