#!/usr/bin/env python
"""Throughput of the generic marginalization engine (`isv_marginalize_generic`, the `MarginalizationInfo`
back end): VINS-Mono-style marginalization of the oldest frame (pose + speed-bias + every feature hosted in
it) of the committed problem tests/golden/problem_F300_host0.npz, batched over independent problems.

    python tools/bench_marg_generic.py [--problems 296] [--steps 10]

Prints one JSON line: problems/s, the per-kernel times and the shapes (pos, m_dense, m_diag, n).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=296)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from is_vins_b200 import FactorProblem, MargBackend, MarginalizationInfo, ResidualBlockInfo
    from is_vins_b200 import capi
    from is_vins_b200.marginalization import isv_ne_factor

    fp = FactorProblem.load(os.path.join(ROOT, "tests", "golden", "problem_F300_host0.npz"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    be = MargBackend(0)
    be.use_torch_stream()
    mi = MarginalizationInfo(be, eps=1e-8, cauchy_a=1.0, constant=[("ex_pose", 0)])
    mi.addResidualBlockInfo(ResidualBlockInfo("imu", [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)],
                                              drop_set=[0, 1], preint=fp.imu_preint[0]))
    for k in range(fp.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in fp.proj_idx[:, k]]
        if i == 0:
            mi.addResidualBlockInfo(ResidualBlockInfo(
                "projection", [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)], drop_set=[0, 3],
                pts_i=fp.proj_obs[0:3, k], pts_j=np.array([fp.proj_obs[3, k], fp.proj_obs[4, k], 1.0])))
    rec = fp.se3_rec[0]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", [("pose", 0)], drop_set=[0], t=rec[0:3], R=rec[3:12].reshape(3, 3).T,
                                              sqrt_info=rec[12:].reshape(6, 6).T))
    rec = fp.rel_rec[0]
    mi.addResidualBlockInfo(ResidualBlockInfo("rel", [("pose", 0), ("pose", 1)], drop_set=[0], delta_t=rec[0:3],
                                              delta_R=rec[3:12].reshape(3, 3).T, sqrt_info=rec[12:].reshape(6, 6).T))
    mi.preMarginalize({"pose": fp.pose, "speed_bias": fp.speed_bias, "ex_pose": fp.ex_pose, "feature": fp.feature})
    mi.marginalize(keep_tables=True)
    assert mi.status == 0
    gi, tabs = mi._gi, mi._tables
    NP, nf = args.problems, gi.n_factors
    # replicate the factor table with problem ids 0..NP-1; values / block table are shared (read-only)
    fa = np.frombuffer(tabs["factors_bytes"], dtype=np.dtype([("res", "<i8"), ("nres", "<i4"), ("nb", "<i4"),
                                                               ("fb", "<i4"), ("prob", "<i4")])).copy()
    big = np.tile(fa, NP)
    big["prob"] = np.repeat(np.arange(NP, dtype=np.int32), nf)
    dev = "cuda:0"
    d_f = torch.from_numpy(big.view(np.uint8)).to(dev)
    pos, n = gi.pos, mi.n
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device=dev)
    o = {"A": z(NP, pos, pos), "b": z(NP, pos), "A_red": z(NP, n, n), "b_red": z(NP, n), "J": z(NP, n, n), "r": z(NP, n),
         "rank": torch.zeros((NP,), dtype=torch.int32, device=dev), "status": torch.zeros((NP,), dtype=torch.int32, device=dev)}
    gi.n_problems, gi.n_factors, gi.factors = NP, NP * nf, d_f.data_ptr()
    go = type(mi._go)(o["A"].data_ptr(), o["b"].data_ptr(), o["A_red"].data_ptr(), o["b_red"].data_ptr(), o["J"].data_ptr(),
                      o["r"].data_ptr(), o["rank"].data_ptr(), o["status"].data_ptr())
    lib = be.lib
    for _ in range(2):
        capi.check(lib.isv_marginalize_generic(be.h, C.byref(gi), C.byref(go)), "isv_marginalize_generic")
    torch.cuda.synchronize()
    assert int(torch.count_nonzero(o["status"]).item()) == 0
    # eigenvalues of unconstrained directions sit at the rounding-noise floor, on either side of eps, and the
    # atomic summation order differs per problem: the rank may differ by those few directions
    assert int((o["rank"] - mi.rank).abs().max().item()) <= 6, o["rank"]
    assert torch.allclose(o["A_red"][NP - 1].T.cpu(), torch.from_numpy(mi.A_red), rtol=1e-9, atol=1e-6)
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_ne = t_all = 0.0
    for _ in range(args.steps):
        e0.record()
        capi.check(lib.isv_build_normal_equations(be.h, C.byref(gi), C.byref(go)), "isv_build_normal_equations")
        e1.record()
        capi.check(lib.isv_marginalize_generic(be.h, C.byref(gi), C.byref(go)), "isv_marginalize_generic")
        e2.record()
        torch.cuda.synchronize()
        t_ne += e0.elapsed_time(e1)
        t_all += e1.elapsed_time(e2)
    t_ne /= args.steps
    t_all /= args.steps
    print(json.dumps({"metric": "problems_marginalized_per_s", "value": NP / (t_all * 1e-3), "unit": "problems/s",
                      "config": {"workload": "VINS-Mono style marginalization of the oldest frame, "
                                             f"{NP} independent problems", "pos": pos, "m_dense": gi.m_dense,
                                 "m_diag": gi.m_diag, "n_keep": n, "residual_blocks_per_problem": nf},
                      "ms_per_step": t_all, "kernels_ms": {"ne_build_kernel(+memset)": t_ne,
                                                           "marg_schur_eig_kernel": t_all - t_ne},
                      "rank": mi.rank}))
    be.close()


if __name__ == "__main__":
    main()
