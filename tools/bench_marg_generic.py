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


def reduced(args):
    import torch
    from is_vins_b200 import FactorProblem, MargBackend, MarginalizationInfo, ResidualBlockInfo, capi
    fp = FactorProblem.load(os.path.join(ROOT, "tests", "golden", "problem_F1000.npz"))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    be = MargBackend(0)
    be.use_torch_stream()
    mi = MarginalizationInfo(be, eps=1e-8, cauchy_a=1.0, constant=[("ex_pose", 0)])
    for k in range(fp.proj_idx.shape[1]):
        i, j, e, f = [int(x) for x in fp.proj_idx[:, k]]
        mi.addResidualBlockInfo(ResidualBlockInfo(
            "projection", [("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)], drop_set=[3],
            pts_i=fp.proj_obs[0:3, k], pts_j=np.array([fp.proj_obs[3, k], fp.proj_obs[4, k], 1.0])))
    for k, (i, j) in enumerate(fp.imu_idx):
        mi.addResidualBlockInfo(ResidualBlockInfo(
            "imu", [("pose", int(i)), ("speed_bias", int(i)), ("pose", int(j)), ("speed_bias", int(j))],
            preint=fp.imu_preint[k]))
    for k, (i, j) in enumerate(fp.rel_idx):
        rec = fp.rel_rec[k]
        mi.addResidualBlockInfo(ResidualBlockInfo("rel", [("pose", int(i)), ("pose", int(j))], delta_t=rec[0:3],
                                                  delta_R=rec[3:12].reshape(3, 3).T, sqrt_info=rec[12:].reshape(6, 6).T))
    rec = fp.se3_rec[0]
    mi.addResidualBlockInfo(ResidualBlockInfo("se3", [("pose", int(fp.se3_idx[0]))], t=rec[0:3], R=rec[3:12].reshape(3, 3).T,
                                              sqrt_info=rec[12:].reshape(6, 6).T))
    rec = fp.vb_rec[0]
    mi.addResidualBlockInfo(ResidualBlockInfo("vb", [("speed_bias", int(fp.vb_idx[0]))], VB=rec[0:9],
                                              sqrt_info=rec[9:].reshape(9, 9).T))
    mi.preMarginalize({"pose": fp.pose, "speed_bias": fp.speed_bias, "ex_pose": fp.ex_pose, "feature": fp.feature})
    mi.marginalize(keep_tables=True, schur_only=True)
    assert mi.status == 0
    gi, tabs = mi._gi, mi._tables
    NP, nf = min(args.problems, 148), gi.n_factors
    fa = np.frombuffer(tabs["factors_bytes"], dtype=np.dtype([("res", "<i8"), ("nres", "<i4"), ("nb", "<i4"),
                                                               ("fb", "<i4"), ("prob", "<i4")])).copy()
    big = np.tile(fa, NP)
    big["prob"] = np.repeat(np.arange(NP, dtype=np.int32), nf)
    dev = "cuda:0"
    d_f = torch.from_numpy(big.view(np.uint8)).to(dev)
    pos, n = gi.pos, mi.n
    z = lambda *s: torch.zeros(s, dtype=torch.float64, device=dev)
    o = {"A": z(NP, pos, pos), "b": z(NP, pos), "A_red": z(NP, n, n), "b_red": z(NP, n),
         "rank": torch.zeros((NP,), dtype=torch.int32, device=dev), "status": torch.zeros((NP,), dtype=torch.int32, device=dev)}
    gi.n_problems, gi.n_factors, gi.factors = NP, NP * nf, d_f.data_ptr()
    go = type(mi._go)(o["A"].data_ptr(), o["b"].data_ptr(), o["A_red"].data_ptr(), o["b_red"].data_ptr(), None, None,
                      o["rank"].data_ptr(), o["status"].data_ptr())
    lib = be.lib
    for _ in range(2):
        capi.check(lib.isv_reduced_system(be.h, C.byref(gi), C.byref(go)), "isv_reduced_system")
    torch.cuda.synchronize()
    assert int(torch.count_nonzero(o["status"]).item()) == 0
    last = o["A_red"][NP - 1].T.cpu().numpy()
    assert float(np.linalg.norm(last - mi.A_red) / np.linalg.norm(mi.A_red)) <= 1e-9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        capi.check(lib.isv_reduced_system(be.h, C.byref(gi), C.byref(go)), "isv_reduced_system")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    nt = (n + 63) // 64
    flops = NP * (nt * (nt + 1) // 2) * 64 * 64 * gi.m_diag * 2.0
    print(json.dumps({"metric": "reduced_camera_systems_per_s", "value": NP / (ms * 1e-3), "unit": "windows/s",
                      "config": {"workload": f"problemSolve() normal equations + Schur over all features, {NP} windows",
                                 "pos": pos, "m_diag": gi.m_diag, "n_keep": n, "residual_blocks_per_window": nf},
                      "ms_per_step": ms, "schur_dmma_flops_executed": flops}))
    be.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=296)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--reduced", action="store_true",
                    help="reduced camera system of a whole 18-frame window (problem_F1000.npz: every factor, all 1000 "
                         "features eliminated, isv_reduced_system) instead of the oldest-frame marginalization")
    args = ap.parse_args()
    if args.reduced:
        return reduced(args)
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
    # problems differ only by the order of the FP64 atomics in A, amplified by the Schur cancellation
    last = o["A_red"][NP - 1].T.cpu().numpy()
    spread = float(np.linalg.norm(last - mi.A_red) / np.linalg.norm(mi.A_red))
    assert spread <= 1e-8, spread
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
