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
    ap.add_argument("--w20", action="store_true",
                    help="BASELINE configs[3] variant (b): WINDOW_SIZE = 20, ~2000 live features, ProjectionTdFactor, "
                         "previous prior over the whole window (tests/golden/problem_W20_F2000_td.npz): n_keep = 307")
    args = ap.parse_args()
    if args.reduced:
        return reduced(args)
    import torch
    from is_vins_b200 import FactorProblem, MargBackend, MarginalizationInfo, PriorState, ResidualBlockInfo, add_margin_old_blocks
    from is_vins_b200 import capi

    name = "problem_W20_F2000_td.npz" if args.w20 else "problem_F300_host0.npz"
    fp = FactorProblem.load(os.path.join(ROOT, "tests", "golden", name))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    be = MargBackend(0)
    be.use_torch_stream()
    if args.w20:
        return w20(args, be, fp, np.load(os.path.join(ROOT, "tests", "golden", name)))
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
    batch_bench(args, be, mi, "VINS-Mono style marginalization of the oldest frame", 6)


def w20(args, be, fp, z):
    from is_vins_b200 import MarginalizationInfo, PriorState, add_margin_old_blocks
    N = fp.pose.shape[0]
    keys = [k for i in range(N) for k in (("pose", i), ("speed_bias", i))] + [("ex_pose", 0), ("td", 0)]
    sizes = [7, 9] * N + [7, 1]
    x0 = np.split(z["prior_x0"], np.cumsum(sizes)[:-1])
    prior = PriorState(keys, z["prior_J"], z["prior_r0"], x0)
    mi = MarginalizationInfo(be, eps=1e-8, cauchy_a=1.0, tr_over_row=float(z["tr_over_row"][0]))
    add_margin_old_blocks(mi, fp, z["td_obs"], prior)
    mi.preMarginalize({"pose": fp.pose, "speed_bias": fp.speed_bias, "ex_pose": fp.ex_pose, "feature": fp.feature, "td": z["td"]})
    mi.marginalize(keep_tables=True)
    assert mi.status == 0 and mi.n == 307, (mi.status, mi.n)
    return batch_bench(args, be, mi, "BASELINE configs[3] (b): WINDOW_SIZE=20, ProjectionTdFactor, previous prior over the whole "
                                     "window, MARGIN_OLD", 0)


def w20_from_bench(be, problems, steps):
    """bench.py's configs[3] (b) line: the same measurement on an existing backend; returns the result dict (with the
    normal equations A, b of one problem for the CPU leg)."""
    import types
    from is_vins_b200 import FactorProblem
    name = os.path.join(ROOT, "tests", "golden", "problem_W20_F2000_td.npz")
    args = types.SimpleNamespace(problems=problems, steps=steps, emit=False, keep_A=True)
    return w20(args, be, FactorProblem.load(name), np.load(name))


def batch_bench(args, be, mi, what, rank_slack):
    """Replicates the one problem `mi` just solved NP times (factor table with problem ids 0..NP-1; values / block
    table shared, read-only) and times isv_build_normal_equations [+ isv_add_marg_prior] + isv_schur_eig."""
    import torch
    from is_vins_b200 import capi
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
    prior = None
    if mi._prior is not None:
        prior = mi._prior_struct(pos_of=lambda k: -1 if k in mi.constant else mi.parameter_block_idx[k])
        pres = mi._prior["d"]["res"]

    def build():
        capi.check(lib.isv_build_normal_equations(be.h, C.byref(gi), C.byref(go)), "isv_build_normal_equations")
        if prior is not None:
            capi.check(lib.isv_add_marg_prior(be.h, C.byref(prior), C.c_void_p(pres.data_ptr()), C.byref(gi), C.byref(go), -1),
                       "isv_add_marg_prior")

    def run():
        build()
        capi.check(lib.isv_schur_eig(be.h, C.byref(gi), C.byref(go), 0), "isv_schur_eig")
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    assert int(torch.count_nonzero(o["status"]).item()) == 0
    # eigenvalues of unconstrained directions sit at the rounding-noise floor, on either side of eps, and the
    # atomic summation order differs per problem: the rank may differ by those few directions
    assert int((o["rank"] - mi.rank).abs().max().item()) <= rank_slack, o["rank"]
    # problems differ only by the order of the FP64 atomics in A, amplified by the Schur cancellation
    last = o["A_red"][NP - 1].T.cpu().numpy()
    spread = float(np.linalg.norm(last - mi.A_red) / np.linalg.norm(mi.A_red))
    assert spread <= 1e-8, spread
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_ne = t_all = 0.0
    for _ in range(args.steps):
        e0.record()
        build()
        e1.record()
        run()
        e2.record()
        torch.cuda.synchronize()
        t_ne += e0.elapsed_time(e1)
        t_all += e1.elapsed_time(e2)
    t_ne /= args.steps
    t_all /= args.steps
    res = ({"metric": "problems_marginalized_per_s", "value": NP / (t_all * 1e-3), "unit": "problems/s",
                      "config": {"workload": f"{what}, {NP} independent problems", "pos": pos, "m_dense": gi.m_dense,
                                 "m_diag": gi.m_diag, "n_keep": n, "residual_blocks_per_problem": nf,
                                 "previous_prior": prior is not None},
                      "ms_per_step": t_all, "kernels_ms": {"normal equations (memset + ne_build_kernel [+ prior])": t_ne,
                                                           "schur_diag_dmma_kernel + marg_schur_eig_kernel": t_all - t_ne},
                      "rank": mi.rank})
    res["A_one"] = res["b_one"] = None
    if getattr(args, "keep_A", False):   # bench.py's CPU leg: the normal equations of one problem before the Schur stage consumes them
        build()
        torch.cuda.synchronize()
        res["A_one"], res["b_one"] = o["A"][0].T.cpu().numpy().copy(), o["b"][0].cpu().numpy().copy()
    res["m"] = gi.m_dense + gi.m_diag
    if getattr(args, "emit", True):
        res.pop("A_one"); res.pop("b_one")
        print(json.dumps(res))
        be.close()
    return res


if __name__ == "__main__":
    main()
