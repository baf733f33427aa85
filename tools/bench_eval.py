#!/usr/bin/env python
"""Roofline of the batched ceres-Evaluate kernels (isv_eval_*_batch): factors evaluated / s and
achieved HBM GB/s (algorithmic bytes / CUDA-event time) against MEASURED_PEAKS.json.

    python tools/bench_eval.py [--windows 256] [--steps 20] [--no-ex-jac]

Algorithmic bytes per ProjectionFactor: in 4*4 (indices) + 5*8 (observation) = 56 B,
out 8*(2 + 14 + 14 [+ 14] + 2) = 256 B (368 B with the extrinsic block); the gathered parameter blocks
are shared by ~10^3 factors and served by L2.  Inputs + outputs exceed the 126 MB L2.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-ex-jac", action="store_true")
    ap.add_argument("--cauchy", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    from is_vins_b200 import DeviceProblem, FactorProblem, MargBackend, eval_problem
    # committed synthetic problemSolve() factor list (tests/golden/make_problem_golden.py, seed 20266100)
    fp = FactorProblem.load(os.path.join(ROOT, "tests", "golden", "problem_F1000.npz")).tile(args.windows)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    be = MargBackend(0)
    be.use_torch_stream()
    dp = DeviceProblem(fp, "cuda:0", want_ex_jac=not args.no_ex_jac)
    P, ni = dp.n_proj, dp.n_imu
    for _ in range(3):
        eval_problem(be, dp, args.cauchy)
    torch.cuda.synchronize()
    assert int(dp.status.item()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = be.launch_count
    e0.record()
    for _ in range(args.steps):
        eval_problem(be, dp, args.cauchy)
    e1.record()
    torch.cuda.synchronize()
    ms_all = e0.elapsed_time(e1) / args.steps
    launches = be.launch_count - l0
    # projection kernel alone
    import ctypes as C
    from is_vins_b200 import capi
    pb = dp.param_blocks()
    pf = capi.isv_proj_factors(P, P, dp.t["proj_idx"].data_ptr(), dp.t["proj_obs"].data_ptr(), args.cauchy)
    po = capi.isv_proj_eval(dp._o("proj_res"), dp._o("proj_ji"), dp._o("proj_jj"), dp._o("proj_je"), dp._o("proj_jf"))
    e0.record()
    for _ in range(args.steps):
        be.lib.isv_eval_projection_batch(be.h, C.byref(pb), C.byref(pf), C.byref(po), None)
    e1.record()
    torch.cuda.synchronize()
    ms_proj = e0.elapsed_time(e1) / args.steps
    out_b = 8 * (2 + 14 + 14 + 2 + (0 if args.no_ex_jac else 14))
    alg = P * (56 + out_b)
    peak = 6650.0
    src = "fallback (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak, src = float(json.load(open(pk))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    ach = alg / (ms_proj * 1e-3) / 1e9
    print(json.dumps({
        "metric": "factors_evaluated_per_s", "value": (P + ni) / (ms_all * 1e-3), "unit": "factors/s",
        "config": {"workload": f"problemSolve() factor list, {args.windows} windows x 1000 features "
                               f"({P} ProjectionFactors, {ni} IMUFactors), Cauchy a={args.cauchy}, "
                               f"ex-pose block {'constant' if args.no_ex_jac else 'evaluated'}",
                   "l2": "inputs_larger_than_l2"},
        "ms_per_step": ms_all, "gpu_launches": launches,
        "kernels_ms": {"eval_projection_kernel": ms_proj},
        "roofline": {"kernel": "eval_projection_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "peak_source": src, "alg_bytes_per_factor": 56 + out_b, "traffic": None},
    }))
    be.close()


if __name__ == "__main__":
    main()
