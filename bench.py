#!/usr/bin/env python
"""bench.py -- windows marginalized / s (BASELINE.json metric) for the B200-native IS-VINS backend.

A "step" = one pass of the hot path (MargForward + MargBackward, i.e. one MARGIN_OLD event per
window) over one batch of independent synthetic windows.  Headline workload at every N: BASELINE.json
configs[1] -- V=8, N=18, ~1000 features hosted in frame 0, K=10 IMU samples -- 9472 windows per GPU
(weak scaling: each rank owns its own contiguous shard, no collective on the data path).

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA path
  python bench.py --impl reference [...]                        the CPU restatement (oracle/isv_ref.c)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Every window of a batch is different: its state (poses, prior factors, pre-integration) is drawn from a pool of 512
seeded estimator-chain events (tests/golden/bench_states_512.npz) and perturbed per copy, its landmark count is
ragged (uniform in [0.75 L, 1.25 L], mean L) and its observations are generated from its own poses.

`value`    : whole-job windows/s with inputs resident in HBM (device pointers, C ABI isv_marg_window_batch), K steps
`sustained`: the same loop kept running for >= 2 s with clocks / power sampled inside it
`e2e`      : the same through isv_marg_window_batch_host: pinned HOST buffers, H2D + kernels + D2H per step
`parity`   : after the timed region, sampled windows of THIS batch are recomputed by oracle/isv_ref.c and compared
`roofline` : the landmark kernel (always the same kernel), algorithmic bytes / CUDA-event time vs measured HBM peak;
             `fp64` adds the FP64-pipe view; `per_kernel` lists every kernel of the step
`configs`  : the other BASELINE.json configs ([0] L=150, [2] MargBackward only, [3] L=2000 and the WINDOW_SIZE=20 td
             marginalization, [4] 4096 windows in total), each with value / e2e / latency / parity / cpu_baseline
`cpu_baseline`: oracle/isv_ref.c ("port": literal restatement of the reference's dense algorithm; the reference
             itself cannot be built here -- no Eigen/Ceres/Sophus) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "windows_marginalized_per_s"
UNIT = "windows/s"
RAGGED = 0.25            # landmark count per window ~ U{(1 - RAGGED) L .. (1 + RAGGED) L}
ROOFLINE_KERNEL = "marg_forward_accum_kernel"   # the landmark kernel: the only one that streams O(L) bytes per window


def alg_bytes(L, which=3, needed_only=False):
    """compulsory bytes per window.  fwd 56 L + 1992, bwd 5920 (sum 56 L + 7912).
    needed_only: IS-VINS' information-only marginalization never reads pts_j (no residual), so 4
    of the 7 doubles per landmark are algorithmically required -> fwd 32 L + 1992.
    The roofline uses this smaller figure."""
    b = 0
    if which & 1:
        b += (32 if needed_only else 56) * L + 8 * (117 + 132)
    if which & 2:
        b += 8 * (589 + 151)
    return b


# Per-kernel algorithmic work per window: bytes each kernel must move (inputs read once,
# outputs written once; the 42-double Gram hand-off between the two forward stages counted on both
# sides) and FP64 flops of the algorithm as implemented (FMA = 2).  L = mean landmarks per window of the batch.
KERNELS = {
    "marg_factor_jac_kernel": {
        "which": 16, "bytes": lambda L: 8 * (14 + 48 + 48 + 5 + 14 + 18 + 70) + 8 * (786 + 12 + 5 + 12 + 12 + 9 + 9),
        "flops": lambda L: 6.0e3,
        "limiter": "uncoalesced global access: one thread per (window, factor) walks a 384-3736 B strided record"},
    "marg_forward_accum_kernel": {
        # 265 flops per landmark with the isotropic ProjectionFactor::sqrt_info of the reference (the chain drops the 2x2
        # weighting: 282 in the general case); 42 of the DFMA are the two 6x6 SYRKs
        "which": 4, "bytes": lambda L: 32 * L + 8 * (14 + 7) + 8 * 42, "flops": lambda L: 265.0 * L,
        "limiter": "FP64 pipe (DFMA) / HBM stream of the landmark observations"},
    "marg_forward_tail_kernel": {
        "which": 8, "bytes": lambda L: 8 * (42 + 14 + 252) + 8 * (36 + 72) + 8, "flops": lambda L: 1.6e4,
        "limiter": "dependent-issue latency: one warp walks a chain of 6x6/12x12 steps"},
    "marg_backward_kernel": {
        "which": 32, "bytes": lambda L: 8 * (534 + 90 + 225 + 151), "flops": lambda L: 1.1e5,
        "limiter": "dependent-issue latency: one warp walks a chain (15x15 Cholesky, 24x30 Householder, 15x21 LQ)"},
}


# ------------------------------------------------------------------------------------------------
# synthetic windows
# ------------------------------------------------------------------------------------------------
FX, FY, CX, CY, IMG_W, IMG_H = 461.6, 460.3, 363.0, 248.1, 752, 480      # config/euroc_config.yaml:5-17


def _quat_R(q):
    """[n,4] (x, y, z, w) unit quaternions -> [n,3,3]"""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _jitter_pose(p, rng, sp, sq):
    """p [...,7]: position += N(0, sp^2), orientation *= exp(N(0, sq^2)), renormalised to 1 ulp"""
    sh = p.shape[:-1]
    p[..., 0:3] += rng.normal(0.0, sp, sh + (3,))
    th = rng.normal(0.0, sq, sh + (3,))
    x, y, z, w = [p[..., 3 + i].copy() for i in range(4)]
    dx, dy, dz, dw = 0.5 * th[..., 0], 0.5 * th[..., 1], 0.5 * th[..., 2], np.ones(sh)
    q = np.stack([w * dx + x * dw + y * dz - z * dy, w * dy - x * dz + y * dw + z * dx,
                  w * dz + x * dy - y * dx + z * dw, w * dw - x * dx - y * dy - z * dz], axis=-1)
    p[..., 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)


def make_batch(L, n_windows, seed, ragged=RAGGED, counts=None):
    """n_windows genuinely different synthetic windows of nominal landmark count L (see module docstring); `counts`
    overrides the per-window landmark counts."""
    from is_vins_b200.batch import WindowBatch
    z = np.load(os.path.join(ROOT, "tests", "golden", "bench_states_512.npz"))
    rng = np.random.default_rng(seed)
    pool = int(z["pose_fwd"].shape[0])
    src = (np.arange(n_windows) + int(rng.integers(0, pool))) % pool
    g = lambda k: np.ascontiguousarray(z[k][src])
    pose_fwd, prior_se3, prior_rel, prior_rp = g("pose_fwd"), g("prior_se3"), g("prior_rel"), g("prior_rp")
    pose_bwd, sb_bwd, prior_vb, preint = g("pose_bwd"), g("sb_bwd"), g("prior_vb"), g("preint")
    imu_raw, imu_init = g("imu_raw"), g("imu_init")
    ex = np.ascontiguousarray(z["ex_pose"])
    # every copy gets its own state: the estimate moves (as between two solver iterations) and the prior
    # informations are rescaled (upper-triangular square-root informations stay valid)
    _jitter_pose(pose_fwd, rng, 0.003, 0.0005)
    _jitter_pose(pose_bwd, rng, 0.003, 0.0005)
    sb_bwd += rng.normal(0.0, 1.0, sb_bwd.shape) * np.array([0.005] * 3 + [0.0005] * 3 + [0.00005] * 3)
    prior_se3[:, 12:] *= rng.uniform(0.8, 1.25, (n_windows, 1))
    prior_rel[:, 12:] *= rng.uniform(0.8, 1.25, (n_windows, 1))
    prior_vb[:, 9:] *= rng.uniform(0.8, 1.25, (n_windows, 1))
    prior_rp[:, 1:] *= rng.uniform(0.8, 1.25, (n_windows, 1))
    lo, hi = int(round((1.0 - ragged) * L)), int(round((1.0 + ragged) * L))
    if counts is None:
        counts = rng.integers(lo, hi + 1, n_windows) if hi > lo else np.full(n_windows, L)
    counts = np.asarray(counts, dtype=np.int64)
    off = np.zeros(n_windows + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    n_lm = int(off[-1])
    obs = np.empty((6, n_lm))
    # landmarks: uniform in the frame-0 image, depth U(1, 8) m, exact projection into frame 1 + 1 px noise on the
    # normalised plane, z == 1 (src/System.cpp:346), inverse depth = truth perturbed by 1 %
    ric = _quat_R(ex[None, 3:7])[0]
    tic = ex[0:3]
    CH = 256
    for w0 in range(0, n_windows, CH):
        w1 = min(w0 + CH, n_windows)
        a, b = int(off[w0]), int(off[w1])
        m = b - a
        if m == 0:
            continue
        wid = np.repeat(np.arange(w1 - w0), counts[w0:w1])
        R0, R1 = _quat_R(pose_fwd[w0:w1, 0, 3:7]), _quat_R(pose_fwd[w0:w1, 1, 3:7])
        P0, P1 = pose_fwd[w0:w1, 0, 0:3], pose_fwd[w0:w1, 1, 0:3]
        u, v = rng.uniform(0, IMG_W, m), rng.uniform(0, IMG_H, m)
        depth = rng.uniform(1.0, 8.0, m)
        # pts_i.x / pts_i.y are cv::Point2f values widened to double in the reference (feature_tracker_simple.h:55,
        # System.cpp:119-122): the synthetic observations are FP32-representable too
        pts_i = np.stack([((u - CX) / FX).astype(np.float32).astype(np.float64),
                          ((v - CY) / FY).astype(np.float32).astype(np.float64), np.ones(m)], axis=1)
        pb = (pts_i * depth[:, None]) @ ric.T + tic                          # body frame 0
        pw = np.einsum("nij,nj->ni", R0[wid], pb) + P0[wid]
        pb1 = np.einsum("nji,nj->ni", R1[wid], pw - P1[wid])                  # R1^T (pw - P1)
        pc1 = (pb1 - tic) @ ric
        obs[0:3, a:b] = pts_i.T
        obs[3, a:b] = pc1[:, 0] / pc1[:, 2] + rng.normal(0, 1.0 / 460.0, m)
        obs[4, a:b] = pc1[:, 1] / pc1[:, 2] + rng.normal(0, 1.0 / 460.0, m)
        obs[5, a:b] = (1.0 / depth) * (1.0 + rng.normal(0, 0.01, m))
    return WindowBatch(n_windows, off, obs, pose_fwd, ex, prior_se3, prior_rel, prior_rp, pose_bwd, sb_bwd, prior_vb,
                       preint, imu_raw, imu_init)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING a timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """t0, t1 (perf_counter): only the samples taken inside [t0, t1] count (default: all)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1):
                continue
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "power_w_median": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_fp64_peak():
    """DFMA / DMMA peak measured on THIS box by tools/fp64_peak (built by __graft_entry__.build()); the committed
    profile is the fallback."""
    best, src = {}, None
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    lines = []
    if os.path.exists(exe):
        try:
            lines = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.splitlines()
            src = "tools/fp64_peak run on this box inside this bench"
        except Exception:
            lines = []
    if not lines:
        for name in ("r02_fp64_peak.jsonl", "r01_fp64_peak.jsonl"):
            p = os.path.join(ROOT, "profiles", name)
            if os.path.exists(p):
                lines, src = open(p).read().splitlines(), f"committed profiles/{name}"
                break
    for line in lines:
        try:
            d = json.loads(line)
        except ValueError:
            continue
        if "kernel" in d:
            best[d["kernel"]] = max(best.get(d["kernel"], 0.0), float(d["tflops"]))
    return best, src


def read_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and per window from the committed `ncu --set full`
    capture of this bench (profiles/r02_traffic.json carries the commit the capture was taken on)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
    return {}


def traffic_for(tr, name, n, n_lm):
    """per-launch DRAM bytes of `name`, if the committed capture was taken on this workload (same windows, same landmark
    total), else None"""
    if tr.get("windows") == n and tr.get("landmarks", n_lm) == n_lm and name in tr:
        return tr[name]
    return None


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle/isv_ref.c): the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step(batch, threads, structured=False, which=3):
    from oracle import ref_c
    t0 = time.perf_counter()
    out = ref_c.marg_window_batch(batch, which, threads, structured)
    dt = time.perf_counter() - t0
    assert int(np.count_nonzero(out.status)) == 0
    return dt


def workload_config(L, n_per_gpu, which=3):
    name = {1000: "configs[1]", 150: "configs[0]", 2000: "configs[3] variant (a), no td"}.get(L, "custom L")
    if which == 2:
        name = "configs[2]"
    what = {3: "MargForward+MargBackward (oldest-frame marginalization + sparsification)",
            2: "MargBackward only (second-newest-frame marginalization: pose / speed-bias, no features)",
            1: "MargForward only"}[which]
    lo, hi = int(round((1 - RAGGED) * L)), int(round((1 + RAGGED) * L))
    return {"workload": f"BASELINE {name}: V=8 N=18 L~U{{{lo}..{hi}}} (mean {L}) K_imu=10, {what}, {n_per_gpu} "
                        f"distinct windows per GPU",
            "L": L, "L_ragged": [lo, hi], "vo_size": 8, "all_buf_size": 18, "k_imu": 10, "windows_per_gpu": n_per_gpu,
            "l2": "inputs_larger_than_l2" if n_per_gpu * alg_bytes(L, which) > 126e6 else "l2_flushed_between_steps",
            "parallelism": "independent windows sharded per GPU, no collective"}


def run_reference(args, L):
    from oracle import ref_c
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_c.build()
    threads = os.cpu_count() or 1
    sample_n = threads
    batch = make_batch(L, sample_n, 7)
    for _ in range(args.warmup):
        cpu_step(batch, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_step(batch, threads)
    value = sample_n * args.steps / t
    sample = (f"{sample_n} windows/step of the same workload (one per thread), literal dense algorithm (FullPivLU of the "
              f"(L+6)^2 block), oracle/isv_ref.c, OpenMP over windows")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(L, args.windows),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def parity_check(batch, out, which, n_sample, n_literal, threads):
    """Recompute sampled windows of `batch` with oracle/isv_ref.c and compare them with the CUDA outputs `out` of the
    timed run.  Structured mode (the O(L) landmark elimination) for n_sample evenly spaced windows, the literal dense
    algorithm of the reference for n_literal of them."""
    from is_vins_b200.batch import outputs_rel_diff
    from oracle import ref_c
    n = batch.n
    idx = np.unique(np.linspace(0, n - 1, min(n_sample, n)).round().astype(np.int64))
    sub = batch.take(idx)
    got = type(out)(*[getattr(out, f)[idx] for f in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
    ref = ref_c.marg_window_batch(sub, which, threads, True)
    err = outputs_rel_diff(got, ref, which)
    cols = [c for c, bit in ((0, 1), (1, 2)) if which & bit]
    res = {"checker": "oracle/isv_ref.c, structured landmark elimination", "windows_checked": int(len(idx)),
           "max_rel_err": float(np.nanmax(err)), "worst_window": int(idx[int(np.nanargmax(err))]),
           "median_rel_err": float(np.nanmedian(err)), "nonfinite": int(np.count_nonzero(~np.isfinite(err))),
           "ranks_equal": bool(np.array_equal(got.rank[:, cols], ref.rank[:, cols])),
           "status_nonzero": int(np.count_nonzero(got.status)), "tolerance": 1e-9}
    if n_literal > 0 and which & 1:
        li = idx[np.unique(np.linspace(0, len(idx) - 1, min(n_literal, len(idx))).round().astype(np.int64))]
        subl = batch.take(li)
        gotl = type(out)(*[getattr(out, f)[li] for f in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
        refl = ref_c.marg_window_batch(subl, which, threads, False)
        el = outputs_rel_diff(gotl, refl, which)
        res["literal"] = {"checker": "oracle/isv_ref.c, literal dense Lamda + FullPivLU (the reference's algorithm)",
                          "windows_checked": int(len(li)), "max_rel_err": float(np.nanmax(el)),
                          "ranks_equal": bool(np.array_equal(gotl.rank[:, cols], refl.rank[:, cols]))}
    res["ok"] = bool(res["max_rel_err"] <= 1e-9 and res["nonfinite"] == 0 and res["ranks_equal"]
                     and res["status_nonzero"] == 0 and res.get("literal", {}).get("max_rel_err", 0.0) <= 1e-9)
    return res


def measure(cx, L, n, which, steps, warmup, seed, full, e2e=True):
    """One workload on this rank's GPU: device-timed K steps, [per-kernel, sustained, latency], e2e; returns the raw
    timings (ms, this rank) plus what rank 0 needs for the report."""
    import torch
    from is_vins_b200 import DeviceBatch, capi
    from is_vins_b200.batch import WindowOutputs
    be, dev = cx.be, cx.dev
    batch = make_batch(L, n, seed)
    dev_fmt = os.environ.get("ISV_BENCH_DEVICE_INPUTS", "abi1")   # experiment switch: abi1 | zone | abi3 (z_one + xy_f32)
    db = DeviceBatch(batch, dev, z_one=dev_fmt in ("zone", "abi3"), xy_f32=dev_fmt == "abi3")
    r = {"batch": batch, "n": n, "n_lm": batch.n_landmarks, "L": L, "which": which}
    flush = None
    if n * alg_bytes(L, which) <= 126e6:
        if cx.flush is None:
            cx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush = cx.flush
    r["flushed"] = flush is not None

    for _ in range(warmup):
        if flush is not None:
            flush.fill_(1)
        be.marg_window_batch(db, which)
    cx.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    clocks = ClockSampler(cx.local).start() if (full and cx.rank == 0) else None
    l0 = be.launch_count
    t_w0 = time.perf_counter()
    if flush is None:
        ev[0].record()
        for _ in range(steps):
            be.marg_window_batch(db, which)
        ev[1].record()
        cx.barrier()
        ms_total = ev[0].elapsed_time(ev[1])
    else:  # L2 flush between steps: time each step on its own
        ms_total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            ev[0].record()
            be.marg_window_batch(db, which)
            ev[1].record()
            torch.cuda.synchronize()
            ms_total += ev[0].elapsed_time(ev[1])
        cx.barrier()
    t_w1 = time.perf_counter()
    r["launches"] = be.launch_count - l0
    r["ms_total"] = ms_total
    if clocks is not None:
        r["clocks"] = clocks.stop()
    assert int(torch.count_nonzero(db.out["status"]).item()) == 0, "a window reported a status flag"
    r["out"] = db.outputs() if cx.rank == 0 else None     # results of the timed run (for the parity sample)

    if full and cx.sustain_s <= 0:
        r["sustained_ms"], r["sustained_steps"], r["sustained_2nd_half_ms_per_step"] = ms_total, steps, ms_total / steps
    if full and cx.sustain_s > 0:
        # ---- sustained: the same loop for >= cx.sustain_s seconds, clocks and power sampled inside ------------------
        chunk = max(8, int(0.05 / max(ms_total / steps * 1e-3, 1e-6)))    # ~50 ms of work per chunk
        sclk = ClockSampler(cx.local).start() if cx.rank == 0 else None
        marks = [torch.cuda.Event(enable_timing=True)]
        cx.barrier()
        ts0 = time.perf_counter()
        marks[0].record()
        n_steps = 0
        while True:
            for _ in range(chunk):
                be.marg_window_batch(db, which)
            n_steps += chunk
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)
            if len(marks) > 2:
                marks[-2].synchronize()          # keep one chunk queued ahead of the GPU, never more
            if time.perf_counter() - ts0 >= cx.sustain_s and len(marks) > 2:
                break
        torch.cuda.synchronize()
        ts1 = time.perf_counter()
        r["sustained_ms"] = marks[0].elapsed_time(marks[-1])
        r["sustained_steps"] = n_steps
        half = len(marks) // 2                    # second half on its own: what the rate settles to
        r["sustained_2nd_half_ms_per_step"] = marks[half].elapsed_time(marks[-1]) / ((len(marks) - 1 - half) * chunk)
        if sclk is not None:
            r["sustained_clocks"] = sclk.stop(ts0 + 0.1, ts1)
        cx.barrier()
    if full:
        # ---- per-kernel timing for the roofline (same stream, CUDA events) ---------------------------------------
        per = {}
        be.marg_window_batch(db, capi.RUN_BOTH)
        for name, k in KERNELS.items():
            tot = 0.0
            reps = max(3, min(steps, 10))
            for _ in range(reps):
                if flush is not None:
                    flush.fill_(1)
                ev[0].record()
                be.marg_window_batch(db, k["which"])
                ev[1].record()
                torch.cuda.synchronize()
                tot += ev[0].elapsed_time(ev[1])
            per[name] = tot / reps
        r["per"] = per

    # ---- end to end through the host-pointer C ABI (pinned host buffers) ----------------------------------------
    if e2e:
        for f in batch.FIELDS:
            a = getattr(batch, f)
            if a is not None:
                t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                cx.keep.append(t)
                setattr(batch, f, t.numpy())
        hout_t = {"se3": (n, capi.SE3_REC), "pg": (n, capi.PG_REC), "rel": (n, capi.REL_REC), "vb": (n, capi.VB_REC),
                  "rp": (n, capi.RP_REC)}
        hout_t = {k: torch.zeros(s, dtype=torch.float64).pin_memory() for k, s in hout_t.items()}
        hout_t["rank"] = torch.zeros((n, 2), dtype=torch.int32).pin_memory()
        hout_t["status"] = torch.zeros((n,), dtype=torch.int32).pin_memory()
        hout = WindowOutputs(*[hout_t[k].numpy() for k in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
        fwd_f = ("lm_offset", "pose_fwd", "ex_pose", "prior_se3", "prior_rel", "prior_rp")
        d2h = sum(hout_t[k].numpy().nbytes for k in ((["se3", "pg"] if which & 1 else []) + (["rel", "vb", "rp"] if which & 2 else [])
                                                    + ["rank", "status"]))
        fams = (["se3", "pg"] if which & 1 else []) + (["rel", "vb", "rp"] if which & 2 else [])
        # (a) ABI 2 inputs: raw IMU samples instead of the 467-double pre-integration record (the library runs
        #     preintegrate_kernel first) and pts_i.z == 1 promised (src/System.cpp:346): 3 of the 6 landmark components
        #     cross PCIe (pts_j is never read by the information-only marginalization, pts_i.z is the constant 1);
        # (b) ABI 1 inputs (the full records), for comparison
        # (a') ABI 3 inputs: (a) plus pts_i.x / pts_i.y as the FP32 values the feature tracker produced (cv::Point2f in the
        #     reference): 4 + 4 + 8 bytes per landmark
        from is_vins_b200.backend import xy_as_f32
        xyf_t = torch.from_numpy(xy_as_f32(batch.lm_obs)).pin_memory()
        cx.keep.append(xyf_t)
        raw_f = ("pose_bwd", "sb_bwd", "prior_vb", "imu_raw", "imu_init")
        # (a'') ABI 4 inputs / outputs: (a') plus every record without its structural zeros -- the sqrt_info blocks are upper
        #     triangular, covRel symmetric: 252 instead of 319 doubles of records in, 191 instead of 289 doubles of results out
        from is_vins_b200.batch import pack_tri_inputs, packed_outputs, unpack_outputs
        tri = {k: (None if v is None else torch.from_numpy(v).pin_memory()) for k, v in pack_tri_inputs(batch).items()}
        cx.keep.append(tri)
        tri_np = {k: (None if v is None else v.numpy()) for k, v in tri.items()}
        pout_t = {k: torch.from_numpy(getattr(packed_outputs(n), k)).pin_memory() for k in ("se3", "pg", "rel", "vb", "rp", "rank", "status")}
        pout = WindowOutputs(*[pout_t[k].numpy() for k in ("se3", "pg", "rel", "vb", "rp", "rank", "status")])
        tri_saved = sum(getattr(batch, f).nbytes - tri_np[f].nbytes for f in ("prior_se3", "prior_rel", "prior_rp", "prior_vb")
                        if tri_np[f] is not None and ((f != "prior_vb" and which & 1) or (f == "prior_vb" and which & 2)))
        d2h_tri = sum(pout_t[k].numpy().nbytes for k in ((["se3", "pg"] if which & 1 else []) + (["rel", "vb", "rp"] if which & 2 else [])
                                                         + ["rank", "status"]))
        # the packed records pay from a few thousand windows on (the two expand / compact kernels add ~0.05 ms of latency to
        # a call: 0.74 vs 0.70 ms at 4096 windows of L ~ 150, 1.29 vs 1.43 ms at 9472): smaller batches take ABI 3
        use_tri = n >= 8192
        r["e2e_abi"] = 4 if use_tri else 3
        abi3_kw = dict(raw_imu=True, z_one=True, xy_f32=xyf_t.numpy())
        for tag, kw, lm_bytes, bwd_f in (("", dict(abi3_kw, tri_in=tri_np, tri_out=True) if use_tri else abi3_kw, 16, raw_f),
                                         ("_abi3", abi3_kw, 16, raw_f),
                                         ("_abi2", dict(raw_imu=True, z_one=True), 24, raw_f),
                                         ("_abi1", dict(), 32, ("pose_bwd", "sb_bwd", "prior_vb", "preint"))):
            if tag and not full:
                continue
            packed = use_tri and not tag
            ho = pout if packed else hout
            h2d = 0
            if which & 1:
                h2d += sum(getattr(batch, f).nbytes for f in fwd_f) + lm_bytes * batch.n_landmarks
            if which & 2:
                h2d += sum(getattr(batch, f).nbytes for f in bwd_f)
            if packed:
                h2d -= tri_saved
            for _ in range(max(1, min(warmup, 3))):
                be.marg_window_batch_host(batch, which, ho, **kw)
            cx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                be.marg_window_batch_host(batch, which, ho, **kw)   # synchronises internally after the D2H
            e1.record()
            cx.barrier()
            r["e2e_ms" + tag] = e0.elapsed_time(e1)
            r["h2d" + tag] = h2d
            r["d2h" + tag] = d2h_tri if packed else d2h
            assert int(np.count_nonzero(ho.status)) == 0
            if cx.rank == 0 and not tag:
                # the e2e path's results against the device path's: equal to rounding (z == 1 is folded into the arithmetic,
                # the pre-integration record was rebuilt on the GPU from the raw samples); the packed results are expanded
                # here, after the timed region
                from is_vins_b200.batch import outputs_rel_diff
                r["e2e_vs_device_path_max_rel"] = float(outputs_rel_diff(unpack_outputs(pout) if packed else hout, r["out"], which).max())
            if cx.rank == 0 and tag == "_abi1":
                r["e2e_abi1_bits_equal"] = bool(all(np.array_equal(getattr(hout, f), getattr(r["out"], f)) for f in fams))
    # ---- the copy ceiling the e2e number runs under: the same byte counts as plain pinned cudaMemcpyAsync, H2D and D2H
    #      overlapped on two streams, every rank at once (what the host / PCIe can feed; tools/h2d_peak.py is the long form)
    if e2e and full:
        hs = torch.empty(r["h2d"], dtype=torch.uint8).pin_memory()
        hd = torch.empty(r["d2h"], dtype=torch.uint8).pin_memory()
        ds = torch.empty(r["h2d"], dtype=torch.uint8, device=dev)
        dd = torch.ones(r["d2h"], dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()

        def copies():
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                ds.copy_(hs, non_blocking=True)
            with torch.cuda.stream(s2):
                hd.copy_(dd, non_blocking=True)
            cur.wait_stream(s1)
            cur.wait_stream(s2)
        copies()
        cx.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            copies()
        c1.record()
        cx.barrier()
        r["copy_ceiling_ms"] = c0.elapsed_time(c1) / 5
        del hs, hd, ds, dd
    del db
    return r


def single_window_latency(cx, batch):
    """One MARGIN_OLD event alone on the GPU, as a real-time estimator issues it: (a) inputs resident, one
    isv_marg_window_batch; (b) through the blocking single-window host entry points (isv_marg_forward +
    isv_marg_backward: pack, H2D, kernels, D2H); (c) through the fused isv_marg_event.  Median of 200 calls."""
    import torch
    from is_vins_b200 import DeviceBatch, capi
    be = cx.be
    lat = {}
    one = batch.slice(0, 1)
    d1 = DeviceBatch(one, cx.dev)
    for _ in range(20):
        be.marg_window_batch(d1, capi.RUN_BOTH)
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        be.marg_window_batch(d1, capi.RUN_BOTH)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    lat["device_resident_us"] = float(np.median(ts) * 1e6)
    o = one.lm_obs
    args1 = (one.pose_fwd[0, 0], one.pose_fwd[0, 1], one.ex_pose if one.ex_pose.ndim == 1 else one.ex_pose[0], o[5],
             np.ascontiguousarray(o[0:3].T), np.ascontiguousarray(np.vstack([o[3:5], np.ones((1, o.shape[1]))]).T),
             one.prior_se3[0], one.prior_rel[0], None if one.prior_rp is None else one.prior_rp[0])
    args2 = (one.pose_bwd[0, 0], one.sb_bwd[0, 0], one.pose_bwd[0, 1], one.sb_bwd[0, 1], one.prior_vb[0], one.preint[0])
    ts = []
    for it in range(220):
        t0 = time.perf_counter()
        be.marg_forward(*args1)
        be.marg_backward(*args2)
        if it >= 20:
            ts.append(time.perf_counter() - t0)
    lat["host_single_window_us"] = float(np.median(ts) * 1e6)
    ts = []
    for it in range(220):
        t0 = time.perf_counter()
        be.marg_event(args1, args2)
        if it >= 20:
            ts.append(time.perf_counter() - t0)
    lat["host_single_event_us"] = float(np.median(ts) * 1e6)
    # the same call timed INSIDE the library (isv_test_event_latency: what a C++ estimator pays per MARGIN_OLD event, without
    # the Python binding's marshalling), on its three routes: zero-copy fused kernel (default), fused kernel behind one H2D /
    # D2H, the five batch kernels behind one H2D / D2H
    ev = {}
    for mode, label in ((0, "zero_copy_fused_kernel"), (1, "staged_fused_kernel"), (2, "staged_batch_kernels")):
        be.set_tuning(capi.TUNE_EVENT_MODE, mode)
        be.event_latency_us(args1, args2, 30)
        us = be.event_latency_us(args1, args2, 300)
        ev[label] = {"median": float(np.median(us)), "p99": float(np.percentile(us, 99))}
    be.set_tuning(capi.TUNE_EVENT_MODE, 0)
    lat["event_c_abi_us"] = ev
    # the batch kernels on the resident window (what `device_resident_us` was before the fused kernel existed)
    be.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 0)
    for _ in range(20):
        be.marg_window_batch(d1, capi.RUN_BOTH)
    torch.cuda.synchronize()
    ts = []
    for _ in range(200):
        t0 = time.perf_counter()
        be.marg_window_batch(d1, capi.RUN_BOTH)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    be.set_tuning(capi.TUNE_FUSED_MAX_WINDOWS, 148)
    lat["device_resident_batch_kernels_us"] = float(np.median(ts) * 1e6)
    lat["note"] = ("device_resident_us / host_*_us are wall times through the Python binding (~15 us of interpreter and ctypes "
                   "overhead per call; host_single_event_us additionally ~30 us of NumPy marshalling); event_c_abi_us is the "
                   "latency at the C ABI")
    lat["landmarks"] = int(one.n_landmarks)
    return lat


def cpu_baseline_for(L, which, threads, n_literal, n_struct, seed=7):
    """oracle/isv_ref.c on the host cores: literal (the reference's dense algorithm) and structured (the CUDA path's
    O(L) elimination) on bounded samples of the same workload."""
    from oracle import ref_c
    ref_c.build()
    sb = make_batch(L, n_literal, seed)
    cpu_step(sb, threads, False, which)
    dt = cpu_step(sb, threads, False, which)
    sb2 = make_batch(L, n_struct, seed + 1)
    dts = cpu_step(sb2, threads, True, which)
    return {"value": n_literal / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_literal} windows of the same workload, literal dense algorithm of the reference (FullPivLU of "
                      f"the (L+6)^2 block), oracle/isv_ref.c, OpenMP over windows; the reference itself cannot be built "
                      f"here (no Eigen/Ceres/Sophus)",
            "structured_cpu_value": n_struct / dts,
            "structured_note": f"same C code with the O(L) landmark elimination the CUDA path uses ({n_struct} windows; not "
                               f"the reference's algorithm): reported so the algorithmic part of the speed-up is visible"}


def run_cuda(args, L):
    import torch
    import torch.distributed as dist
    from is_vins_b200 import MargBackend, capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    # a real (non-null) stream, shared with the library, so torch.cuda.Event sees the kernels
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on STDOUT: keep stdout = one JSON line
        # NCCL prints its version banner on STDOUT when the communicator is created: send fd 1 to stderr until the first
        # collective has run, so that stdout carries the JSON line and nothing else
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device=f"cuda:{local}")
            dist.all_reduce(warm)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    cx = Ctx()
    cx.world, cx.rank, cx.local, cx.dev = world, rank, local, f"cuda:{local}"
    cx.flush, cx.keep, cx.sustain_s = None, [], args.sustain

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    cx.barrier = barrier

    def maxr(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=cx.dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    cx.be = be = MargBackend(local)
    be.use_torch_stream()
    threads = os.cpu_count() or 1
    cpu_legs = world == 1 and not args.no_cpu

    # ---- headline: BASELINE configs[1] (or --features) -----------------------------------------------------------
    n = args.windows
    if args.quick:   # kernel development: device-timed value and the per-kernel times only
        cx.sustain_s = 0.0
        h = measure(cx, L, n, capi.RUN_BOTH, args.steps, args.warmup, 1000 + rank, full=True, e2e=False)
        if rank == 0:
            print(json.dumps({"quick": True, "value": world * n * args.steps / (h["ms_total"] * 1e-3), "ms_per_step": h["ms_total"] / args.steps,
                              "kernels_ms": h["per"], "L": L, "windows": n, "lib": os.environ.get("ISV_B200_LIB", "default")}), flush=True)
        be.close()
        return
    h = measure(cx, L, n, capi.RUN_BOTH, args.steps, args.warmup, 1000 + rank, full=True)
    names = list(KERNELS)
    tl = maxr([h["ms_total"], h["e2e_ms"], h["sustained_ms"] / h["sustained_steps"], h["sustained_2nd_half_ms_per_step"],
               h["e2e_ms_abi1"], h["copy_ceiling_ms"], h["e2e_ms_abi2"], h["e2e_ms_abi3"]] + [h["per"][k] for k in names])
    ms_total, e2e_ms, sus_ms, sus2_ms, e2e1_ms, ceil_ms, e2e2_ms, e2e3_ms = tl[0:8]
    per = dict(zip(names, tl[8:]))
    tot_lm = maxr([float(h["n_lm"])])[0]
    line = None
    if rank == 0:
        lat = single_window_latency(cx, h["batch"])
        value = world * n * args.steps / (ms_total * 1e-3)
        e2e = world * n * args.steps / (e2e_ms * 1e-3)
        peak, peak_src = read_peaks()
        fp, fp_src = measure_fp64_peak()
        fp_peak = fp.get("dfma")
        traffic = read_traffic()
        Lm = h["n_lm"] / n        # mean landmarks per window of rank 0's batch
        t_sum = sum(per.values())
        table = []
        for name in names:
            k = KERNELS[name]
            t = per[name] * 1e-3
            gbs = n * k["bytes"](Lm) / t / 1e9
            tf = n * k["flops"](Lm) / t / 1e12
            table.append({"kernel": name, "ms": per[name], "share": per[name] / t_sum,
                          "alg_bytes_per_window": k["bytes"](Lm), "hbm_gbs": gbs, "hbm_frac": gbs / peak,
                          "alg_flops_per_window": k["flops"](Lm), "fp64_tflops": tf,
                          "fp64_frac": (tf / fp_peak) if fp_peak else None,
                          "traffic": traffic_for(traffic, name, n, h["n_lm"]), "limiter": k["limiter"]})
        dom = [r for r in table if r["kernel"] == ROOFLINE_KERNEL][0]
        sus_value = world * n / (sus_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(L, n),
            "latency_us_per_window": ms_total / args.steps * 1e3 / n,
            "mean_landmarks_per_window": Lm,
            "sustained": {"value": sus_value, "unit": UNIT, "ms_per_step": sus_ms, "steps": h["sustained_steps"],
                          "seconds": h["sustained_ms"] * 1e-3, "second_half_value": world * n / (sus2_ms * 1e-3),
                          "vs_value": sus_value / value, "clocks": h.get("sustained_clocks"),
                          "note": "the K-step loop of `value` kept running back to back for >= 2 s; clocks / power sampled "
                                  "by nvidia-smi inside it (rank 0's GPU)"},
            "single_window_latency_us": lat,
            "clocks": h["clocks"],
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h["h2d"], "d2h_bytes_per_step": h["d2h"],
                    "ms_per_step": e2e_ms / args.steps, "gb_per_s": (h["h2d"] + h["d2h"]) / (e2e_ms / args.steps * 1e-3) / 1e9,
                    "input_abi": h.get("e2e_abi"),
                    "inputs": "ABI 4 = ABI 3 + every record without its structural zeros, in and out (upper-triangular sqrt_info "
                              "blocks as 21 / 45 / 3 numbers, the symmetric covRel as 21: 252 instead of 319 doubles of records per "
                              "window in, 191 instead of 289 doubles of results out; expanded / compacted on the device).  "
                              "ABI 3: raw IMU samples (12 + 7 K doubles; preintegrate_kernel runs inside the call) instead of the "
                              "467-double pre-integration record; pts_i.z == 1 promised (ISV_IN_PTS_I_Z_ONE); pts_i.x / pts_i.y as "
                              "the FP32 values the reference's feature tracker produces (cv::Point2f, widened exactly on the GPU): "
                              "4 + 4 + 8 bytes per landmark",
                    "max_rel_diff_vs_device_path": h.get("e2e_vs_device_path_max_rel"),
                    "copy_ceiling": {"ms_per_step": ceil_ms, "value": world * n / (ceil_ms * 1e-3), "unit": UNIT,
                                     "frac": (e2e_ms / args.steps and ceil_ms / (e2e_ms / args.steps)),
                                     "note": "the same H2D + D2H byte counts as plain pinned cudaMemcpyAsync on two streams, all "
                                             "ranks at once, max over ranks: what host memory / PCIe can feed on this box (the "
                                             "aggregate saturates near 115 GB/s at 2-4 GPUs and 186 GB/s at 8: "
                                             "profiles/r02n_h2d_peak_*gpu.json)"},
                    "abi3": {"value": world * n * args.steps / (e2e3_ms * 1e-3), "h2d_bytes_per_step": h["h2d_abi3"],
                             "d2h_bytes_per_step": h["d2h_abi3"], "ms_per_step": e2e3_ms / args.steps,
                             "inputs": "ABI 3: as above with the full (square) records in and out"},
                    "abi2": {"value": world * n * args.steps / (e2e2_ms * 1e-3), "h2d_bytes_per_step": h["h2d_abi2"],
                             "ms_per_step": e2e2_ms / args.steps,
                             "inputs": "ABI 2: as above with pts_i.x / pts_i.y as doubles (24 bytes per landmark)"},
                    "abi1": {"value": world * n * args.steps / (e2e1_ms * 1e-3), "h2d_bytes_per_step": h["h2d_abi1"],
                             "ms_per_step": e2e1_ms / args.steps, "results_bit_equal_device_path": h.get("e2e_abi1_bits_equal"),
                             "inputs": "ABI 1: the full records (467-double pre-integration, 4 doubles per landmark)"}},
            "gpu_launches": h["launches"],
            "kernels_ms": per,
            # The contract's object, always for the same kernel (the landmark kernel: the one that streams O(L) bytes per
            # window).  The window kernels are FP64 small-matrix chains, so the HBM fraction is low by construction; `fp64`
            # and `per_kernel` carry the FP64-pipe view and the same figures for every kernel of the step.
            "roofline": {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["hbm_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": dom["hbm_frac"], "traffic": dom["traffic"],
                         "traffic_source": {k: traffic.get(k) for k in ("file", "commit", "windows", "landmarks")} if traffic else None,
                         "peak_source": peak_src, "alg_bytes_per_window": dom["alg_bytes_per_window"],
                         "share_of_step": dom["share"], "limiter": dom["limiter"],
                         "fp64": {"achieved_tflops": dom["fp64_tflops"], "peak_tflops": fp_peak,
                                  "frac": dom["fp64_frac"], "alg_flops_per_window": dom["alg_flops_per_window"],
                                  "peak_source": fp_src, "dmma_peak_tflops": fp.get("dmma884"),
                                  "note": "FLOPs of the algorithm as implemented (DESIGN.md 4) vs the DFMA peak"},
                         "per_kernel": table},
        }
        if numa:
            line["rank0_cpu_affinity"] = f"{len(numa)} CPUs local to the GPU (NVML)"
        # parity of THIS batch's results (rank 0's shard) against the C oracle: test infrastructure used as the checker
        line["parity"] = parity_check(h["batch"], h["out"], 3, args.parity_windows, 4 if L <= 1000 else 0, threads)
        line["parity_max_rel_err"] = line["parity"]["max_rel_err"]
        if cpu_legs:
            line["cpu_baseline"] = cpu_baseline_for(L, 3, threads, threads, threads * 64)
    h = None
    cx.keep.clear()

    # ---- the other BASELINE configs, short -------------------------------------------------------------------------
    cfgs = []
    if not args.no_configs:
        n4 = max(1, 4096 // world)
        plan = [("configs[0]", 150, n, capi.RUN_BOTH, "MH_01-shaped window, L ~ 150 (the reference's own CPU-runnable case)"),
                ("configs[2]", 150, n, capi.RUN_BACKWARD, "second-newest-frame marginalization: MargBackward only"),
                ("configs[3](a)", 2000, n // 2, capi.RUN_BOTH, "~2000 features per window, no td (runtime-L engine; the "
                                                               "reference caps NUM_OF_F at 1000)"),
                ("configs[4]", 150, n4, capi.RUN_BOTH, f"4096 windows IN TOTAL sharded over {world} GPU(s) "
                                                       f"({n4} per GPU: a fraction of one wave -- latency regime)")]
        for ci, (name, Lc, nc, which, what) in enumerate(plan):
            r = measure(cx, Lc, nc, which, args.steps, args.warmup, 2000 + 100 * ci + rank, full=False)
            ms_c, e2e_c = maxr([r["ms_total"], r["e2e_ms"]])
            if rank == 0:
                c = {"config": name, "what": what, "workload": workload_config(Lc, nc, which)["workload"],
                     "l2": "flushed_between_steps" if r["flushed"] else "inputs_larger_than_l2",
                     "windows_per_gpu": nc, "mean_landmarks_per_window": r["n_lm"] / nc,
                     "value": world * nc * args.steps / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c / args.steps,
                     "latency_us_per_window": ms_c / args.steps * 1e3 / nc,
                     "e2e": {"value": world * nc * args.steps / (e2e_c * 1e-3), "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
                             "d2h_bytes_per_step": r["d2h"], "ms_per_step": e2e_c / args.steps, "input_abi": r.get("e2e_abi")},
                     "gpu_launches": r["launches"],
                     "parity": parity_check(r["batch"], r["out"], which, 32, 2 if Lc <= 1000 else 0, threads)}
                if cpu_legs:
                    nl = {150: 256, 2000: max(4, threads // 2)}.get(Lc, threads)
                    c["cpu_baseline"] = cpu_baseline_for(Lc, which, threads, nl, 1024)
                    c["vs_cpu_literal_e2e"] = c["e2e"]["value"] / c["cpu_baseline"]["value"]
                    c["vs_cpu_structured_e2e"] = c["e2e"]["value"] / c["cpu_baseline"]["structured_cpu_value"]
                cfgs.append(c)
            r = None
            cx.keep.clear()
        # configs[3] (b): WINDOW_SIZE = 20 with online td estimation through the generic MarginalizationInfo engine
        # (one GPU's worth of problems per rank; rank 0 reports its own rate x world: the engine shards like the windows)
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bench_marg_generic", os.path.join(ROOT, "tools", "bench_marg_generic.py"))
            bmg = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(bmg)
            g = bmg.w20_from_bench(be, 148, max(3, min(args.steps, 10)))
            ms_g = maxr([g["ms_per_step"]])[0]
            if rank == 0:
                c = {"config": "configs[3](b)", "what": "WINDOW_SIZE=20, ~2000 features, online td (ProjectionTdFactor), "
                     "VINS-Mono style MarginalizationInfo::marginalize of the oldest frame, 148 problems per GPU",
                     "workload": g["config"], "value": world * 148 / (ms_g * 1e-3), "unit": "problems/s", "ms_per_step": ms_g,
                     "latency_us_per_window": ms_g * 1e3 / 148, "kernels_ms": g["kernels_ms"]}
                if cpu_legs:
                    from oracle import isv_oracle as O
                    O.vins_mono_schur_eig(g["A_one"], g["b_one"], g["m"])
                    t0 = time.perf_counter()
                    for _ in range(5):
                        O.vins_mono_schur_eig(g["A_one"], g["b_one"], g["m"])
                    dt = (time.perf_counter() - t0) / 5
                    c["cpu_baseline"] = {"value": 1.0 / dt, "unit": "problems/s", "cores": threads, "kind": "port",
                                         "sample": "5 repetitions of one problem: the back half of marginalize() only (joint eigh "
                                                   "pseudo-inverse of A_mm, Schur complement, eigh of the reduced system) with "
                                                   "NumPy/LAPACK on the assembled normal equations -- assembling them "
                                                   "(ThreadsConstructA) is NOT included, so this flatters the CPU"}
                cfgs.append(c)
        except Exception as e:   # the headline must survive a failure of this auxiliary line
            if rank == 0:
                cfgs.append({"config": "configs[3](b)", "error": repr(e)})

    if rank == 0:
        line["configs"] = cfgs
        print(json.dumps(line), flush=True)
    be.close()
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(local):
    """One process per GPU: pin this rank to the CPUs NVML reports as local to its GPU BEFORE any pinned host buffer is
    allocated, so the e2e leg's H2D / D2H sources sit on the GPU's own NUMA node (first touch) instead of crossing the
    socket interconnect.  Returns the CPU list, or None when NVML / the topology is not available."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1 and 64 * i + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--windows", type=int, default=9472,
                    help="windows per GPU; default 9472 = 148 SMs x 16 resident warps x 4: one warp owns one window "
                         "and the window kernels keep 16 warps per SM resident, so the grid is a whole number of waves")
    ap.add_argument("--features", type=int, default=1000, help="mean L (1000 = configs[1]; 150 = configs[0]/[4])")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained block")
    ap.add_argument("--parity-windows", type=int, default=64, help="windows of the timed batch re-computed by the C oracle")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-configs", action="store_true", help="headline workload only")
    ap.add_argument("--quick", action="store_true", help="kernel development: `value` and the per-kernel times only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        run_reference(args, args.features)
    else:
        run_cuda(args, args.features)


if __name__ == "__main__":
    main()
