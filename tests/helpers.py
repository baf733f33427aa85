"""Shared comparison helpers: CUDA outputs (WindowOutputs) vs the oracle's MargEvent outputs."""
import numpy as np


def rel_err(x, ref):
    x, ref = np.asarray(x, float), np.asarray(ref, float)
    d = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / d) if d > 0 else float(np.linalg.norm(x))


def compare_event(out, w, ev, which=3):
    """Relative Frobenius errors of every recovered factor of window w against the oracle."""
    errs = {}
    if which & 1:
        fo = ev.fwd_out
        errs["se3_t"] = rel_err(out.se3_t(w), fo.se3_t)
        errs["se3_R"] = rel_err(out.se3_R(w), fo.se3_R)
        errs["se3_sqrt_info"] = rel_err(out.se3_sqrt_info(w), fo.se3_sqrt_info)
        errs["pg_dt"] = rel_err(out.pg_dt(w), fo.pg_dt)
        errs["pg_dR"] = rel_err(out.pg_dR(w), fo.pg_dR)
        errs["pg_sqrt_info"] = rel_err(out.pg_sqrt_info(w), fo.pg_sqrt_info)
        errs["pg_covRel"] = rel_err(out.pg_covRel(w), fo.pg_covRel)
        errs["pg_distance"] = rel_err(out.pg_distance(w), fo.pg_distance)
        if fo.pg_covAbs is not None:
            errs["pg_covAbs"] = rel_err(out.pg_covAbs(w), fo.pg_covAbs)
    if which & 2:
        bo = ev.bwd_out
        errs["rel_dt"] = rel_err(out.rel_dt(w), bo.rel_dt)
        errs["rel_dR"] = rel_err(out.rel_dR(w), bo.rel_dR)
        errs["rel_sqrt_info"] = rel_err(out.rel_sqrt_info(w), bo.rel_sqrt_info)
        errs["vb_VB"] = rel_err(out.vb_VB(w), bo.vb)
        errs["vb_sqrt_info"] = rel_err(out.vb_sqrt_info(w), bo.vb_sqrt_info)
        errs["rp_R"] = rel_err(out.rp_R(w), bo.rp_R)
        errs["rp_sqrt_info"] = rel_err(out.rp_sqrt_info(w), bo.rp_sqrt_info)
    return errs


def expected_ranks(ev):
    fo, bo = ev.fwd_out, ev.bwd_out
    return (fo.qr_rank if not fo.used_eig_path else fo.eig_rank), bo.rank


def oracle_outputs(events):
    """Oracle outputs of a list of MargEvents in the C ABI's record layout (WindowOutputs)."""
    from is_vins_b200 import capi
    from is_vins_b200.batch import WindowOutputs, se3_record, vb_record
    n = len(events)
    o = WindowOutputs(np.zeros((n, capi.SE3_REC)), np.zeros((n, capi.PG_REC)), np.zeros((n, capi.REL_REC)),
                      np.zeros((n, capi.VB_REC)), np.zeros((n, capi.RP_REC)), np.zeros((n, 2), np.int32),
                      np.zeros((n,), np.int32))
    for w, ev in enumerate(events):
        fo, bo = ev.fwd_out, ev.bwd_out
        o.se3[w] = se3_record(fo.se3_t, fo.se3_R, fo.se3_sqrt_info)
        o.pg[w, 0:48] = se3_record(fo.pg_dt, fo.pg_dR, fo.pg_sqrt_info)
        o.pg[w, 48:84] = fo.pg_covRel.flatten(order="F")
        o.pg[w, 84] = fo.pg_distance
        if fo.pg_covAbs is not None:
            o.pg[w, 85:89] = fo.pg_covAbs.flatten(order="F")
        o.rel[w] = se3_record(bo.rel_dt, bo.rel_dR, bo.rel_sqrt_info)
        o.vb[w] = vb_record(bo.vb, bo.vb_sqrt_info)
        o.rp[w, 0:9] = bo.rp_R.flatten(order="F")
        o.rp[w, 9:13] = bo.rp_sqrt_info.flatten(order="F")
        o.rank[w] = expected_ranks(ev)
    return o


RECORDS = ("se3", "pg", "rel", "vb", "rp")


def compare_outputs(out, ref, which=3):
    """Worst relative Frobenius error per record family between two WindowOutputs (sub-blocks of a
    record are compared separately so a large-magnitude block cannot mask a small one)."""
    blocks = {"se3": [(0, 3), (3, 12), (12, 48)], "pg": [(0, 3), (3, 12), (12, 48), (48, 84), (84, 85), (85, 89)],
              "rel": [(0, 3), (3, 12), (12, 48)], "vb": [(0, 9), (9, 90)], "rp": [(0, 9), (9, 13)]}
    fam = [f for f in RECORDS if (which & 1 and f in ("se3", "pg")) or (which & 2 and f in ("rel", "vb", "rp"))]
    worst = {}
    for f in fam:
        a, b = getattr(out, f), getattr(ref, f)
        e = 0.0
        for w in range(a.shape[0]):
            for lo, hi in blocks[f]:
                e = max(e, rel_err(a[w, lo:hi], b[w, lo:hi]))
        worst[f] = e
    return worst


def save_batch(path, batch, ref_out, extra=None):
    d = {f: getattr(batch, f) for f in batch.FIELDS if getattr(batch, f) is not None}
    d.update({"ref_" + f: getattr(ref_out, f) for f in RECORDS + ("rank", "status")})
    d.update(extra or {})
    np.savez_compressed(path, **d)


def load_batch(path):
    from is_vins_b200.batch import WindowBatch, WindowOutputs
    z = np.load(path)
    g = lambda k: z[k] if k in z.files else None
    n = int(z["pose_fwd"].shape[0])
    b = WindowBatch(n, z["lm_offset"], z["lm_obs"], z["pose_fwd"], z["ex_pose"], z["prior_se3"], z["prior_rel"],
                    g("prior_rp"), z["pose_bwd"], z["sb_bwd"], z["prior_vb"], z["preint"], g("imu_raw"), g("imu_init"))
    ref = WindowOutputs(z["ref_se3"], z["ref_pg"], z["ref_rel"], z["ref_vb"], z["ref_rp"], z["ref_rank"],
                        z["ref_status"])
    return b, ref, z
