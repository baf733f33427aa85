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
