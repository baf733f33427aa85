"""-m gpu: the C++ host layer (is_vins_b200/host/isv_estimator_host.hpp: `Estimator::initFactorGraph /
MargForward / MargBackward / slideWindow` with the reference's member names) compiled with g++, linked
against libisv_b200.so and driven by tests/cpp/host_shim_test.cpp on a fixture of the oracle's chain."""
import os
import subprocess

import numpy as np
import pytest

from oracle import sim
from tests.helpers import oracle_outputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_shim_test.cpp")


SRC_MI = os.path.join(ROOT, "tests", "cpp", "marginalization_info_test.cpp")
SRC_CHAIN = os.path.join(ROOT, "tests", "cpp", "marginalization_chain_test.cpp")
O_SIZE = {"pose": 7, "speed_bias": 9, "ex_pose": 7, "feature": 1, "td": 1}


def build_host_test(out_path, src=SRC):
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", out_path, src, "-L" + os.path.join(ROOT, "is_vins_b200"),
           "-lisv_b200", "-Wl,-rpath," + os.path.join(ROOT, "is_vins_b200")]
    subprocess.run(cmd, check=True)


def _pre(p):
    out = [float(len(p.dt_buf))] + list(p.linearized_acc) + list(p.linearized_gyr) + list(p.linearized_ba) + list(p.linearized_bg)
    for dt, a, g in zip(p.dt_buf, p.acc_buf, p.gyr_buf):
        out += [dt] + list(a) + list(g)
    return out


def _rec48(t, R, s):
    return list(np.asarray(t).ravel()) + list(np.asarray(R).T.ravel()) + list(np.asarray(s).T.ravel())


def write_fixture(path, chain):
    V = chain.cfg.vo_size
    d = [float(len(chain.events)), float(V)]
    ii, io = chain.init_in, chain.init_out
    d += list(ii.poses.ravel()) + list(ii.sbs.ravel())
    for p in ii.pres:
        d += _pre(p)
    d += [float(io.rank)]
    for i in range(V - 1):
        d += _rec48(io.rel_dt[i], io.rel_dR[i], io.rel_sqrt_info[i])
    d += _rec48(io.se3_t, io.se3_R, io.se3_sqrt_info)
    d += list(io.vb) + list(io.vb_sqrt_info.T.ravel())
    ref = oracle_outputs(chain.events)
    for r, e in enumerate(chain.events):
        f, b = e.fwd_in, e.bwd_in
        L = len(f.inv_dep)
        d += [float(L)] + list(f.pose0) + list(f.pose1) + list(f.ex_pose) + list(f.inv_dep)
        d += list(np.asarray(f.pts_i).ravel()) + list(np.asarray(f.pts_j).ravel())
        d += _rec48(f.prior_t, f.prior_R, f.prior_sqrt_info) + _rec48(f.rel_dt, f.rel_dR, f.rel_sqrt_info)
        d += [1.0 if f.rp_valid else 0.0]
        d += list(b.pose_i) + list(b.sb_i) + list(b.pose_j) + list(b.sb_j) + list(b.vb_prior) + list(b.vb_sqrt_info.T.ravel())
        d += [e.pg_meta["ts"]] + list(e.pg_meta["Ri"].T.ravel()) + list(e.pg_meta["ti"])
        d += _pre(b.pre)
        d += [float(ref.rank[r, 0]), float(ref.rank[r, 1])]
        d += list(ref.se3[r]) + list(ref.pg[r]) + list(ref.rel[r]) + list(ref.vb[r]) + list(ref.rp[r])
        idx = sorted(e.state_after["rp"].keys())
        d += [float(len(idx))] + [float(i) for i in idx]
    np.asarray(d, dtype="<f8").tofile(path)


@pytest.mark.gpu
def test_cpp_host_estimator_against_oracle(tmp_path):
    exe = str(tmp_path / "host_shim_test")
    build_host_test(exe)
    chain = sim.make_chain(sim.seed_for(8, 0), L=[60, 33, 150, 1, 97, 32, 64, 40, 45, 50], rounds=10)
    fx = str(tmp_path / "fixture.bin")
    write_fixture(fx, chain)
    p = subprocess.run([exe, fx], capture_output=True, text=True, timeout=300)
    print(p.stdout, p.stderr)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "0 mismatches" in p.stdout


def write_marginalization_fixture(path, p):
    """The MARGIN_OLD marginalization problem of tests/test_marginalization_info_gpu.py and its 80-bit
    reduced system, in the order the C++ test adds the residual blocks."""
    from is_vins_b200.marginalization import LOCAL_SIZE
    from oracle import isv_oracle as O
    N, F = p.poses.shape[0], len(p.feat)
    hosted = [k for k in range(p.proj_idx.shape[1]) if int(p.proj_idx[0, k]) == 0]
    d = [float(N), float(F), float(len(hosted))] + list(p.poses.ravel()) + list(p.sbs.ravel()) + list(p.ex.ravel()) + list(p.feat)
    d += list(p.imu_pre[0].pack())
    # ordering by first appearance: pose0, sb0 | features (in factor order) | pose1, sb1, other poses
    order = [("pose", 0), ("speed_bias", 0)]
    feats, kept = [], [("pose", 1), ("speed_bias", 1)]
    ofac = []
    r, js = O.IMUFactor(p.imu_pre[0]).EvaluateCeres([p.poses[0], p.sbs[0], p.poses[1], p.sbs[1]])
    ofac.append((r, js, [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)]))
    s = p.cfg.proj_sqrt_info
    for k in hosted:
        i, j, e, f = [int(x) for x in p.proj_idx[:, k]]
        pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
        d += [float(j), float(f)] + list(pts_i) + list(pts_j)
        if ("feature", f) not in feats:
            feats.append(("feature", f))
        if ("pose", j) not in kept:
            kept.append(("pose", j))
        r, js = O.ProjectionFactor(pts_i, pts_j, s).EvaluateCeres([p.poses[i], p.poses[j], p.ex[e], p.feat[f:f + 1]])
        ofac.append(sim.cauchy_correct(r, js, 1.0) + ([("pose", i), ("pose", j), ("ex_pose", e), ("feature", f)],))
    se3, rel = p.se3[0], p.rel[0]
    d += _rec48(se3.t, se3.R, se3.sqrt_info) + _rec48(rel.delta_t, rel.delta_R, rel.sqrt_info)
    ofac.append(sim.cauchy_correct(*se3.EvaluateCeres([p.poses[0]]), 1.0) + ([("pose", 0)],))
    ofac.append(sim.cauchy_correct(*rel.EvaluateCeres([p.poses[0], p.poses[1]]), 1.0) + ([("pose", 0), ("pose", 1)],))
    idx, pos = {}, 0
    for key in order + feats + kept:
        idx[key] = pos
        pos += LOCAL_SIZE[key[0]]
    m = 15 + len(feats)
    facs = [(r, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k[0] != "ex_pose"])
            for r, js, keys in ofac]
    ref = O.vins_mono_marginalize(facs, pos, m, eps=1e-8)
    S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], m)
    e_ref = np.linalg.norm(ref["A_red"] - S_hp) / np.linalg.norm(S_hp)
    tol = max(1e-9, 2.0 * e_ref)
    d += [float(m), float(pos - m), tol] + list(S_hp.T.ravel()) + list(s_hp)
    np.asarray(d, dtype="<f8").tofile(path)


@pytest.mark.gpu
def test_cpp_marginalization_info_against_oracle(tmp_path):
    """The C++ MarginalizationInfo (north_star API) vs the 80-bit reduced system of the same problem."""
    exe = str(tmp_path / "marginalization_info_test")
    build_host_test(exe, SRC_MI)
    p = sim.make_problem(sim.seed_for(9, 1), n_features=120, max_track=9, host0=0.6)
    fx = str(tmp_path / "mi_fixture.bin")
    write_marginalization_fixture(fx, p)
    r = subprocess.run([exe, fx], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "0 mismatches" in r.stdout


def _chain_problem(p, use_td, seed):
    """Both rounds of the chain as (fixture doubles, oracle factor builders).  Round 2 = the window slid by one
    frame with every estimate moved a little (the next solve), frame 1 of `p` being the new oldest frame."""
    from oracle import isv_oracle as O
    rng = np.random.default_rng(seed)
    N, F = p.poses.shape[0], len(p.feat)
    tr = 0.033 / 480 if use_td else 0.0
    td1, td2 = (0.007, 0.0062) if use_td else (0.0, 0.0)
    s = p.cfg.proj_sqrt_info
    d = [float(N), float(F), 1.0 if use_td else 0.0, tr]
    d += list(p.poses.ravel()) + list(p.sbs.ravel()) + list(p.ex.ravel()) + list(p.feat) + [td1]
    poses2 = p.poses[1:].copy()
    poses2[:, 0:3] += rng.normal(0, 0.01, poses2[:, 0:3].shape)
    for q in poses2:
        dq = O.q_mul(O.quat_from_pose(q), np.concatenate([[1.0], rng.normal(0, 0.002, 3)]))
        dq /= np.linalg.norm(dq)
        q[3:6], q[6] = dq[1:4], dq[0]
    poses2[2, 3:7] *= -1.0                     # q and -q: the `w < 0` branch of MarginalizationFactor::Evaluate
    sbs2 = p.sbs[1:] + rng.normal(0, 0.003, p.sbs[1:].shape)
    feat2 = p.feat * (1.0 + rng.normal(0, 0.01, p.feat.shape))
    rounds = []
    for host, poses, sbs, feat, td, shift in ((0, p.poses, p.sbs, p.feat, td1, 0), (1, poses2, sbs2, feat2, td2, 1)):
        ofac = []
        keys = [("pose", 0), ("speed_bias", 0), ("pose", 1), ("speed_bias", 1)]
        r, js = O.IMUFactor(p.imu_pre[host]).EvaluateCeres([poses[0], sbs[0], poses[1], sbs[1]])
        ofac.append((r, js, keys))
        d += list(p.imu_pre[host].pack())
        hosted = [k for k in range(p.proj_idx.shape[1]) if int(p.proj_idx[0, k]) == host]
        d += [float(len(hosted))]
        for k in hosted:
            _, j, e, f = [int(x) for x in p.proj_idx[:, k]]
            j -= shift
            pts_i, pts_j = p.proj_obs[0:3, k], np.array([p.proj_obs[3, k], p.proj_obs[4, k], 1.0])
            d += [float(j), float(f)] + list(pts_i) + list(pts_j)
            keys = [("pose", 0), ("pose", j), ("ex_pose", e), ("feature", f)]
            if use_td:
                m = [rng.normal(0, 0.3, 2), rng.normal(0, 0.3, 2), rng.normal(0, 0.004), rng.normal(0, 0.004),
                     rng.uniform(-240, 240), rng.uniform(-240, 240)]
                d += list(m[0]) + list(m[1]) + m[2:]
                fac = O.ProjectionTdFactor(pts_i, pts_j, m[0], m[1], m[2], m[3], m[4], m[5], s, tr)
                r, js = fac.EvaluateCeres([poses[0], poses[j], p.ex[e], feat[f:f + 1], np.array([td])])
                keys = keys + [("td", 0)]
            else:
                r, js = O.ProjectionFactor(pts_i, pts_j, s).EvaluateCeres([poses[0], poses[j], p.ex[e], feat[f:f + 1]])
            ofac.append(sim.cauchy_correct(r, js, 1.0) + (keys,))
        if host == 0:
            se3, rel = p.se3[0], p.rel[0]
            d += _rec48(se3.t, se3.R, se3.sqrt_info) + _rec48(rel.delta_t, rel.delta_R, rel.sqrt_info)
            ofac.append(sim.cauchy_correct(*se3.EvaluateCeres([poses[0]]), 1.0) + ([("pose", 0)],))
            ofac.append(sim.cauchy_correct(*rel.EvaluateCeres([poses[0], poses[1]]), 1.0) + ([("pose", 0), ("pose", 1)],))
            d += list(poses2.ravel()) + list(sbs2.ravel()) + list(feat2) + [td2]
        rounds.append({"ofac": ofac, "pose": poses, "speed_bias": sbs, "feature": feat, "td": np.array([td]), "ex_pose": p.ex})
    return d, rounds


def _read_round(a, o, N, F):
    m, n, status, rank = [int(v) for v in a[o:o + 4]]
    o += 4
    out = {"m": m, "n": n, "status": status, "rank": rank}
    out["A_red"] = a[o:o + n * n].reshape(n, n).T; o += n * n
    out["b_red"] = a[o:o + n]; o += n
    out["J"] = a[o:o + n * n].reshape(n, n).T; o += n * n
    out["r"] = a[o:o + n]; o += n
    nk = int(a[o]); o += 1
    out["keep"] = [(int(a[o + 2 * k]), int(a[o + 2 * k + 1])) for k in range(nk)]; o += 2 * nk
    idx = {}
    for fam, cnt in (("pose", N), ("speed_bias", N), ("feature", F), ("td", 1)):
        for i in range(cnt):
            if a[o + i] >= 0:
                idx[(fam, i)] = int(a[o + i])
        o += cnt
    out["idx"] = idx
    return out, o


@pytest.mark.gpu
@pytest.mark.parametrize("use_td", [False, True])
def test_cpp_marginalization_chain_two_rounds(tmp_path, use_td):
    """C++ MarginalizationInfo, two chained MARGIN_OLD rounds (MarginalizationFactor kind; with use_td every
    visual factor is a ProjectionTdFactor and para_Td a kept block -- BASELINE configs[3]).  Each round's reduced
    system vs the 80-bit truth of the oracle's factors in the block order the C++ class chose; the round-2 prior
    block is the oracle's MarginalizationFactor built from the round-1 outputs."""
    from is_vins_b200.marginalization import LOCAL_SIZE
    from oracle import isv_oracle as O
    from tests.helpers import rel_err
    exe = str(tmp_path / "marginalization_chain_test")
    build_host_test(exe, SRC_CHAIN)
    p = sim.make_problem(sim.seed_for(9, 21 + int(use_td)), n_features=320, max_track=9, host0=0.4)
    d, rounds = _chain_problem(p, use_td, 77)
    fx, dump = str(tmp_path / "chain_fixture.bin"), str(tmp_path / "chain_dump.bin")
    np.asarray(d, dtype="<f8").tofile(fx)
    r = subprocess.run([exe, fx, dump], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stderr[-2000:]
    a = np.fromfile(dump, dtype="<f8")
    N, F = p.poses.shape[0], len(p.feat)
    r1, o = _read_round(a, 0, N, F)
    r2, o = _read_round(a, o, N, F)
    assert o == len(a)
    prev = None
    for rnd, got in zip(rounds, (r1, r2)):
        idx, m, n = got["idx"], got["m"], got["n"]
        assert got["status"] == 0 and idx[("pose", 0)] == 0 and idx[("speed_bias", 0)] == 6
        n_diag = sum(1 for k in idx if k[0] == "feature")
        assert n_diag >= 4 and m == 15 + n_diag
        assert sorted(v for k, v in idx.items() if k[0] == "feature") == list(range(15, m))
        if use_td:
            assert idx[("td", 0)] >= m                      # the time offset is kept
        ofac = list(rnd["ofac"])
        if prev is not None:     # the round-1 prior as a residual block of round 2 (VINS-Mono MarginalizationFactor)
            g1, rnd1 = prev
            inv = {v: k for k, v in g1["idx"].items()}
            keys1 = [inv[i] for _, i in g1["keep"]]
            assert [O_SIZE[k[0]] for k in keys1] == [sz for sz, _ in g1["keep"]]
            x0 = [np.atleast_1d(rnd1[k[0]][k[1]]) for k in keys1]
            keys2 = [((k[0], k[1] - 1) if k[0] in ("pose", "speed_bias") else k) for k in keys1]
            mf = O.MarginalizationFactor(g1["J"], g1["r"], [(sz, i - g1["m"]) for sz, i in g1["keep"]], x0)
            r_pr, j_pr = mf.EvaluateCeres([np.atleast_1d(rnd[k[0]][k[1]]) for k in keys2])
            ofac.append((r_pr, j_pr, keys2))
        pos = m + n
        facs = [(r_, [(idx[k], np.asarray(j)[:, :LOCAL_SIZE[k[0]]]) for k, j in zip(keys, js) if k[0] != "ex_pose"])
                for r_, js, keys in ofac]
        ref = O.vins_mono_marginalize(facs, pos, m, eps=1e-8)
        assert ref["min_eig_Amm"] > 1e-8
        S_hp, s_hp = O.schur_complement_longdouble(ref["A"], ref["b"], m)
        e_ref, e_gpu = rel_err(ref["A_red"], S_hp), rel_err(got["A_red"], S_hp)
        print(f"round {1 if prev is None else 2} A_red vs 80-bit truth: literal FP64 oracle {e_ref:.2e}, C++/CUDA {e_gpu:.2e}")
        assert e_gpu <= max(1e-9, 2.0 * e_ref)
        assert rel_err(got["b_red"], s_hp) <= max(1e-9, 2.0 * rel_err(ref["b_red"], s_hp))
        assert rel_err(got["J"].T @ got["J"], got["A_red"]) <= 1e-7
        cols = 0
        for sz, i in got["keep"]:
            assert i == m + cols
            cols += 6 if sz == 7 else sz
        assert cols == n
        prev = (got, rnd)
    # the kept set of round 2 carries every block the round-1 prior touched, minus the dropped frame
    kept1 = {((k[0], k[1] - 1) if k[0] in ("pose", "speed_bias") else k) for k, v in r1["idx"].items() if v >= r1["m"]}
    kept2 = {k for k, v in r2["idx"].items() if v >= r2["m"]}
    assert kept2 >= kept1 - {("pose", 0), ("speed_bias", 0)}


def test_cpp_host_layer_compiles(tmp_path):
    """CPU: the host layer, the C++ MarginalizationInfo and their test drivers compile and link against the
    C ABI (no compute)."""
    build_host_test(str(tmp_path / "host_shim_test"))
    build_host_test(str(tmp_path / "marginalization_info_test"), SRC_MI)
    build_host_test(str(tmp_path / "marginalization_chain_test"), SRC_CHAIN)
