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


def build_host_test(out_path):
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", out_path, SRC, "-L" + os.path.join(ROOT, "is_vins_b200"),
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


def test_cpp_host_layer_compiles(tmp_path):
    """CPU: the host layer and its test driver compile and link against the C ABI (no compute)."""
    build_host_test(str(tmp_path / "host_shim_test"))
