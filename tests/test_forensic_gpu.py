"""-m gpu: forensic mode (VERDICT r1 "missing" 3 + 4).  The GPU's own dense route -- Lamda assembled by the ceres-Evaluate
kernels + ne_build_kernel, the whole (L + 6)^2 block inverted with full pivoting (src/estimator.cpp:1286-1288) -- must
reproduce the structured kernels' Lamda_prior, both must match the NumPy oracle's literal algorithm, and the KLD
diagnostics of :1333-1345 / :1522-1534 must match the oracle's."""
import numpy as np
import pytest

from is_vins_b200 import DeviceBatch, capi, forensic_batch, literal_forward, pack_events
from oracle import sim
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def test_literal_route_equals_structured_route_and_kld_matches(backend):
    ch = sim.make_chain(sim.seed_for(3, 1), L=[150, 33, 64, 1, 97], rounds=5, max_gap=2)
    events = ch.events
    batch = pack_events(events)
    db = DeviceBatch(batch, "cuda:0")
    fz = forensic_batch(backend, db)
    out = db.outputs()
    assert not out.status.any()
    for w, ev in enumerate(events):
        fo, bo = ev.fwd_out, ev.bwd_out
        lit = literal_forward(backend, batch, w)
        assert lit["status"] == 0 and lit["rank_mm"] == len(ev.fwd_in.inv_dep) + 6
        # (1) the dense Lamda assembled on the GPU == the oracle's block loop (:1168-1242)
        assert rel_err(lit["Lamda"], fo.Lamda) <= 1e-12, w
        # (2) GPU literal == GPU structured == oracle literal, 1e-9
        e_ls = rel_err(lit["Lamda_prior"], fz["lamda_prior_fwd"][w])
        e_so = rel_err(fz["lamda_prior_fwd"][w], fo.Lamda_prior)
        e_lo = rel_err(lit["Lamda_prior"], fo.Lamda_prior)
        print(f"window {w} L={len(ev.fwd_in.inv_dep)}: literal vs structured {e_ls:.2e}, structured vs oracle {e_so:.2e}, "
              f"literal vs oracle {e_lo:.2e}; kld fwd {fz['kld_fwd'][w]:.2e} (oracle {fo.kld:.2e}), "
              f"kld bwd {fz['kld_bwd'][w]:.6e} (oracle {bo.kld:.6e})")
        assert max(e_ls, e_so, e_lo) <= 1e-9, w
        # (3) backward intermediates: G^T G = Lamda_prior (:1419), its spectrum, the discarded abs / yaw informations
        assert rel_err(fz["lamda_prior_bwd"][w], bo.Lamda_prior) <= 1e-9
        lam_o = np.sort(bo.eigvals)
        assert np.allclose(fz["eig_bwd"][w][-15:], lam_o[-15:], rtol=1e-9)
        assert np.all(np.abs(fz["eig_bwd"][w][:6]) <= 1e-9 * lam_o[-1])
        assert rel_err(fz["info_abs"][w], bo.abs_info) <= 1e-8 and abs(fz["info_yaw"][w] - bo.yaw_info.ravel()[0]) <= 1e-8 * abs(bo.yaw_info.ravel()[0])
        # (4) KLD: forward is identically zero in exact arithmetic (phi == Lamda_prior when Jr is square and regular);
        #     backward is the information lost by the block-diagonal recovery: non-negative, equal to the oracle's
        assert abs(fz["kld_fwd"][w]) <= 1e-6 and abs(fo.kld) <= 1e-6
        assert fz["kld_bwd"][w] >= -1e-9
        assert abs(fz["kld_bwd"][w] - bo.kld) <= 1e-7 * max(1.0, abs(bo.kld)), (fz["kld_bwd"][w], bo.kld)


def test_literal_schur_zero_fills_dependent_unknowns(backend):
    """rank-deficient marginalized block: accepted pivots = rank, the rest zero-filled like FullPivLU::solve."""
    import ctypes as C

    import torch
    rng = np.random.default_rng(4)
    B = rng.normal(size=(8, 5))
    M = B @ B.T                                   # 8 x 8, rank 5
    X = rng.normal(size=(3, 8))
    A = np.block([[np.eye(3) * 50 + X @ X.T, X @ M], [M @ X.T, M]])
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).to("cuda:0")
    prior = torch.zeros((9,), dtype=torch.float64, device="cuda:0")
    inv = torch.zeros((64,), dtype=torch.float64, device="cuda:0")
    rank = torch.zeros((1,), dtype=torch.int32, device="cuda:0")
    capi.check(backend.lib.isv_literal_schur(backend.h, 1, 11, 3, C.c_void_p(dA.data_ptr()), C.c_void_p(prior.data_ptr()),
                                             C.c_void_p(inv.data_ptr()), C.c_void_p(rank.data_ptr())), "isv_literal_schur")
    backend.synchronize()
    assert int(rank.item()) == 5
    Mi = inv.cpu().numpy().reshape(8, 8).T
    assert np.linalg.norm(M @ Mi @ M - M) <= 1e-9 * np.linalg.norm(M)          # a generalised inverse on range(M)
    from oracle import isv_oracle as O
    ref = O.full_piv_lu_solve_identity(M)
    P = prior.cpu().numpy().reshape(3, 3).T
    assert rel_err(P, A[:3, :3] - A[:3, 3:] @ ref @ A[3:, :3]) <= 1e-9
