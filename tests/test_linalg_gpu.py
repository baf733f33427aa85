"""-m gpu: the warp-level PSD eigensolver (pivoted Cholesky + one-sided Jacobi) against numpy.eigh."""
import ctypes as C

import numpy as np
import pytest

from is_vins_b200 import capi

pytestmark = pytest.mark.gpu


def _run(backend, A):
    nb, n, _ = A.shape
    Af = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))   # column-major per matrix
    G = np.zeros((nb, n, n))
    lam = np.zeros((nb, n))
    info = np.zeros((nb, 2), np.int32)
    p = lambda a: a.ctypes.data_as(capi.c_double_p)
    capi.check(backend.lib.isv_test_psd_eig(backend.h, nb, n, p(Af), p(G), p(lam), info.ctypes.data_as(capi.c_int32_p)))
    return G, lam, info


@pytest.mark.parametrize("n,rank", [(6, 6), (6, 3), (21, 15), (21, 21), (57, 42)])
def test_psd_eig_matches_eigh(backend, n, rank):
    rng = np.random.default_rng(n * 100 + rank)
    nb = 64
    A = np.zeros((nb, n, n))
    for b in range(nb):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        w = np.zeros(n)
        w[:rank] = 10.0 ** rng.uniform(2, 9, rank)       # graded spectrum like the marginal information
        A[b] = (Q * w) @ Q.T
        A[b] = 0.5 * (A[b] + A[b].T)
    G, lam, info = _run(backend, A)
    for b in range(nb):
        w_ref = np.linalg.eigvalsh(A[b])[::-1]
        r = int(info[b, 0])
        assert info[b, 1] < 30
        kept = np.sort(lam[b][lam[b] > 0.1])[::-1]
        assert kept.size == rank
        # eigvalsh is only ABSOLUTELY accurate (eps*||A||); the Jacobi path is relatively accurate
        assert np.allclose(kept, w_ref[:rank], rtol=1e-9, atol=1e-14 * w_ref[0] * n)
        Gb = G[b][:r]
        # rows orthogonal, A reproduced
        GG = Gb @ Gb.T
        off = GG - np.diag(np.diag(GG))
        assert np.abs(off).max() <= 1e-13 * np.abs(np.diag(GG)).max()
        assert np.linalg.norm(Gb.T @ Gb - A[b]) <= 1e-13 * np.linalg.norm(A[b])
        # truncated pseudo-inverse
        keep = lam[b][:r] > 0.1
        Sig = (Gb[keep].T / lam[b][:r][keep] ** 2) @ Gb[keep]
        ww, V = np.linalg.eigh(A[b])
        k = ww > 0.1
        Sig_ref = (V[:, k] / ww[k]) @ V[:, k].T
        assert np.linalg.norm(Sig - Sig_ref) <= 1e-8 * np.linalg.norm(Sig_ref)


@pytest.mark.parametrize("n,nb", [(1, 4), (2, 8), (6, 16), (57, 32), (58, 8), (129, 8), (307, 4), (600, 2)])
def test_sym_eig_matches_eigh(backend, n, nb):
    """The CTA-level symmetric eigensolver of the generic engine's reduced system (Householder tridiagonalization +
    implicit QL = the algorithm class of Eigen's SelfAdjointEigenSolver) against LAPACK: graded spectra like a
    marginal information matrix, incl. rank-deficient ones and a block-diagonal one (interior deflation)."""
    rng = np.random.default_rng(1000 + n)
    A = np.zeros((nb, n, n))
    for b in range(nb):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        w = 10.0 ** rng.uniform(-2, 9, n)
        if b % 3 == 1:
            w[: n // 3] = 0.0                              # rank-deficient (unobservable directions)
        A[b] = (Q * w) @ Q.T
        if b % 4 == 2 and n >= 4:                          # decoupled blocks: zero sub-diagonals inside the tridiagonal
            A[b][: n // 2, n // 2:] = 0.0
            A[b][n // 2:, : n // 2] = 0.0
        A[b] = 0.5 * (A[b] + A[b].T)
    Af = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
    lam, V = np.zeros((nb, n)), np.zeros((nb, n, n))
    info = np.zeros((nb, 2), np.int32)
    p = lambda a: a.ctypes.data_as(capi.c_double_p)
    capi.check(backend.lib.isv_test_sym_eig(backend.h, nb, n, p(Af), p(lam), p(V), info.ctypes.data_as(capi.c_int32_p)))
    assert not info[:, 0].any()
    assert np.all(info[:, 1] <= 3 * n * n)                 # rotations logged (typically ~1.1 n^2)
    print(f"n={n}: {info[:, 1].mean() / max(n * n, 1):.2f} n^2 QL rotations")
    for b in range(nb):
        w_ref = np.linalg.eigvalsh(A[b])
        nrm = np.abs(w_ref).max()
        Vb = V[b].T                                        # columns = eigenvectors
        assert np.all(np.diff(lam[b]) >= 0)
        assert np.abs(lam[b] - w_ref).max() <= 1e-13 * n * nrm
        assert np.abs(Vb.T @ Vb - np.eye(n)).max() <= 1e-13 * n
        assert np.linalg.norm((Vb * lam[b]) @ Vb.T - A[b]) <= 1e-13 * n * np.linalg.norm(A[b])


def test_sym_eig_flags_non_finite_input_and_terminates(backend):
    """A NaN in the reduced system must not hang the serial QL (bounded sweeps, bounded rotation log) and must be
    reported: info[.,0] != 0 -> ISV_W_EIG_NOCONV in the marginalization status."""
    n, nb = 24, 3
    rng = np.random.default_rng(5)
    A = np.zeros((nb, n, n))
    for b in range(nb):
        M = rng.normal(size=(n, n))
        A[b] = M @ M.T
    A[1][3, 7] = A[1][7, 3] = np.nan
    Af = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
    lam, V = np.zeros((nb, n)), np.zeros((nb, n, n))
    info = np.zeros((nb, 2), np.int32)
    p = lambda a: a.ctypes.data_as(capi.c_double_p)
    capi.check(backend.lib.isv_test_sym_eig(backend.h, nb, n, p(Af), p(lam), p(V), info.ctypes.data_as(capi.c_int32_p)))
    assert info[1, 0] != 0 and info[0, 0] == 0 and info[2, 0] == 0
    assert info[1, 1] <= 3 * n * n + 64
    for b in (0, 2):                                       # the neighbours of the bad problem are untouched
        assert np.abs(lam[b] - np.linalg.eigvalsh(A[b])).max() <= 1e-12 * np.abs(lam[b]).max()


def test_fast_rsqrt_and_rcp_are_faithful(backend):
    """The lean reciprocal square root / reciprocal of the hot chains (MUFU seed + two Newton steps instead of CUDA's rsqrt()
    and division): within 2 ulp of the correctly rounded value over 300 decades of normal positive arguments, and exact
    powers of two come out exact."""
    import ctypes as C
    from is_vins_b200 import capi
    rng = np.random.default_rng(5)
    x = np.concatenate([10.0 ** rng.uniform(-150, 150, 200000), rng.uniform(0.5, 2.0, 100000), 2.0 ** np.arange(-200, 201, 2.0)])
    rs, rc = np.zeros_like(x), np.zeros_like(x)
    dp = lambda a: a.ctypes.data_as(capi.c_double_p)
    capi.check(backend.lib.isv_test_fast_special(backend.h, len(x), dp(x), dp(rs), dp(rc)), "isv_test_fast_special")
    ref_rs = (1.0 / np.sqrt(x.astype(np.longdouble))).astype(np.float64)
    ref_rc = (1.0 / x.astype(np.longdouble)).astype(np.float64)
    ulp = lambda a, b: np.abs(a - b) / np.spacing(np.abs(b))
    print(f"\nfast_rsqrt worst {ulp(rs, ref_rs).max():.2f} ulp, fast_rcp worst {ulp(rc, ref_rc).max():.2f} ulp")
    assert ulp(rs, ref_rs).max() <= 2.0 and ulp(rc, ref_rc).max() <= 2.0
    p2 = x[-201:]
    assert np.array_equal(rs[-201:], 1.0 / np.sqrt(p2)) and np.array_equal(rc[-201:], 1.0 / p2)
