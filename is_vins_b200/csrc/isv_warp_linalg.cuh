// Warp-cooperative dense FP64 linear algebra on small matrices held in shared memory.
// One warp owns one window; every routine is called by all 32 lanes of that warp with identical
// arguments and ends with a __syncwarp(), so results are visible to all lanes on return.
// Matrices are COLUMN-MAJOR with explicit leading dimension (element (i,j) at A[i + j*ld]).
//
// These replace the un-vendored Eigen decompositions the reference calls on this path
// (SURVEY.md 8c): FullPivLU::solve / MatrixXd::inverse() -> w_inverse (partial-pivot Gauss-Jordan),
// LLT -> w_chol_lower, SelfAdjointEigenSolver -> w_jacobi_eig (parallel-order two-sided Jacobi),
// BDCSVD pseudo-inverse -> eigen-decomposition of J J^T (see fwd tail).  Any backward-stable
// substitute is equivalent within the conditioning-scaled tolerance; discrete outcomes (ranks,
// the `> ALPHA` cut) are compared exactly by the tests.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace isv {

constexpr unsigned kFullMask = 0xffffffffu;

// The window kernels are long straight-line programs executed once per warp: with every routine
// inlined at every call site they are 170-220 KB of SASS and instruction fetch becomes a limiter
// (ncu: stall_no_instruction).  -DISV_HEAVY=__noinline__ compiles the heavy routines out of line (one
// copy per template instantiation); measured on B200 it is 3-5 % SLOWER than full inlining (the
// run-time loop bounds cost more than the fetches save), so inlining stays the default.
#ifndef ISV_HEAVY
#define ISV_HEAVY __forceinline__
#endif

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

__device__ __forceinline__ void w_fill(double* A, int count, double v, int lane) {
  for (int i = lane; i < count; i += 32) A[i] = v;
  __syncwarp();
}
__device__ __forceinline__ void w_copy(double* dst, const double* src, int count, int lane) {
  for (int i = lane; i < count; i += 32) dst[i] = src[i];
  __syncwarp();
}
// dst (m x n, ld ldd) = src (m x n, ld lds)
__device__ __forceinline__ void w_copy2d(double* dst, int ldd, const double* src, int lds, int m, int n, int lane) {
  for (int idx = lane; idx < m * n; idx += 32) {
    int i = idx % m, j = idx / m;
    dst[i + j * ldd] = src[i + j * lds];
  }
  __syncwarp();
}

// C (m x n) {=, +=, -=} op(A) (m x k) * op(B) (k x n).   mode: 0 set, 1 add, -1 subtract.
// TA: A is stored k x m (use A^T);  TB: B is stored n x k (use B^T).
template <bool TA, bool TB>
__device__ ISV_HEAVY void w_gemm(int m, int n, int k, const double* A, int lda, const double* B, int ldb,
                                       double* C, int ldc, int mode, int lane) {
  for (int idx = lane; idx < m * n; idx += 32) {
    int i = idx % m, j = idx / m;
    double acc = 0.0;
    for (int l = 0; l < k; ++l) {
      double a = TA ? A[l + i * lda] : A[i + l * lda];
      double b = TB ? B[j + l * ldb] : B[l + j * ldb];
      acc = fma(a, b, acc);
    }
    double* c = &C[i + j * ldc];
    if (mode == 0) *c = acc;
    else if (mode > 0) *c += acc;
    else *c -= acc;
  }
  __syncwarp();
}

// Compile-time-shaped variant: index arithmetic folds to shifts / multiplies, the k loop is fully unrolled
// with two accumulators and all its shared-memory loads issued up front (the window kernels are bound by
// dependent-issue latency, so exposed LDS round trips matter more than instruction count).
template <bool TA, bool TB, int M, int N, int K>
__device__ __forceinline__ void w_gemm_t(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                         double* __restrict__ C, int ldc, int mode, int lane) {
#pragma unroll
  for (int idx0 = 0; idx0 < M * N; idx0 += 32) {
    const int idx = idx0 + lane;
    if (idx < M * N) {
      const int i = idx % M, j = idx / M;
      double av[K], bv[K];
#pragma unroll
      for (int l = 0; l < K; ++l) {
        av[l] = TA ? A[l + i * lda] : A[i + l * lda];
        bv[l] = TB ? B[j + l * ldb] : B[l + j * ldb];
      }
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int l = 0; l < K; ++l) {
        if (l & 1) a1 = fma(av[l], bv[l], a1);
        else a0 = fma(av[l], bv[l], a0);
      }
      const double acc = a0 + a1;
      double* c = &C[i + j * ldc];
      if (mode == 0) *c = acc;
      else if (mode > 0) *c += acc;
      else *c -= acc;
    }
  }
  __syncwarp();
}

// Mirror the lower triangle onto the upper one (n x n).
__device__ __forceinline__ void w_symmetrize_from_lower(double* A, int ld, int n, int lane) {
  for (int idx = lane; idx < n * n; idx += 32) {
    int i = idx % n, j = idx / n;
    if (i < j) A[i + j * ld] = A[j + i * ld];
  }
  __syncwarp();
}

// In-place inverse by Gauss-Jordan with partial (row) pivoting on the augmented matrix [A | I].
// A: n x n (ld lda) is overwritten by A^-1.  work: n * 2n doubles.  Returns 1 if a zero pivot met.
__device__ ISV_HEAVY int w_inverse(double* A, int lda, int n, double* work, int lane) {
  const int n2 = 2 * n;
  // W is n x 2n, column-major, ld = n
  for (int idx = lane; idx < n * n2; idx += 32) {
    int i = idx % n, j = idx / n;
    work[idx] = (j < n) ? A[i + j * lda] : ((j - n) == i ? 1.0 : 0.0);
  }
  __syncwarp();
  int singular = 0;
  for (int k = 0; k < n; ++k) {
    // pivot search over rows k..n-1 of column k (n <= 32)
    double v = (lane >= k && lane < n) ? fabs(work[lane + k * n]) : -1.0;
    int piv = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(kFullMask, v, o);
      int op = __shfl_xor_sync(kFullMask, piv, o);
      if (ov > v || (ov == v && op < piv)) { v = ov; piv = op; }
    }
    if (!(v > 0.0)) singular = 1;
    if (piv != k) {
      for (int j = lane; j < n2; j += 32) {
        double t = work[k + j * n];
        work[k + j * n] = work[piv + j * n];
        work[piv + j * n] = t;
      }
    }
    __syncwarp();
    double p = work[k + k * n];
    __syncwarp();
    // multipliers f_i = W[i][k] / p  (as LU does), stored in column k in place; row k scaled last
    if (lane < n && lane != k) work[lane + k * n] = work[lane + k * n] / p;
    __syncwarp();
    const int ncols = n2 - (k + 1);
    for (int idx = lane; idx < n * ncols; idx += 32) {
      int i = idx % n, j = k + 1 + idx / n;
      if (i != k) work[i + j * n] = fma(-work[i + k * n], work[k + j * n], work[i + j * n]);
    }
    __syncwarp();
    for (int j = k + 1 + lane; j < n2; j += 32) work[k + j * n] = work[k + j * n] / p;
    __syncwarp();
  }
  for (int idx = lane; idx < n * n; idx += 32) {
    int i = idx % n, j = idx / n;
    A[i + j * lda] = work[i + (j + n) * n];
  }
  __syncwarp();
  return singular;
}

// In-place Cholesky A = L L^T reading only the lower triangle (Eigen::LLT semantics);
// on return the lower triangle holds L (upper triangle untouched).  Returns 1 if not SPD.
__device__ ISV_HEAVY int w_chol_lower(double* A, int ld, int n, int lane) {
  int bad = 0;
  for (int j = 0; j < n; ++j) {
    double d = A[j + j * ld];
    if (!(d > 0.0)) bad = 1;
    double r = sqrt(d);
    __syncwarp();
    if (lane == 0) A[j + j * ld] = r;
    for (int i = j + 1 + lane; i < n; i += 32) A[i + j * ld] = A[i + j * ld] / r;
    __syncwarp();
    // trailing update of the lower triangle: A[i][c] -= L[i][j] * L[c][j], i >= c > j
    int t = n - j - 1;
    for (int idx = lane; idx < t * t; idx += 32) {
      int i = j + 1 + idx % t, c = j + 1 + idx / t;
      if (i >= c) A[i + c * ld] = fma(-A[i + j * ld], A[c + j * ld], A[i + c * ld]);
    }
    __syncwarp();
  }
  return bad;
}

// out (n x n, ld ldo, column-major) = L^T where L is the lower triangle of A:
// Eigen `LLT(M).matrixL().transpose()` -- upper triangular, zeros below the diagonal.
__device__ __forceinline__ void w_store_upper_from_chol(const double* A, int ld, int n, double* out, int ldo,
                                                        int lane) {
  for (int idx = lane; idx < n * n; idx += 32) {
    int i = idx % n, j = idx / n;
    out[i + j * ldo] = (i <= j) ? A[j + i * ld] : 0.0;
  }
  __syncwarp();
}

// Symmetric eigen-decomposition by two-sided Jacobi with round-robin (parallel) ordering.
// A: n x n symmetric (full storage, ld lda), overwritten; eigenvalues end on its diagonal.
// V: n x n (ld ldv) receives the eigenvectors as columns (A_in = V diag V^T).
// cs: scratch of 6*ceil(n/2) doubles.  n <= 63.  Returns the number of sweeps used
// (>= max_sweeps -> not converged).
// Classical formulation: after each rotation the pivot pair is set to exactly zero and the two
// diagonal entries are updated with  a_pp -= t a_pq , a_qq += t a_pq , so rounding noise of size
// eps*|a_pp| is never re-injected into the off-diagonal.  A pair is skipped once
// |a_pq| <= 0.1 * eps * ||A||_F, ten times below what a backward-stable QR iteration guarantees.
__device__ __forceinline__ void jacobi_pair(int kp, int r, int m, int& p, int& q) {
  int a, b;
  if (kp == 0) { a = m - 1; b = r; }
  else { a = r + kp; if (a >= m - 1) a -= m - 1; b = r - kp; if (b < 0) b += m - 1; }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

__device__ ISV_HEAVY int w_jacobi_eig(double* A, int lda, double* V, int ldv, int n, double* cs, int lane,
                                            int max_sweeps = 30) {
  for (int idx = lane; idx < n * n; idx += 32) {
    int i = idx % n, j = idx / n;
    V[i + j * ldv] = (i == j) ? 1.0 : 0.0;
  }
  const int m = (n + 1) & ~1;  // players in the tournament (one dummy if n is odd)
  const int half = m / 2;
  double fro = 0.0;
  for (int idx = lane; idx < n * n; idx += 32) {
    double a = A[(idx % n) + (idx / n) * lda];
    fro = fma(a, a, fro);
  }
  fro = warp_sum(fro);
  __syncwarp();
  const double thr = 0.1 * 2.220446049250313e-16 * sqrt(fro);
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    int rotated = 0;
    for (int r = 0; r < m - 1; ++r) {
      // --- rotation parameters, one pair per lane ---
      for (int kp = lane; kp < half; kp += 32) {
        int p, q;
        jacobi_pair(kp, r, m, p, q);
        double c = 1.0, s = 0.0, app = 0.0, aqq = 0.0, tapq = 0.0;
        if (q < n) {
          double apq = A[p + q * lda];
          if (fabs(apq) > thr) {
            app = A[p + p * lda];
            aqq = A[q + q * lda];
            double tau = (aqq - app) / (2.0 * apq);
            double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            tapq = t * apq;
            rotated = 1;
          }
        }
        cs[6 * kp] = c;
        cs[6 * kp + 1] = s;
        cs[6 * kp + 2] = app - tapq;
        cs[6 * kp + 3] = aqq + tapq;
      }
      __syncwarp();
      // --- column update of A and V:  [x_p x_q] <- [c x_p - s x_q , s x_p + c x_q] ---
      for (int idx = lane; idx < half * n; idx += 32) {
        int kp = idx / n, i = idx - kp * n;
        int p, q;
        jacobi_pair(kp, r, m, p, q);
        double c = cs[6 * kp], s = cs[6 * kp + 1];
        if (q < n && s != 0.0) {
          double xp = A[i + p * lda], xq = A[i + q * lda];
          A[i + p * lda] = c * xp - s * xq;
          A[i + q * lda] = s * xp + c * xq;
          double vp = V[i + p * ldv], vq = V[i + q * ldv];
          V[i + p * ldv] = c * vp - s * vq;
          V[i + q * ldv] = s * vp + c * vq;
        }
      }
      __syncwarp();
      // --- row update of A ---
      for (int idx = lane; idx < half * n; idx += 32) {
        int kp = idx / n, j = idx - kp * n;
        int p, q;
        jacobi_pair(kp, r, m, p, q);
        double c = cs[6 * kp], s = cs[6 * kp + 1];
        if (q < n && s != 0.0) {
          double xp = A[p + j * lda], xq = A[q + j * lda];
          A[p + j * lda] = c * xp - s * xq;
          A[q + j * lda] = s * xp + c * xq;
        }
      }
      __syncwarp();
      // --- exact pivot block ---
      for (int kp = lane; kp < half; kp += 32) {
        int p, q;
        jacobi_pair(kp, r, m, p, q);
        if (q < n && cs[6 * kp + 1] != 0.0) {
          A[p + p * lda] = cs[6 * kp + 2];
          A[q + q * lda] = cs[6 * kp + 3];
          A[p + q * lda] = 0.0;
          A[q + p * lda] = 0.0;
        }
      }
      __syncwarp();
    }
    if (!__any_sync(kFullMask, rotated)) break;
  }
  return sweep;
}

// -------------------------------------------------------------------------------------------------
// PSD eigen-decomposition in factored form: pivoted Cholesky + one-sided (Hestenes) Jacobi.
//
// Replaces SelfAdjointEigenSolver on the marginal information matrices of this path
// (/root/reference/src/estimator.cpp:920, :1311, :1479), which are PSD up to rounding (Schur
// complements of J^T W J).  Step 1 factors A ~= G^T G with G (r x n) by outer-product Cholesky with
// diagonal pivoting, stopping at pivots <= n*eps*max_diag (what is dropped is rounding noise many
// orders below ALPHA; negative rounding-level eigenvalues disappear with it).  Step 2 rotates the
// ROWS of G until they are mutually orthogonal; then  A = sum_k g_k g_k^T  with eigenvalues
// lam_k = |g_k|^2 and eigenvectors g_k/|g_k|, so that the reference's
//     U D^-1 U^T = sum_{lam_k > ALPHA} g_k g_k^T / lam_k^2 .
// One-sided Jacobi works on r x n (15 x 21 for MargBackward) instead of n x n, needs no eigenvector
// accumulation, and is accurate to high RELATIVE precision for every kept eigenvalue.
// -------------------------------------------------------------------------------------------------

// A: n x n symmetric (lower triangle read, ld lda) -- destroyed.  G: receives r rows (row-major,
// ld ldg >= n).  d: n doubles scratch.  Returns r.
__device__ ISV_HEAVY int w_pivoted_cholesky_rows(double* A, int lda, int n, double* G, int ldg, double* d,
                                                       int lane) {
  w_symmetrize_from_lower(A, lda, n, lane);
  double dmax0 = 0.0;
  for (int i = lane; i < n; i += 32) {
    double v = A[i + i * lda];
    d[i] = v;
    dmax0 = fmax(dmax0, v);
  }
  dmax0 = warp_max(dmax0);
  __syncwarp();
  const double tol = dmax0 * (double)n * 2.220446049250313e-16;
  int r = 0;
  for (; r < n; ++r) {
    // pivot = largest remaining diagonal (selected indices carry d = -1)
    double best = -1.0;
    int bi = 0;
    for (int i = lane; i < n; i += 32)
      if (d[i] > best) { best = d[i]; bi = i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(kFullMask, best, o);
      int oi = __shfl_xor_sync(kFullMask, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (!(best > tol)) break;
    const double inv = 1.0 / sqrt(best);
    double* g = G + r * ldg;
    for (int i = lane; i < n; i += 32) g[i] = (d[i] < 0.0) ? 0.0 : A[i + bi * lda] * inv;
    __syncwarp();
    for (int idx = lane; idx < n * n; idx += 32) {
      int i = idx % n, j = idx / n;
      A[i + j * lda] = fma(-g[i], g[j], A[i + j * lda]);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) d[i] = (d[i] < 0.0 || i == bi) ? -1.0 : A[i + i * lda];
    __syncwarp();
  }
  return r;
}

// Rotation that orthogonalises two rows with squared norms a, b and inner product g:
//   p' = c p - s q ,  q' = s p + c q ,  tan(2 theta) = 2 g / (b - a), |theta| <= pi/4.
// cos(2 theta) = |d|/r, r = sqrt(d^2 + 4 g^2), d = b - a;  c = sqrt((1 + cos 2theta)/2),
// s = sign(d) g / (r c).  Two rsqrt instead of three divisions and two square roots.
__device__ __forceinline__ void jacobi_cs(double a, double b, double g, double& c, double& s) {
  const double d = b - a;
  const double inv_r = rsqrt(fma(d, d, 4.0 * g * g));
  const double h = fma(0.5 * fabs(d), inv_r, 0.5);  // cos^2(theta) in [1/2, 1]
  const double inv_c = rsqrt(h);
  c = h * inv_c;
  s = (d >= 0.0 ? g : -g) * inv_r * inv_c;
}

// Orthogonalise the r rows (n elements each; element e of row p at G[p*ldg + e*es]) of G by
// one-sided Jacobi with round-robin ordering; LPP lanes cooperate on one row pair and keep their
// NE = ceil(n/LPP) elements of both rows in registers between the inner product and the update.
// Squared row norms are cached in nrm[] (updated analytically by each rotation, recomputed exactly
// at the start of every sweep -- the dgesvj strategy).  lam[k] = |g_k|^2 on return (lam may alias
// nrm).  Returns the number of sweeps used (>= max_sweeps -> not converged).
template <int LPP, int NE>
__device__ ISV_HEAVY int w_onesided_jacobi_rows(double* G, int ldg, int r, int n, double* lam, int lane,
                                                      int max_sweeps = 30, int es = 1) {
  constexpr int kGroups = 32 / LPP;
  const int sub = lane % LPP, grp = lane / LPP;
  const int m = (r + 1) & ~1;
  const int half = m / 2;
  const double tol2 = 2.220446049250313e-16 * 2.220446049250313e-16 * (double)n;
  double* nrm = lam;
  int sweep = 0;
  for (; sweep < max_sweeps && r >= 2; ++sweep) {
    // exact squared norms
    for (int k = lane; k < r; k += 32) {
      double a = 0.0;
      for (int e = 0; e < n; ++e) a = fma(G[k * ldg + e * es], G[k * ldg + e * es], a);
      nrm[k] = a;
    }
    __syncwarp();
    int rotated = 0;
    for (int rd = 0; rd < m - 1; ++rd) {
      for (int kb = 0; kb < half; kb += kGroups) {
        const int kp = kb + grp;
        int p = 0, q = 0;
        bool act = kp < half;
        if (act) {
          int a, b;
          if (kp == 0) { a = m - 1; b = rd; }
          else { a = rd + kp; if (a >= m - 1) a -= m - 1; b = rd - kp; if (b < 0) b += m - 1; }
          p = a < b ? a : b;
          q = a < b ? b : a;
          act = q < r;
        }
        double* gp = G + p * ldg + sub * es;
        double* gq = G + q * ldg + sub * es;
        double x[NE > 0 ? NE : 1], y[NE > 0 ? NE : 1];
        double gg = 0.0;
        if constexpr (NE > 0) {
#pragma unroll
          for (int i = 0; i < NE; ++i) {
            const bool in = act && (sub + i * LPP < n);
            x[i] = in ? gp[i * LPP * es] : 0.0;
            y[i] = in ? gq[i * LPP * es] : 0.0;
            gg = fma(x[i], y[i], gg);
          }
        } else {  // rows stay in shared memory (long rows: initFactorGraph's 42 x 57 factor)
          if (act)
            for (int e = sub; e < n; e += LPP) gg = fma(gp[(e - sub) * es], gq[(e - sub) * es], gg);
        }
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) gg += __shfl_xor_sync(kFullMask, gg, o);
        const double aa = nrm[p], bb = nrm[q];
        if (act && gg * gg > tol2 * aa * bb) {
          double c, s;
          jacobi_cs(aa, bb, gg, c, s);
          if constexpr (NE > 0) {
#pragma unroll
            for (int i = 0; i < NE; ++i)
              if (sub + i * LPP < n) {
                gp[i * LPP * es] = c * x[i] - s * y[i];
                gq[i * LPP * es] = s * x[i] + c * y[i];
              }
          } else {
            for (int e = sub; e < n; e += LPP) {
              const double xv = gp[(e - sub) * es], yv = gq[(e - sub) * es];
              gp[(e - sub) * es] = c * xv - s * yv;
              gq[(e - sub) * es] = s * xv + c * yv;
            }
          }
          if (sub == 0) {
            const double cc = c * c, ss = s * s, csg = 2.0 * c * s * gg;
            nrm[p] = cc * aa - csg + ss * bb;
            nrm[q] = ss * aa + csg + cc * bb;
          }
          rotated = 1;
        }
      }
      __syncwarp();
    }
    if (!__any_sync(kFullMask, rotated)) break;
  }
  for (int k = lane; k < r; k += 32) {
    double a = 0.0;
    for (int e = 0; e < n; ++e) a = fma(G[k * ldg + e * es], G[k * ldg + e * es], a);
    lam[k] = a;
  }
  __syncwarp();
  return sweep;
}

// -------------------------------------------------------------------------------------------------
// Square-root-information marginalisation: Householder QR on the columns [c0, c0+nc) of the stacked
// whitened Jacobian M (rows x cols, column-major, ld).  On return rows [nc, rows) hold, in the
// columns outside [c0, c0+nc), a factor Gm with  Gm^T Gm = Lam_rr - Lam_rm Lam_mm^-1 Lam_mr  where
// Lam = M^T M -- the Schur complement of /root/reference/src/estimator.cpp:1413-1419 (and :808-816)
// without ever forming Lam or inverting Lam_mm, so no cancellation noise enters the null space.
// vbuf: rows doubles of scratch.
// -------------------------------------------------------------------------------------------------
__device__ ISV_HEAVY void w_householder_marginalize(double* M, int ld, int rows, int cols, int c0, int nc,
                                                          double* vbuf, int lane) {
  for (int k = 0; k < nc; ++k) {
    double* ck = M + (size_t)ld * (c0 + k);
    // reflector for ck[k:rows]
    double t2 = 0.0;
    for (int i = k + 1 + lane; i < rows; i += 32) t2 = fma(ck[i], ck[i], t2);
    t2 = warp_sum(t2);
    const double x0 = ck[k];
    __syncwarp();
    if (t2 > 0.0) {
      double beta = sqrt(x0 * x0 + t2);
      if (x0 >= 0.0) beta = -beta;
      const double inv = 1.0 / (x0 - beta);
      const double tau = (beta - x0) / beta;
      for (int i = k + lane; i < rows; i += 32) vbuf[i] = (i == k) ? 1.0 : ck[i] * inv;
      __syncwarp();
      // apply (I - tau v v^T) to every other live column: one column per lane
      for (int j = lane; j < cols; j += 32) {
        if (j >= c0 && j <= c0 + k) continue;  // eliminated columns (and ck itself, set below)
        double* cj = M + (size_t)ld * j;
        double s = 0.0;
        for (int i = k; i < rows; ++i) s = fma(vbuf[i], cj[i], s);
        s *= tau;
        for (int i = k; i < rows; ++i) cj[i] = fma(-s, vbuf[i], cj[i]);
      }
      if (lane == 0) ck[k] = beta;
      for (int i = k + 1 + lane; i < rows; i += 32) ck[i] = 0.0;
      __syncwarp();
    }
  }
}

// sqrt_info = LLT(cov^-1).matrixL().transpose() computed without forming cov^-1:
// cov = U1 U1^T with U1 upper triangular (Cholesky started from the last row), then
// cov^-1 = (U1^-1)^T (U1^-1) and U1^-1 is the unique upper-triangular factor with positive diagonal,
// i.e. exactly the matrix the reference obtains as  Eigen::LLT(covi.inverse()).matrixL().transpose()
// (/root/reference/src/estimator.cpp:950,1349,1503,1508,1515).
// A: N x N SPD in shared memory (upper triangle read, destroyed); out: N x N column-major (global).
template <int N>
__device__ ISV_HEAVY int w_sqrt_info_from_cov(double* A, int ld, double* out, int lane, int& nonfinite) {
  int bad = 0;
  for (int j = N - 1; j >= 0; --j) {
    double d = A[j + j * ld];
    if (!(d > 0.0)) bad = 1;
    double r = sqrt(d);
    __syncwarp();
    if (lane == 0) A[j + j * ld] = r;
    for (int i = lane; i < j; i += 32) A[i + j * ld] = A[i + j * ld] / r;
    __syncwarp();
    for (int idx = lane; idx < j * j; idx += 32) {
      int i = idx % j, c = idx / j;
      if (i <= c) A[i + c * ld] = fma(-A[i + j * ld], A[c + j * ld], A[i + c * ld]);
    }
    __syncwarp();
  }
  // column `lane` of X = U1^-1 by back substitution, held in registers
  if (lane < N) {
    double x[N];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = i + 1; k < N; ++k) s = fma(-A[i + k * ld], x[k], s);
      x[i] = (i <= lane) ? s / A[i + i * ld] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (!isfinite(x[i])) nonfinite = 1;
      out[i + N * lane] = x[i];
    }
  }
  __syncwarp();
  return bad;
}

// SPD inverse without pivot searches: A = U1 U1^T (reverse-order Cholesky), X = U1^-1 (one column per
// lane, in registers), A^-1 = X^T X.  A (smem, ld) is overwritten by A^-1; Xs: N*N doubles scratch.
// Returns 1 when A is not SPD (result then meaningless: the caller falls back to w_inverse).
template <int N>
__device__ ISV_HEAVY int w_spd_inverse(double* A, int ld, double* Xs, int lane) {
  int bad = 0;
  for (int j = N - 1; j >= 0; --j) {
    double d = A[j + j * ld];
    if (!(d > 0.0)) bad = 1;
    double r = sqrt(d);
    __syncwarp();
    if (lane == 0) A[j + j * ld] = r;
    for (int i = lane; i < j; i += 32) A[i + j * ld] = A[i + j * ld] / r;
    __syncwarp();
    for (int idx = lane; idx < j * j; idx += 32) {
      int i = idx % j, c = idx / j;
      if (i <= c) A[i + c * ld] = fma(-A[i + j * ld], A[c + j * ld], A[i + c * ld]);
    }
    __syncwarp();
  }
  if (lane < N) {
    double x[N];
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = i + 1; k < N; ++k) s = fma(-A[i + k * ld], x[k], s);
      x[i] = (i <= lane) ? s / A[i + i * ld] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) Xs[i + N * lane] = x[i];
  }
  __syncwarp();
  for (int idx = lane; idx < N * N; idx += 32) {
    int i = idx % N, j = idx / N;
    double acc = 0.0;
    // (X^T X)[i][j] = sum_k X[k][i] X[k][j], X upper triangular: k <= min(i, j)
    const int km = i < j ? i : j;
    for (int k = 0; k <= km; ++k) acc = fma(Xs[k + N * i], Xs[k + N * j], acc);
    A[i + j * ld] = acc;
  }
  __syncwarp();
  return bad;
}

// Register-resident form of w_sqrt_info_from_cov (same result: the upper-triangular U1^-1 with
// cov = U1 U1^T): every lane loads the upper triangle (broadcast LDS), runs the reverse-order
// Cholesky serially in registers (no barriers, no shuffles; SIMT makes the redundancy free), then
// lane c < N back-substitutes column c of U1^-1 and stores it.  A is only read.
template <int N>
__device__ __forceinline__ int w_sqrt_info_from_cov_regs(const double* A, int ld, double* out, int lane,
                                                         int& nonfinite) {
  double u[N][N];   // upper triangle used: u[i][j], i <= j
  double rinv[N];
  int bad = 0;
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i <= j; ++i) u[i][j] = A[i + j * ld];
#pragma unroll
  for (int j = N - 1; j >= 0; --j) {
    const double d = u[j][j];
    if (!(d > 0.0)) bad = 1;
    const double ri = fast_rsqrt(d);
    rinv[j] = ri;
#pragma unroll
    for (int i = 0; i < j; ++i) u[i][j] *= ri;
#pragma unroll
    for (int c = 0; c < j; ++c)
#pragma unroll
      for (int i = 0; i <= c; ++i) u[i][c] = fma(-u[i][j], u[c][j], u[i][c]);
  }
  // U1 has diagonal sqrt(d_j) = 1 / rinv[j]; column `lane` of X = U1^-1: x_i = (e_i - sum_{k>i} U_ik x_k) rinv_i
  double x[N];
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = i + 1; k < N; ++k) sacc = fma(-u[i][k], x[k], sacc);
    x[i] = (i <= lane) ? sacc * rinv[i] : 0.0;
  }
  if (lane < N) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (!isfinite(x[i])) nonfinite = 1;
      out[i + N * lane] = x[i];
    }
  }
  __syncwarp();
  return bad;
}

// Eigen `LLT(M).matrixL().transpose()` of a small SPD matrix entirely in registers: every lane loads the
// lower triangle (broadcast LDS) and runs the same serial Cholesky (SIMT makes the redundancy free; one
// rsqrt per column, no barriers), lane j < N stores column j of the upper-triangular factor (zeros below
// the diagonal).  M (smem, ld) is only read.  out: N x N column-major.  Returns 1 when M is not SPD.
template <int N>
__device__ __forceinline__ int w_llt_upper_regs(const double* M, int ld, double* out, int lane, int& nonfinite) {
  double l[N][N];
  int bad = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) l[i][j] = M[i + j * ld];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = l[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma(-l[j][k], l[j][k], d);
    if (!(d > 0.0)) bad = 1;
    const double ri = fast_rsqrt(d);
    l[j][j] = d * ri;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double v = l[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-l[i][k], l[j][k], v);
      l[i][j] = v * ri;
    }
  }
  if (lane < N) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double v = 0.0;   // U[i][lane] = L[lane][i] for i <= lane
#pragma unroll
      for (int c = 0; c < N; ++c)
        if (c == lane && i <= c) v = l[c][i];
      if (!isfinite(v)) nonfinite = 1;
      out[i + N * lane] = v;
    }
  }
  __syncwarp();
  return bad;
}

// Several small SPD matrices turned into their sqrt-information factors AT ONCE, one column per lane:
// lane (group g, column c) owns column c of its group's covariance A_g (N_g x N_g, column-major, ld N_g)
// in registers; the reverse-order Cholesky A = U1 U1^T runs right-looking with the finished column
// broadcast through the group's scratch Us (N*N + N doubles), then every lane back-substitutes its own
// column of U1^-1 (the matrix w_sqrt_info_from_cov produces) and stores it.  The groups share one
// instruction stream, so three matrices cost what the largest one costs.  All 32 lanes must call it;
// `active` = this lane owns a column.  NMAX >= every N.
template <int NMAX>
__device__ __forceinline__ int w_sqrt_info_multi(const double* A, int N, int c, bool active, double* Us, double* out,
                                                 int& nonfinite) {
  double a[NMAX];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) a[i] = (active && i <= c) ? A[i + N * c] : 0.0;
  int bad = 0;
#pragma unroll
  for (int j = NMAX - 1; j >= 0; --j) {
    if (active && c == j) {
      const double d = a[j];
      if (!(d > 0.0)) bad = 1;
      const double ri = fast_rsqrt(d);
      Us[N * N + j] = ri;
#pragma unroll
      for (int i = 0; i < NMAX; ++i)
        if (i < j) Us[j * N + i] = a[i] * ri;
    }
    __syncwarp();
    if (active && c < j && j < N) {
      const double ucj = Us[j * N + c];
#pragma unroll
      for (int i = 0; i < NMAX; ++i)
        if (i <= c) a[i] = fma(-Us[j * N + i], ucj, a[i]);
    }
  }
  __syncwarp();
  // column c of X = U1^-1 : x_i = ((i == c) - sum_{i < k <= c} U1[i][k] x_k) / U1[i][i]
  double x[NMAX];
#pragma unroll
  for (int i = NMAX - 1; i >= 0; --i) {
    double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int k = i + 1; k < NMAX; ++k)
      if (active && k <= c) sacc = fma(-Us[k * N + i], x[k], sacc);
    x[i] = (active && i <= c) ? sacc * Us[N * N + i] : 0.0;
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
      if (i < N) {
        if (!isfinite(x[i])) nonfinite = 1;
        out[i + N * c] = x[i];
      }
  }
  __syncwarp();
  return bad;
}

// SPD inverse of a small matrix entirely in registers, no shuffles and no barriers inside: EVERY lane
// loads the lower triangle (broadcast LDS) and runs the same serial LDL^T factorisation (SIMT makes the
// redundancy free), then lane j < N solves for column j of A^-1 and writes it back.  ~4x fewer warp
// instructions than the cooperative w_spd_inverse for N = 6, and a much shorter dependent chain.
// A (smem, ld) is overwritten by A^-1 (full, symmetric).  Returns 1 when A is not SPD.
template <int N>
__device__ __forceinline__ int w_spd_inverse_regs(double* A, int ld, int lane) {
  double l[N][N];   // unit lower factor (strict lower part used), d on the diagonal slot
  double dinv[N];
  int bad = 0;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) l[i][j] = A[i + j * ld];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double d = l[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma(-l[j][k] * l[j][k], l[k][k], d);   // l[k][k] holds d_k
    if (!(d > 0.0)) bad = 1;
    l[j][j] = d;
    dinv[j] = fast_rcp(d);
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      double v = l[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v = fma(-l[i][k] * l[k][k], l[j][k], v);
      l[i][j] = v * dinv[j];
    }
  }
  // column `lane` of A^-1:  L y = e ;  z = D^-1 y ;  L^T x = z
  double x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double v = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) v = fma(-l[i][k], x[k], v);
    x[i] = v;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] *= dinv[i];
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double v = x[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) v = fma(-l[k][i], x[k], v);
    x[i] = v;
  }
  if (lane < N) {
#pragma unroll
    for (int i = 0; i < N; ++i) A[i + lane * ld] = x[i];
  }
  __syncwarp();
  return bad;
}

}  // namespace isv
