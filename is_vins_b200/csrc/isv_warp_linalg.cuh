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
__device__ __forceinline__ void w_gemm(int m, int n, int k, const double* A, int lda, const double* B, int ldb,
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
__device__ __forceinline__ int w_inverse(double* A, int lda, int n, double* work, int lane) {
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
__device__ __forceinline__ int w_chol_lower(double* A, int ld, int n, int lane) {
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

__device__ __forceinline__ int w_jacobi_eig(double* A, int lda, double* V, int ldv, int n, double* cs, int lane,
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

}  // namespace isv
