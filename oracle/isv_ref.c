/*
 * isv_ref.c -- plain-C CPU restatement of IS-VINS MargForward + MargBackward.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/isv_oracle.py header): used by tests/ as a second,
 * independent restatement of the reference algorithm and by bench.py as the reported CPU baseline
 * ("cpu_baseline.kind = port", `--impl reference`).  The product (is_vins_b200/) never links,
 * loads or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests/golden vectors for this path and cannot be built
 * here (Eigen 3.3.4 / Ceres 2.0.0 / Sophus / OpenCV are neither vendored nor installed), so this is
 * a restatement, function by function, of
 *   /root/reference/src/estimator.cpp:1149-1352   Estimator::MargForward
 *   /root/reference/src/estimator.cpp:1354-1539   Estimator::MargBackward
 *   /root/reference/src/factor/projection_factor.cpp:124-196, include/factor/imu_factor.h:161-265,
 *   include/factor/{relative_pose,se3_prior,rollpitch}_factor.h, include/utility/utility.h,
 *   include/utility/sophus_utils.hpp:194-236
 * with the same ALGORITHM CLASSES as the un-vendored Eigen calls: dense column-major Lamda built by
 * the block loop, FullPivLU + identity solve for Lamda_mm^-1 (O((L+6)^3), the reference's dominant
 * cost), PartialPivLU inverse, LLT, FullPivHouseholderQR rank/solve, a tridiagonal-QL symmetric
 * eigensolver (SelfAdjointEigenSolver) and a one-sided-Jacobi thin SVD (BDCSVD) pseudo-inverse.
 * The reference is single-threaded per window; threads here run ACROSS windows (OpenMP).
 * `structured != 0` eliminates the diagonal landmark block first (same result, O(L)); it is
 * reported separately and never called "the reference".
 *
 * I/O uses the product's SoA layout (include/isv_capi.h) with HOST pointers.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/isv_capi.h"

#define EPS_D 2.220446049250313e-16
#define SOPHUS_EPS 1e-10
#define SOPHUS_EPS_SQRT 1e-5
#define A_(M, ld, i, j) (M)[(size_t)(i) + (size_t)(ld) * (size_t)(j)]

/* ------------------------------------------------------------------------------------------
 * small dense helpers (column-major)
 * ---------------------------------------------------------------------------------------- */
/* C (m x n) = / += / -= op(A) op(B); ta/tb: use transpose; mode 0 set, 1 add, -1 sub */
static void mm(int m, int n, int k, const double* A, int lda, int ta, const double* B, int ldb, int tb, double* C,
               int ldc, int mode) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double acc = 0.0;
      for (int l = 0; l < k; ++l) {
        double a = ta ? A_(A, lda, l, i) : A_(A, lda, i, l);
        double b = tb ? A_(B, ldb, j, l) : A_(B, ldb, l, j);
        acc += a * b;
      }
      if (mode == 0) A_(C, ldc, i, j) = acc;
      else if (mode > 0) A_(C, ldc, i, j) += acc;
      else A_(C, ldc, i, j) -= acc;
    }
}

/* Eigen::FullPivLU(A).solve(Identity): X (n x n).  A is destroyed.  Returns the rank. */
static int fullpiv_lu_solve_identity(int n, double* A, double* X) {
  int* rowp = (int*)malloc(sizeof(int) * (size_t)n * 2);
  int* colp = rowp + n;
  for (int i = 0; i < n; ++i) { rowp[i] = i; colp[i] = i; }
  double maxpivot = 0.0;
  int nonzero = n;
  for (int k = 0; k < n; ++k) {
    int pr = k, pc = k;
    double big = -1.0;
    for (int c = k; c < n; ++c) {
      const double* col = A + (size_t)n * c;
      for (int r = k; r < n; ++r) {
        double v = fabs(col[r]);
        if (v > big) { big = v; pr = r; pc = c; }
      }
    }
    if (big == 0.0) { nonzero = k; break; }
    if (big > maxpivot) maxpivot = big;
    if (pr != k) {
      for (int c = 0; c < n; ++c) { double t = A_(A, n, k, c); A_(A, n, k, c) = A_(A, n, pr, c); A_(A, n, pr, c) = t; }
      int t = rowp[k]; rowp[k] = rowp[pr]; rowp[pr] = t;
    }
    if (pc != k) {
      double* c1 = A + (size_t)n * k; double* c2 = A + (size_t)n * pc;
      for (int r = 0; r < n; ++r) { double t = c1[r]; c1[r] = c2[r]; c2[r] = t; }
      int t = colp[k]; colp[k] = colp[pc]; colp[pc] = t;
    }
    if (k < n - 1) {
      double* ck = A + (size_t)n * k;
      double inv = ck[k];
      for (int r = k + 1; r < n; ++r) ck[r] /= inv;
      for (int c = k + 1; c < n; ++c) {
        double* cc = A + (size_t)n * c;
        double f = cc[k];
        if (f != 0.0)
          for (int r = k + 1; r < n; ++r) cc[r] -= ck[r] * f;
      }
    }
  }
  int rank = 0;
  double thr = EPS_D * n;
  for (int k = 0; k < nonzero; ++k)
    if (fabs(A_(A, n, k, k)) > maxpivot * thr) ++rank;
  /* X = Q * U^-1 * L^-1 * P * I */
  double* c = (double*)malloc(sizeof(double) * (size_t)n);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) c[i] = (rowp[i] == j) ? 1.0 : 0.0;
    for (int k = 0; k < n; ++k) { /* unit-lower forward substitution, column oriented */
      double f = c[k];
      if (f != 0.0) {
        const double* ck = A + (size_t)n * k;
        for (int r = k + 1; r < n; ++r) c[r] -= ck[r] * f;
      }
    }
    for (int k = rank - 1; k >= 0; --k) {
      const double* ck = A + (size_t)n * k;
      c[k] /= ck[k];
      double f = c[k];
      for (int r = 0; r < k; ++r) c[r] -= ck[r] * f;
    }
    double* xj = X + (size_t)n * j;
    for (int i = 0; i < n; ++i) xj[i] = 0.0;
    for (int i = 0; i < rank; ++i) xj[colp[i]] = c[i];
  }
  free(c);
  free(rowp);
  return rank;
}

/* MatrixXd::inverse() (PartialPivLU): Gauss-Jordan with row pivoting, n <= 32.  A -> A^-1 */
static void partial_piv_inverse(int n, double* A) {
  double W[32 * 64];
  int n2 = 2 * n;
  for (int j = 0; j < n2; ++j)
    for (int i = 0; i < n; ++i) W[i + n * j] = j < n ? A_(A, n, i, j) : ((j - n) == i ? 1.0 : 0.0);
  for (int k = 0; k < n; ++k) {
    int p = k;
    double big = fabs(W[k + n * k]);
    for (int r = k + 1; r < n; ++r)
      if (fabs(W[r + n * k]) > big) { big = fabs(W[r + n * k]); p = r; }
    if (p != k)
      for (int j = 0; j < n2; ++j) { double t = W[k + n * j]; W[k + n * j] = W[p + n * j]; W[p + n * j] = t; }
    double piv = W[k + n * k];
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      double f = W[i + n * k] / piv;
      for (int j = k + 1; j < n2; ++j) W[i + n * j] -= f * W[k + n * j];
    }
    for (int j = k + 1; j < n2; ++j) W[k + n * j] /= piv;
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) A_(A, n, i, j) = W[i + n * (j + n)];
}

/* Eigen::LLT(M).matrixL().transpose() -> U (upper, zeros below).  Reads the lower triangle. */
static int llt_upper(int n, const double* M, double* U) {
  double Lm[32 * 32];
  int bad = 0;
  memset(Lm, 0, sizeof(double) * (size_t)n * n);
  for (int j = 0; j < n; ++j) {
    double d = A_(M, n, j, j);
    for (int l = 0; l < j; ++l) d -= Lm[j + n * l] * Lm[j + n * l];
    if (!(d > 0.0)) bad = 1;
    double r = sqrt(d);
    Lm[j + n * j] = r;
    for (int i = j + 1; i < n; ++i) {
      double s = A_(M, n, i, j);
      for (int l = 0; l < j; ++l) s -= Lm[i + n * l] * Lm[j + n * l];
      Lm[i + n * j] = s / r;
    }
  }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) A_(U, n, i, j) = (i <= j) ? Lm[j + n * i] : 0.0;
  return bad;
}

/* SelfAdjointEigenSolver: Householder tridiagonalisation + implicit QL (EISPACK tred2/tql2 class).
 * A (n x n, lower triangle read) -> eigenvalues w ascending, eigenvectors in the columns of Z. */
static double hyp(double a, double b) { return hypot(a, b); }
static void sym_eig(int n, const double* Ain, double* w, double* Z) {
  double e[64];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) A_(Z, n, i, j) = (i >= j) ? A_(Ain, n, i, j) : A_(Ain, n, j, i);
  /* tred2 */
  for (int j = 0; j < n; ++j) w[j] = A_(Z, n, n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += fabs(w[k]);
    if (scale == 0.0) {
      e[i] = w[i - 1];
      for (int j = 0; j < i; ++j) { w[j] = A_(Z, n, i - 1, j); A_(Z, n, i, j) = 0.0; A_(Z, n, j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { w[k] /= scale; h += w[k] * w[k]; }
      double f = w[i - 1];
      double g = sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      w[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = w[j];
        A_(Z, n, j, i) = f;
        g = e[j] + A_(Z, n, j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += A_(Z, n, k, j) * w[k]; e[k] += A_(Z, n, k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * w[j]; }
      double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * w[j];
      for (int j = 0; j < i; ++j) {
        f = w[j];
        g = e[j];
        for (int k = j; k <= i - 1; ++k) A_(Z, n, k, j) -= (f * e[k] + g * w[k]);
        w[j] = A_(Z, n, i - 1, j);
        A_(Z, n, i, j) = 0.0;
      }
    }
    w[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    A_(Z, n, n - 1, i) = A_(Z, n, i, i);
    A_(Z, n, i, i) = 1.0;
    double h = w[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) w[k] = A_(Z, n, k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += A_(Z, n, k, i + 1) * A_(Z, n, k, j);
        for (int k = 0; k <= i; ++k) A_(Z, n, k, j) -= g * w[k];
      }
    }
    for (int k = 0; k <= i; ++k) A_(Z, n, k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { w[j] = A_(Z, n, n - 1, j); A_(Z, n, n - 1, j) = 0.0; }
  A_(Z, n, n - 1, n - 1) = 1.0;
  e[0] = 0.0;
  /* tql2 */
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  for (int l = 0; l < n; ++l) {
    double t = fabs(w[l]) + fabs(e[l]);
    if (t > tst1) tst1 = t;
    int m = l;
    while (m < n) {
      if (fabs(e[m]) <= EPS_D * tst1) break;
      ++m;
    }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = w[l];
        double p = (w[l + 1] - g) / (2.0 * e[l]);
        double r = hyp(p, 1.0);
        if (p < 0) r = -r;
        w[l] = e[l] / (p + r);
        w[l + 1] = e[l] * (p + r);
        double dl1 = w[l + 1];
        double h = g - w[l];
        for (int i = l + 2; i < n; ++i) w[i] -= h;
        f += h;
        p = w[m];
        double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = hyp(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * w[i] - s * g;
          w[i + 1] = h + s * (c * g + s * w[i]);
          for (int k = 0; k < n; ++k) {
            h = A_(Z, n, k, i + 1);
            A_(Z, n, k, i + 1) = s * A_(Z, n, k, i) + c * h;
            A_(Z, n, k, i) = c * A_(Z, n, k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        w[l] = c * p;
      } while (fabs(e[l]) > EPS_D * tst1 && iter < 60);
    }
    w[l] += f;
    e[l] = 0.0;
  }
  for (int i = 0; i < n - 1; ++i) { /* ascending order */
    int k = i;
    double p = w[i];
    for (int j = i + 1; j < n; ++j)
      if (w[j] < p) { k = j; p = w[j]; }
    if (k != i) {
      w[k] = w[i];
      w[i] = p;
      for (int j = 0; j < n; ++j) { double t = A_(Z, n, j, i); A_(Z, n, j, i) = A_(Z, n, j, k); A_(Z, n, j, k) = t; }
    }
  }
}

/* Eigen::FullPivHouseholderQR(A) with setThreshold(thr): rank; X = solve(I) when full rank */
static int fullpiv_qr_inverse(int n, const double* Ain, double thr, double* X) {
  double A[36], hc[6], c[6];
  int rowt[6], colp[6];
  memcpy(A, Ain, sizeof(double) * (size_t)n * n);
  double precision = EPS_D * n, biggest = 0.0, maxpivot = 0.0;
  int nonzero = n;
  for (int k = 0; k < n; ++k) colp[k] = k;
  for (int k = 0; k < n; ++k) {
    int pr = k, pc = k;
    double big = -1.0;
    for (int cc = k; cc < n; ++cc)
      for (int r = k; r < n; ++r) {
        double v = fabs(A[r + n * cc]);
        if (v > big) { big = v; pr = r; pc = cc; }
      }
    if (k == 0) biggest = big;
    if (big <= biggest * precision || big == 0.0) {
      nonzero = k;
      for (int i = k; i < n; ++i) { rowt[i] = i; hc[i] = 0.0; }
      break;
    }
    rowt[k] = pr;
    if (pr != k)
      for (int cc = k; cc < n; ++cc) { double t = A[k + n * cc]; A[k + n * cc] = A[pr + n * cc]; A[pr + n * cc] = t; }
    if (pc != k) {
      for (int r = 0; r < n; ++r) { double t = A[r + n * k]; A[r + n * k] = A[r + n * pc]; A[r + n * pc] = t; }
      int t = colp[k]; colp[k] = colp[pc]; colp[pc] = t;
    }
    double c0 = A[k + n * k], tail2 = 0.0, beta, tau;
    for (int r = k + 1; r < n; ++r) tail2 += A[r + n * k] * A[r + n * k];
    if (tail2 <= 2.2250738585072014e-308) {
      tau = 0.0;
      beta = c0;
      for (int r = k + 1; r < n; ++r) A[r + n * k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      for (int r = k + 1; r < n; ++r) A[r + n * k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k + n * k] = beta;
    hc[k] = tau;
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
    if (tau != 0.0)
      for (int cc = k + 1; cc < n; ++cc) {
        double s = A[k + n * cc];
        for (int r = k + 1; r < n; ++r) s += A[r + n * k] * A[r + n * cc];
        s *= tau;
        A[k + n * cc] -= s;
        for (int r = k + 1; r < n; ++r) A[r + n * cc] -= s * A[r + n * k];
      }
  }
  int rank = 0;
  for (int k = 0; k < nonzero; ++k)
    if (fabs(A[k + n * k]) > maxpivot * thr) ++rank;
  if (rank < n) return rank;
  for (int j = 0; j < n; ++j) {
    for (int r = 0; r < n; ++r) c[r] = (r == j) ? 1.0 : 0.0;
    for (int k = 0; k < n; ++k) {
      int r = rowt[k];
      if (r != k) { double t = c[k]; c[k] = c[r]; c[r] = t; }
      double s = c[k];
      for (int i = k + 1; i < n; ++i) s += A[i + n * k] * c[i];
      s *= hc[k];
      c[k] -= s;
      for (int i = k + 1; i < n; ++i) c[i] -= s * A[i + n * k];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = c[i];
      for (int l = i + 1; l < n; ++l) s -= A[i + n * l] * c[l];
      c[i] = s / A[i + n * i];
    }
    for (int i = 0; i < n; ++i) X[colp[i] + n * j] = c[i];
  }
  return rank;
}

/* Utility::pseudoInverse(J, epsilon) for J (6 x 12): thin SVD by one-sided Jacobi on J^T (12 x 6),
 * Eigen's relative threshold epsilon*max(rows,cols)*sigma_max.  Jp: 12 x 6. */
static void pseudo_inverse_6x12(const double* J, double epsilon, double* Jp) {
  double G[72], Vr[36], sig[6]; /* G = J^T (12x6, ld 12): columns get orthogonalised; J^T = G_final Vr^T */
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 12; ++j) G[j + 12 * i] = J[i + 6 * j];
  for (int i = 0; i < 36; ++i) Vr[i] = 0.0;
  for (int i = 0; i < 6; ++i) Vr[i + 6 * i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rot = 0;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        double a = 0, b = 0, g = 0;
        for (int k = 0; k < 12; ++k) { a += G[k + 12 * p] * G[k + 12 * p]; b += G[k + 12 * q] * G[k + 12 * q]; g += G[k + 12 * p] * G[k + 12 * q]; }
        if (fabs(g) <= EPS_D * sqrt(a * b) || g == 0.0) continue;
        rot = 1;
        double zeta = (b - a) / (2.0 * g);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 12; ++k) {
          double x = G[k + 12 * p], y = G[k + 12 * q];
          G[k + 12 * p] = c * x - s * y;
          G[k + 12 * q] = s * x + c * y;
        }
        for (int k = 0; k < 6; ++k) {
          double x = Vr[k + 6 * p], y = Vr[k + 6 * q];
          Vr[k + 6 * p] = c * x - s * y;
          Vr[k + 6 * q] = s * x + c * y;
        }
      }
    if (!rot) break;
  }
  double smax = 0.0;
  for (int i = 0; i < 6; ++i) {
    double a = 0;
    for (int k = 0; k < 12; ++k) a += G[k + 12 * i] * G[k + 12 * i];
    sig[i] = sqrt(a);
    if (sig[i] > smax) smax = sig[i];
  }
  /* J = Vr diag(sig) Un^T with Un = G/sig  ->  J^+ = Un diag(1/sig) Vr^T = sum_i G_i/sig_i^2 * Vr_i^T */
  double thr = epsilon * 12.0 * smax;
  if (thr < 2.2250738585072014e-308) thr = 2.2250738585072014e-308;
  for (int i = 0; i < 72; ++i) Jp[i] = 0.0;
  for (int i = 0; i < 6; ++i) {
    if (!(sig[i] > thr)) continue;
    double inv2 = 1.0 / (sig[i] * sig[i]);
    for (int r = 0; r < 12; ++r)
      for (int c = 0; c < 6; ++c) Jp[r + 12 * c] += G[r + 12 * i] * inv2 * Vr[c + 6 * i];
  }
}

/* ------------------------------------------------------------------------------------------
 * geometry with Eigen / Sophus semantics; quaternion = {w,x,y,z}; 3x3 matrices column-major
 * ---------------------------------------------------------------------------------------- */
typedef struct { double w, x, y, z; } quat;
static quat q_from_pose(const double* ps) { quat q = {ps[6], ps[3], ps[4], ps[5]}; return q; }
static quat q_mul(quat a, quat b) {
  quat r = {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
  return r;
}
static quat q_conj(quat q) { quat r = {q.w, -q.x, -q.y, -q.z}; return r; }
static quat q_inv(quat q) {
  double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  quat r = {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return r;
}
static quat q_normalized(quat q) {
  double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  quat r = {q.w / n, q.x / n, q.y / n, q.z / n};
  return r;
}
static void q_rot(quat q, const double* v, double* o) {
  double uv0 = 2.0 * (q.y * v[2] - q.z * v[1]), uv1 = 2.0 * (q.z * v[0] - q.x * v[2]), uv2 = 2.0 * (q.x * v[1] - q.y * v[0]);
  o[0] = v[0] + q.w * uv0 + (q.y * uv2 - q.z * uv1);
  o[1] = v[1] + q.w * uv1 + (q.z * uv0 - q.x * uv2);
  o[2] = v[2] + q.w * uv2 + (q.x * uv1 - q.y * uv0);
}
static void q_to_R(quat q, double* R) { /* column-major */
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z, twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  A_(R, 3, 0, 0) = 1 - (tyy + tzz); A_(R, 3, 0, 1) = txy - twz; A_(R, 3, 0, 2) = txz + twy;
  A_(R, 3, 1, 0) = txy + twz; A_(R, 3, 1, 1) = 1 - (txx + tzz); A_(R, 3, 1, 2) = tyz - twx;
  A_(R, 3, 2, 0) = txz - twy; A_(R, 3, 2, 1) = tyz + twx; A_(R, 3, 2, 2) = 1 - (txx + tyy);
}
static quat R_to_q(const double* m) {
  double t = A_(m, 3, 0, 0) + A_(m, 3, 1, 1) + A_(m, 3, 2, 2);
  double q[4];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (A_(m, 3, 2, 1) - A_(m, 3, 1, 2)) * t;
    q[2] = (A_(m, 3, 0, 2) - A_(m, 3, 2, 0)) * t;
    q[3] = (A_(m, 3, 1, 0) - A_(m, 3, 0, 1)) * t;
  } else {
    int i = 0;
    if (A_(m, 3, 1, 1) > A_(m, 3, 0, 0)) i = 1;
    if (A_(m, 3, 2, 2) > A_(m, 3, i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(A_(m, 3, i, i) - A_(m, 3, j, j) - A_(m, 3, k, k) + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (A_(m, 3, k, j) - A_(m, 3, j, k)) * t;
    q[1 + j] = (A_(m, 3, j, i) + A_(m, 3, i, j)) * t;
    q[1 + k] = (A_(m, 3, k, i) + A_(m, 3, i, k)) * t;
  }
  quat r = {q[0], q[1], q[2], q[3]};
  return r;
}
static void skew(const double* v, double* S) {
  A_(S, 3, 0, 0) = 0; A_(S, 3, 0, 1) = -v[2]; A_(S, 3, 0, 2) = v[1];
  A_(S, 3, 1, 0) = v[2]; A_(S, 3, 1, 1) = 0; A_(S, 3, 1, 2) = -v[0];
  A_(S, 3, 2, 0) = -v[1]; A_(S, 3, 2, 1) = v[0]; A_(S, 3, 2, 2) = 0;
}
static void so3_log(quat q, double* o) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z, f;
  if (n2 < SOPHUS_EPS * SOPHUS_EPS) f = 2.0 / q.w - (2.0 / 3.0) * n2 / (q.w * q.w * q.w);
  else {
    double n = sqrt(n2);
    double at = (q.w < 0) ? atan2(-n, -q.w) : atan2(n, q.w);
    f = 2.0 * at / n;
  }
  o[0] = f * q.x; o[1] = f * q.y; o[2] = f * q.z;
}
static quat so3_mul(quat a, quat b) {
  quat q = q_mul(a, b);
  double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  if (n2 != 1.0) { double s = 2.0 / (1.0 + n2); q.w *= s; q.x *= s; q.y *= s; q.z *= s; }
  return q;
}
static void right_jacobian_inv(const double* phi, double* J) {
  double n2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], H[9], H2[9], coef;
  skew(phi, H);
  mm(3, 3, 3, H, 3, 0, H, 3, 0, H2, 3, 0);
  if (n2 > SOPHUS_EPS) {
    double n = sqrt(n2);
    if (n < M_PI - SOPHUS_EPS_SQRT) coef = 1.0 / n2 - (1.0 + cos(n)) / (2.0 * n * sin(n));
    else coef = 1.0 / (M_PI * M_PI);
  } else coef = 1.0 / 12.0;
  for (int i = 0; i < 9; ++i) J[i] = 0.5 * H[i] + coef * H2[i];
  J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

/* ------------------------------------------------------------------------------------------
 * factors (tangent-space twins)
 * ---------------------------------------------------------------------------------------- */
/* RelativePoseFactor::EvaluateOnlyJacobians (relative_pose_factor.h:72-101): Ji, Jj 6x6 */
static void relpose_jac(const double* PSi, const double* PSj, const double* dR, double* Ji, double* Jj) {
  quat Qi = q_from_pose(PSi), Qj = q_from_pose(PSj);
  double Ri[9], Rj[9], d[3] = {PSj[0] - PSi[0], PSj[1] - PSi[1], PSj[2] - PSi[2]}, tij[3];
  q_to_R(Qi, Ri);
  q_to_R(Qj, Rj);
  q_rot(q_inv(Qi), d, tij);
  double A[9], B[9], lg[3], Jr[9], S[9], RitRj[9], JR[9];
  mm(3, 3, 3, dR, 3, 0, Rj, 3, 1, A, 3, 0);
  mm(3, 3, 3, A, 3, 0, Ri, 3, 0, B, 3, 0);
  so3_log(R_to_q(B), lg);
  right_jacobian_inv(lg, Jr);
  skew(tij, S);
  mm(3, 3, 3, Ri, 3, 1, Rj, 3, 0, RitRj, 3, 0);
  mm(3, 3, 3, Jr, 3, 0, RitRj, 3, 0, JR, 3, 0);
  memset(Ji, 0, sizeof(double) * 36);
  memset(Jj, 0, sizeof(double) * 36);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      A_(Ji, 6, r, c) = A_(Ri, 3, c, r);
      A_(Ji, 6, r, 3 + c) = -A_(S, 3, r, c);
      A_(Ji, 6, 3 + r, 3 + c) = A_(Jr, 3, r, c);
      A_(Jj, 6, r, c) = -A_(Ri, 3, c, r);
      A_(Jj, 6, 3 + r, 3 + c) = -A_(JR, 3, r, c);
    }
}
/* SE3PriorFactor::EvaluateOnlyJacobians (se3_prior_factor.h:53-71) */
static void se3prior_jac(const double* PS, const double* Rprior, double* J) {
  quat ri = q_normalized(q_from_pose(PS)), rp = R_to_q(Rprior);
  double lg[3], Jr[9];
  so3_log(so3_mul(q_conj(rp), ri), lg);
  right_jacobian_inv(lg, Jr);
  memset(J, 0, sizeof(double) * 36);
  A_(J, 6, 0, 0) = 1; A_(J, 6, 1, 1) = 1; A_(J, 6, 2, 2) = 1;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) A_(J, 6, 3 + r, 3 + c) = A_(Jr, 3, r, c);
}
/* RollPitchFactor::EvaluateOnlyJacobians (rollpitch_factor.h:59-76): J 2x6 */
static void rollpitch_jac(const double* PS, const double* Rmeas, double* J) {
  quat ri = q_normalized(q_from_pose(PS)), rm = R_to_q(Rmeas);
  double nz[3] = {0, 0, -1}, a[3], r3[3], S[9], Rm[9], SR[9];
  q_rot(q_conj(ri), nz, a);
  q_rot(rm, a, r3);
  skew(r3, S);
  q_to_R(rm, Rm);
  mm(3, 3, 3, S, 3, 0, Rm, 3, 0, SR, 3, 0);
  memset(J, 0, sizeof(double) * 12);
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c) A_(J, 2, r, 3 + c) = A_(SR, 3, r, c);
}
/* YawFactor::EvaluateOnlyJacobians (yaw_factor.h:51-65): J 1x6 */
static void yaw_jac(const double* PS, double* J) {
  quat q = q_from_pose(PS), ri = q_normalized(q);
  double ex[3] = {1, 0, 0}, ym[3], R[9], S[9], RS[9];
  q_rot(q_inv(q), ex, ym);
  q_to_R(ri, R);
  skew(ym, S);
  mm(3, 3, 3, R, 3, 0, S, 3, 0, RS, 3, 0);
  memset(J, 0, sizeof(double) * 6);
  for (int c = 0; c < 3; ++c) J[3 + c] = -A_(RS, 3, 1, c);
}
/* ProjectionFactor::EvaluateOnlyJacobians (projection_factor.cpp:124-196): Ji, Jj 2x6, Jf 2x1 */
static void projection_jac(const double* PSi, const double* PSj, const double* PSic, double inv_dep, const double* pts_i,
                           double* Ji, double* Jj, double* Jf) {
  quat Qi = q_from_pose(PSi), Qj = q_from_pose(PSj), qic = q_from_pose(PSic);
  double pc[3] = {pts_i[0] / inv_dep, pts_i[1] / inv_dep, pts_i[2] / inv_dep}, pim[3], pw[3], d[3], pj[3], cj[3];
  q_rot(qic, pc, pim);
  for (int i = 0; i < 3; ++i) pim[i] += PSic[i];
  q_rot(Qi, pim, pw);
  for (int i = 0; i < 3; ++i) d[i] = pw[i] + PSi[i] - PSj[i];
  q_rot(q_inv(Qj), d, pj);
  for (int i = 0; i < 3; ++i) d[i] = pj[i] - PSic[i];
  q_rot(q_inv(qic), d, cj);
  double dep = cj[2];
  double Ri[9], Rj[9], ric[9], red[6] = {1.0 / dep, 0, 0, 1.0 / dep, -cj[0] / (dep * dep), -cj[1] / (dep * dep)};
  q_to_R(Qi, Ri);
  q_to_R(Qj, Rj);
  q_to_R(qic, ric);
  double B[9], C[9], S[9], T[9], jaco[18];
  mm(3, 3, 3, ric, 3, 1, Rj, 3, 1, B, 3, 0);  /* ric^T Rj^T */
  mm(3, 3, 3, B, 3, 0, Ri, 3, 0, C, 3, 0);    /* ric^T Rj^T Ri */
  skew(pim, S);
  mm(3, 3, 3, C, 3, 0, S, 3, 0, T, 3, 0);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { A_(jaco, 3, r, c) = A_(B, 3, r, c); A_(jaco, 3, r, 3 + c) = -A_(T, 3, r, c); }
  mm(2, 6, 3, red, 2, 0, jaco, 3, 0, Ji, 2, 0);
  skew(pj, S);
  mm(3, 3, 3, ric, 3, 1, S, 3, 0, T, 3, 0);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { A_(jaco, 3, r, c) = -A_(B, 3, r, c); A_(jaco, 3, r, 3 + c) = A_(T, 3, r, c); }
  mm(2, 6, 3, red, 2, 0, jaco, 3, 0, Jj, 2, 0);
  double Ap[9], f[3];
  mm(3, 3, 3, C, 3, 0, ric, 3, 0, Ap, 3, 0);
  for (int r = 0; r < 3; ++r)
    f[r] = (A_(Ap, 3, r, 0) * pts_i[0] + A_(Ap, 3, r, 1) * pts_i[1] + A_(Ap, 3, r, 2) * pts_i[2]) * -1.0 / (inv_dep * inv_dep);
  mm(2, 1, 3, red, 2, 0, f, 3, 0, Jf, 2, 0);
}

/* IMUFactor::Evaluate tangent twin (imu_factor.h:161-265): J0 15x6, J1 15x9, J2 15x6, J3 15x9
 * (unweighted), info = covariance^-1 (= sqrt_info^T sqrt_info with sqrt_info = LLT(cov^-1).L^T). */
static void qleft(quat q, double* M) { /* 4x4 col-major, Utility::Qleft */
  double v[3] = {q.x, q.y, q.z}, S[9];
  skew(v, S);
  memset(M, 0, sizeof(double) * 16);
  A_(M, 4, 0, 0) = q.w;
  for (int i = 0; i < 3; ++i) { A_(M, 4, 0, 1 + i) = -v[i]; A_(M, 4, 1 + i, 0) = v[i]; }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) A_(M, 4, 1 + r, 1 + c) = (r == c ? q.w : 0.0) + A_(S, 3, r, c);
}
static void qright(quat q, double* M) {
  double v[3] = {q.x, q.y, q.z}, S[9];
  skew(v, S);
  memset(M, 0, sizeof(double) * 16);
  A_(M, 4, 0, 0) = q.w;
  for (int i = 0; i < 3; ++i) { A_(M, 4, 0, 1 + i) = -v[i]; A_(M, 4, 1 + i, 0) = v[i]; }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) A_(M, 4, 1 + r, 1 + c) = (r == c ? q.w : 0.0) - A_(S, 3, r, c);
}
static void imu_eval(const double* PSi, const double* VBi, const double* PSj, const double* VBj, const double* pre,
                     const double* G, double* J0, double* J1, double* J2, double* J3, double* sqrt_info) {
  quat dq = {pre[6], pre[3], pre[4], pre[5]};
  const double *lin_bg = pre + 13, *Jp = pre + 17, *cov = pre + 17 + 225;
  double s = pre[16];
  quat Qi = q_from_pose(PSi), Qj = q_from_pose(PSj), Qi_inv = q_inv(Qi);
  double Ri_inv[9];
  q_to_R(Qi_inv, Ri_inv);
  double dbg[3] = {VBi[6] - lin_bg[0], VBi[7] - lin_bg[1], VBi[8] - lin_bg[2]}, th[3];
  for (int r = 0; r < 3; ++r) th[r] = A_(Jp, 15, 3 + r, 12) * dbg[0] + A_(Jp, 15, 3 + r, 13) * dbg[1] + A_(Jp, 15, 3 + r, 14) * dbg[2];
  quat dth = {1.0, th[0] / 2, th[1] / 2, th[2] / 2};
  quat cq = q_mul(dq, dth);
  double a1[3], a2[3], v1[3], v2[3], S1[9], S2[9];
  for (int k = 0; k < 3; ++k) {
    a1[k] = 0.5 * G[k] * s * s + PSj[k] - PSi[k] - VBi[k] * s;
    a2[k] = G[k] * s + VBj[k] - VBi[k];
  }
  q_rot(Qi_inv, a1, v1);
  q_rot(Qi_inv, a2, v2);
  skew(v1, S1);
  skew(v2, S2);
  double QL[16], QR[16], QLR[16], QL1[16], QL2[16];
  quat QjinvQi = q_mul(q_inv(Qj), Qi);
  qleft(QjinvQi, QL);
  qright(cq, QR);
  mm(4, 4, 4, QL, 4, 0, QR, 4, 0, QLR, 4, 0);
  qleft(q_mul(QjinvQi, dq), QL1);
  qleft(q_mul(q_mul(q_inv(cq), Qi_inv), Qj), QL2);
  memset(J0, 0, sizeof(double) * 90);
  memset(J1, 0, sizeof(double) * 135);
  memset(J2, 0, sizeof(double) * 90);
  memset(J3, 0, sizeof(double) * 135);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      A_(J0, 15, r, c) = -A_(Ri_inv, 3, r, c);
      A_(J0, 15, r, 3 + c) = A_(S1, 3, r, c);
      A_(J0, 15, 3 + r, 3 + c) = -A_(QLR, 4, 1 + r, 1 + c);
      A_(J0, 15, 6 + r, 3 + c) = A_(S2, 3, r, c);
      A_(J1, 15, r, c) = -A_(Ri_inv, 3, r, c) * s;
      A_(J1, 15, r, 3 + c) = -A_(Jp, 15, r, 9 + c);
      A_(J1, 15, r, 6 + c) = -A_(Jp, 15, r, 12 + c);
      double qd = 0;
      for (int k = 0; k < 3; ++k) qd += A_(QL1, 4, 1 + r, 1 + k) * A_(Jp, 15, 3 + k, 12 + c);
      A_(J1, 15, 3 + r, 6 + c) = -qd;
      A_(J1, 15, 6 + r, c) = -A_(Ri_inv, 3, r, c);
      A_(J1, 15, 6 + r, 3 + c) = -A_(Jp, 15, 6 + r, 9 + c);
      A_(J1, 15, 6 + r, 6 + c) = -A_(Jp, 15, 6 + r, 12 + c);
      A_(J1, 15, 9 + r, 3 + c) = r == c ? -1.0 : 0.0;
      A_(J1, 15, 12 + r, 6 + c) = r == c ? -1.0 : 0.0;
      A_(J2, 15, r, c) = A_(Ri_inv, 3, r, c);
      A_(J2, 15, 3 + r, 3 + c) = A_(QL2, 4, 1 + r, 1 + c);
      A_(J3, 15, 6 + r, c) = A_(Ri_inv, 3, r, c);
      A_(J3, 15, 9 + r, 3 + c) = r == c ? 1.0 : 0.0;
      A_(J3, 15, 12 + r, 6 + c) = r == c ? 1.0 : 0.0;
    }
  double inv[225];
  memcpy(inv, cov, sizeof(inv));
  partial_piv_inverse(15, inv);
  llt_upper(15, inv, sqrt_info);
}

/* the block loop (estimator.cpp:1183-1201 etc.): Lamda += J_j^T info J_k (+ transpose) */
static void accumulate(double* Lam, int n, int nb, const int* off, const int* dim, const double* const* jac, int rows,
                       const double* info) {
  double JtW[9 * 15], H[81];
  for (int j = 0; j < nb; ++j) {
    if (off[j] < 0) continue;
    mm(dim[j], rows, rows, jac[j], rows, 1, info, rows, 0, JtW, dim[j], 0);
    for (int k = j; k < nb; ++k) {
      if (off[k] < 0) continue;
      mm(dim[j], dim[k], rows, JtW, dim[j], 0, jac[k], rows, 0, H, dim[j], 0);
      for (int c = 0; c < dim[k]; ++c)
        for (int r = 0; r < dim[j]; ++r) {
          A_(Lam, n, off[j] + r, off[k] + c) += A_(H, dim[j], r, c);
          if (j != k) A_(Lam, n, off[k] + c, off[j] + r) += A_(H, dim[j], r, c);
        }
    }
  }
}

/* cov = (J U) Dinv (J U)^T ; Omega = cov^-1 ; sqrt_info = LLT(Omega).L^T   (m <= 9, n x rank) */
static int recover(int m, int n, const double* J, int ldj, const double* U, const double* dinv, int nk, double* sqrt_info) {
  double JU[9 * 64], cov[81];
  mm(m, nk, n, J, ldj, 0, U, n, 0, JU, m, 0);
  for (int c = 0; c < m; ++c)
    for (int r = 0; r < m; ++r) {
      double acc = 0;
      for (int k = 0; k < nk; ++k) acc += A_(JU, m, r, k) * dinv[k] * A_(JU, m, c, k);
      A_(cov, m, r, c) = acc;
    }
  partial_piv_inverse(m, cov);
  return llt_upper(m, cov, sqrt_info);
}

/* ------------------------------------------------------------------------------------------
 * MargForward (estimator.cpp:1149-1352) for one window
 * ---------------------------------------------------------------------------------------- */
static int marg_forward_one(const isv_config* cfg, int L, const double* obs, int64_t st, const double* pose0,
                            const double* pose1, const double* ex, const double* pse3, const double* prel,
                            const double* prp, double* o_se3, double* o_pg, int32_t* o_rank, int structured) {
  int status = 0;
  const int n = L + 12;
  double* Lam = (double*)calloc((size_t)n * n, sizeof(double));
  double info_p[4];
  mm(2, 2, 2, cfg->proj_sqrt_info, 2, 1, cfg->proj_sqrt_info, 2, 0, info_p, 2, 0);
  for (int k = 0; k < L; ++k) { /* :1168-1202 */
    double pts_i[3] = {obs[k], obs[st + k], obs[2 * st + k]};
    double Ji[12], Jj[12], Jf[2];
    projection_jac(pose0, pose1, ex, obs[5 * st + k], pts_i, Ji, Jj, Jf);
    int off[4] = {6, 0, -1, 12 + k}, dim[4] = {6, 6, 6, 1};
    const double* jac[4] = {Ji, Jj, NULL, Jf};
    accumulate(Lam, n, 4, off, dim, jac, 2, info_p);
  }
  { /* vioPosePriorEdge :1203-1211 */
    double Jp[36], info[36], T[36];
    se3prior_jac(pose0, pse3 + 3, Jp);
    mm(6, 6, 6, pse3 + 12, 6, 1, pse3 + 12, 6, 0, info, 6, 0);
    mm(6, 6, 6, Jp, 6, 1, info, 6, 0, T, 6, 0);
    mm(6, 6, 6, T, 6, 0, Jp, 6, 0, Lam + 6 + (size_t)n * 6, n, 1);
  }
  { /* vioRelativePoseEdges[1] :1212-1238 */
    double Ji[36], Jj[36], info[36];
    relpose_jac(pose0, pose1, prel + 3, Ji, Jj);
    mm(6, 6, 6, prel + 12, 6, 1, prel + 12, 6, 0, info, 6, 0);
    int off[2] = {6, 0}, dim[2] = {6, 6};
    const double* jac[2] = {Ji, Jj};
    accumulate(Lam, n, 2, off, dim, jac, 6, info);
  }
  /* pose-graph factor :1243-1259 */
  quat Qi = q_from_pose(pose0), Qj = q_from_pose(pose1);
  double d[3] = {pose1[0] - pose0[0], pose1[1] - pose0[1], pose1[2] - pose0[2]}, tij[3], Rij[9];
  q_rot(q_inv(Qi), d, tij);
  q_to_R(q_mul(q_inv(Qi), Qj), Rij);
  double J[72], Jp[72], Lrp[144], T[72], rpOmega[36], rpCov[36];
  relpose_jac(pose0, pose1, Rij, J, J + 36);
  pseudo_inverse_6x12(J, 1e-8, Jp);
  for (int c = 0; c < 12; ++c)
    for (int r = 0; r < 12; ++r) Lrp[r + 12 * c] = A_(Lam, n, r, c);
  mm(12, 6, 12, Lrp, 12, 0, Jp, 12, 0, T, 12, 0);
  mm(6, 6, 12, Jp, 12, 1, T, 12, 0, rpOmega, 6, 0);
  memcpy(rpCov, rpOmega, sizeof(rpCov));
  partial_piv_inverse(6, rpCov);
  memcpy(o_pg, tij, sizeof(double) * 3);
  memcpy(o_pg + 3, Rij, sizeof(double) * 9);
  if (llt_upper(6, rpOmega, o_pg + 12)) status |= ISV_W_NOT_SPD;
  memcpy(o_pg + 48, rpCov, sizeof(rpCov));
  o_pg[84] = sqrt(tij[0] * tij[0] + tij[1] * tij[1] + tij[2] * tij[2]);
  for (int i = 0; i < 4; ++i) o_pg[85 + i] = 0.0;
  if (prp && prp[0] != 0.0) {
    double m[4];
    mm(2, 2, 2, prp + 1, 2, 1, prp + 1, 2, 0, m, 2, 0);
    partial_piv_inverse(2, m);
    memcpy(o_pg + 85, m, sizeof(m));
  }
  /* Schur complement :1286-1288 */
  double Lprior[36];
  if (!structured) {
    const int m = L + 6;
    double* Lmm = (double*)malloc(sizeof(double) * (size_t)m * m * 2);
    double* Linv = Lmm + (size_t)m * m;
    for (int c = 0; c < m; ++c) memcpy(Lmm + (size_t)m * c, Lam + 6 + (size_t)n * (6 + c), sizeof(double) * (size_t)m);
    fullpiv_lu_solve_identity(m, Lmm, Linv);
    /* Lamda_prior = Lamda_rr - Lamda_rm * Linv * Lamda_rm^T */
    double* T2 = (double*)malloc(sizeof(double) * 6 * (size_t)m);
    for (int c = 0; c < m; ++c)
      for (int r = 0; r < 6; ++r) {
        double acc = 0;
        for (int l = 0; l < m; ++l) acc += A_(Lam, n, r, 6 + l) * Linv[l + (size_t)m * c];
        T2[r + 6 * c] = acc;
      }
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) {
        double acc = 0;
        for (int l = 0; l < m; ++l) acc += T2[r + 6 * l] * A_(Lam, n, c, 6 + l);
        Lprior[r + 6 * c] = A_(Lam, n, r, c) - acc;
      }
    free(T2);
    free(Lmm);
  } else {
    double S[144], Smm[36], T3[36];
    for (int c = 0; c < 12; ++c)
      for (int r = 0; r < 12; ++r) {
        double acc = A_(Lam, n, r, c);
        for (int k = 0; k < L; ++k) acc -= A_(Lam, n, r, 12 + k) * A_(Lam, n, c, 12 + k) / A_(Lam, n, 12 + k, 12 + k);
        S[r + 12 * c] = acc;
      }
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) Smm[r + 6 * c] = S[6 + r + 12 * (6 + c)];
    partial_piv_inverse(6, Smm);
    mm(6, 6, 6, S + 12 * 6, 12, 0, Smm, 6, 0, T3, 6, 0);
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) Lprior[r + 6 * c] = S[r + 12 * c];
    mm(6, 6, 6, T3, 6, 0, S + 12 * 6, 12, 1, Lprior, 6, -1);
  }
  /* recovered SE3 prior on T1 :1291-1349 */
  double R1[9], Jr[36], cov[36], covi[36], T4[36];
  q_to_R(Qj, R1);
  se3prior_jac(pose1, R1, Jr);
  int rank = fullpiv_qr_inverse(6, Lprior, pow(10.0, (double)cfg->qr_rank_eps_log10), cov);
  *o_rank = rank;
  if (rank == 6) {
    mm(6, 6, 6, Jr, 6, 0, cov, 6, 0, T4, 6, 0);
    mm(6, 6, 6, T4, 6, 0, Jr, 6, 1, covi, 6, 0);
  } else {
    status |= ISV_W_RANK_DEFICIENT;
    double w[6], Z[36], U[36], dinv[6], JU[36];
    sym_eig(6, Lprior, w, Z);
    int nk = 0;
    for (int i = 0; i < 6; ++i)
      if (w[i] > cfg->alpha) { memcpy(U + 6 * nk, Z + 6 * i, sizeof(double) * 6); dinv[nk] = 1.0 / w[i]; ++nk; }
    *o_rank = nk;
    mm(6, nk, 6, Jr, 6, 0, U, 6, 0, JU, 6, 0);
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) {
        double acc = 0;
        for (int k = 0; k < nk; ++k) acc += JU[r + 6 * k] * dinv[k] * JU[c + 6 * k];
        covi[r + 6 * c] = acc;
      }
  }
  partial_piv_inverse(6, covi);
  memcpy(o_se3, pose1, sizeof(double) * 3);
  memcpy(o_se3 + 3, R1, sizeof(double) * 9);
  if (llt_upper(6, covi, o_se3 + 12)) status |= ISV_W_NOT_SPD;
  free(Lam);
  return status;
}

/* ------------------------------------------------------------------------------------------
 * MargBackward (estimator.cpp:1354-1539) for one window
 * ---------------------------------------------------------------------------------------- */
static int marg_backward_one(const isv_config* cfg, const double* pose_i, const double* sb_i, const double* pose_j,
                             const double* sb_j, const double* pvb, const double* pre, double* o_rel, double* o_vb,
                             double* o_rp, int32_t* o_rank) {
  int status = 0;
  double Lam[900];
  memset(Lam, 0, sizeof(Lam));
  mm(9, 9, 9, pvb + 9, 9, 1, pvb + 9, 9, 0, Lam + 21 + 30 * 21, 30, 1); /* :1372-1380 */
  double J0[90], J1[135], J2[90], J3[135], sI[225], omegaI[225];
  imu_eval(pose_i, sb_i, pose_j, sb_j, pre, cfg->g, J0, J1, J2, J3, sI);
  mm(15, 15, 15, sI, 15, 1, sI, 15, 0, omegaI, 15, 0);
  {
    int off[4] = {15, 21, 0, 6}, dim[4] = {6, 9, 6, 9};
    const double* jac[4] = {J0, J1, J2, J3};
    accumulate(Lam, 30, 4, off, dim, jac, 15, omegaI);
  }
  double Lmm[81], Linv[81], T[189], Lprior[441];
  for (int c = 0; c < 9; ++c)
    for (int r = 0; r < 9; ++r) Lmm[r + 9 * c] = Lam[21 + r + 30 * (21 + c)];
  fullpiv_lu_solve_identity(9, Lmm, Linv);
  mm(21, 9, 9, Lam + 30 * 21, 30, 0, Linv, 9, 0, T, 21, 0);
  for (int c = 0; c < 21; ++c)
    for (int r = 0; r < 21; ++r) Lprior[r + 21 * c] = Lam[r + 30 * c];
  mm(21, 21, 9, T, 21, 0, Lam + 30 * 21, 30, 1, Lprior, 21, -1);
  /* recovered factors and Jr :1424-1477 */
  quat Qi = q_from_pose(pose_i), Qj = q_from_pose(pose_j);
  double d[3] = {pose_j[0] - pose_i[0], pose_j[1] - pose_i[1], pose_j[2] - pose_i[2]}, tij[3], Rij[9], Rw[9];
  q_rot(q_inv(Qi), d, tij);
  q_to_R(q_mul(q_inv(Qi), Qj), Rij);
  q_to_R(Qi, Rw);
  double Ji[36], Jj[36], Jrp[12], Jy[6], Jr[441];
  relpose_jac(pose_i, pose_j, Rij, Ji, Jj);
  rollpitch_jac(pose_i, Rw, Jrp);
  yaw_jac(pose_i, Jy);
  memset(Jr, 0, sizeof(Jr));
  for (int c = 0; c < 6; ++c) {
    for (int r = 0; r < 6; ++r) { Jr[r + 21 * (15 + c)] += Ji[r + 6 * c]; Jr[r + 21 * c] += Jj[r + 6 * c]; }
    for (int r = 0; r < 2; ++r) Jr[15 + r + 21 * (15 + c)] += Jrp[r + 2 * c];
    Jr[20 + 21 * (15 + c)] += Jy[c];
  }
  for (int i = 0; i < 9; ++i) Jr[6 + i + 21 * (6 + i)] += 1.0;
  for (int i = 0; i < 3; ++i) Jr[17 + i + 21 * (15 + i)] += 1.0;
  /* truncated eigen :1479-1497 */
  double w[21], Z[441], U[441], dinv[21];
  sym_eig(21, Lprior, w, Z);
  int nk = 0;
  for (int i = 0; i < 21; ++i)
    if (w[i] > cfg->alpha) { memcpy(U + 21 * nk, Z + 21 * i, sizeof(double) * 21); dinv[nk] = 1.0 / w[i]; ++nk; }
  *o_rank = nk;
  memcpy(o_rel, tij, sizeof(double) * 3);
  memcpy(o_rel + 3, Rij, sizeof(double) * 9);
  if (recover(6, 21, Jr, 21, U, dinv, nk, o_rel + 12)) status |= ISV_W_NOT_SPD;
  memcpy(o_vb, sb_j, sizeof(double) * 9);
  if (recover(9, 21, Jr + 6, 21, U, dinv, nk, o_vb + 9)) status |= ISV_W_NOT_SPD;
  memcpy(o_rp, Rw, sizeof(double) * 9);
  if (recover(2, 21, Jr + 15, 21, U, dinv, nk, o_rp + 9)) status |= ISV_W_NOT_SPD;
  double tmp[9];
  recover(3, 21, Jr + 17, 21, U, dinv, nk, tmp); /* abs position / yaw: computed and discarded (:1518-1519) */
  recover(1, 21, Jr + 20, 21, U, dinv, nk, tmp);
  return status;
}

/* ------------------------------------------------------------------------------------------
 * batch entry point: HOST pointers in the product's SoA layout
 * ---------------------------------------------------------------------------------------- */
int isv_ref_marg_window_batch(const isv_config* cfg, const isv_batch_in* in, const isv_batch_out* out, int which,
                              int n_threads, int structured) {
  const int n = in->n_windows;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int w = 0; w < n; ++w) {
    int status = 0;
    if (which & ISV_RUN_FORWARD) {
      int64_t a = in->lm_offset[w];
      int L = (int)(in->lm_offset[w + 1] - a);
      const double* ex = in->ex_pose + (in->ex_pose_shared ? 0 : (size_t)w * 7);
      status |= marg_forward_one(cfg, L, in->lm_obs + a, in->lm_stride, in->pose_fwd + (size_t)w * 14,
                                 in->pose_fwd + (size_t)w * 14 + 7, ex, in->prior_se3 + (size_t)w * ISV_SE3_REC,
                                 in->prior_rel + (size_t)w * ISV_REL_REC,
                                 in->prior_rp ? in->prior_rp + (size_t)w * ISV_RP_IN_REC : NULL,
                                 out->se3_out + (size_t)w * ISV_SE3_REC, out->pg_out + (size_t)w * ISV_PG_REC,
                                 out->rank + 2 * (size_t)w, structured);
    }
    if (which & ISV_RUN_BACKWARD) {
      status |= marg_backward_one(cfg, in->pose_bwd + (size_t)w * 14, in->sb_bwd + (size_t)w * 18,
                                  in->pose_bwd + (size_t)w * 14 + 7, in->sb_bwd + (size_t)w * 18 + 9,
                                  in->prior_vb + (size_t)w * ISV_VB_REC, in->preint + (size_t)w * ISV_PREINT_REC,
                                  out->rel_out + (size_t)w * ISV_REL_REC, out->vb_out + (size_t)w * ISV_VB_REC,
                                  out->rp_out + (size_t)w * ISV_RP_REC, out->rank + 2 * (size_t)w + 1);
    }
    if (out->status) out->status[w] = status;
  }
  return 0;
}

int isv_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
