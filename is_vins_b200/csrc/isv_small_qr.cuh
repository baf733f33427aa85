// Single-lane (serial) Eigen::FullPivHouseholderQR restated for tiny matrices in shared memory.
// Used only for the 6x6 rank decision + solve of MargForward (/root/reference/src/estimator.cpp:
// 1304-1309): `qr(Lamda_prior); qr.setThreshold(eps); if (qr.rank()==6) cov = qr.solve(I)`.
// The discrete outcome (rank) must match the reference exactly, so the pivoting rule, the
// "biggest-in-corner <= |biggest| * eps * size" early stop and the rank count
// |R_kk| > threshold * maxpivot follow Eigen 3.3's FullPivHouseholderQR::computeInPlace.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace isv {

// A: n x n column-major (ld n), overwritten by the QR factors.  X: n x n receives solve(I) when
// rank == n (untouched otherwise).  hc: n doubles scratch.  Returns the rank.
template <int N>
__device__ inline int serial_fullpiv_qr_inverse(double* A, double* X, double threshold) {
  double hc[N];
  int rowt[N], colp[N];
  const double eps = 2.220446049250313e-16;
  const double precision = eps * N;
  double biggest = 0.0, maxpivot = 0.0;
  int nonzero = N;
  for (int k = 0; k < N; ++k) colp[k] = k;
  for (int k = 0; k < N; ++k) {
    // largest |entry| in the bottom-right corner; Eigen scans column-major and keeps the first max
    int pr = k, pc = k;
    double big = -1.0;
    for (int c = k; c < N; ++c)
      for (int r = k; r < N; ++r) {
        double v = fabs(A[r + N * c]);
        if (v > big) { big = v; pr = r; pc = c; }
      }
    if (k == 0) biggest = big;
    if (big <= biggest * precision || big == 0.0) {  // isMuchSmallerThan(biggest, precision)
      nonzero = k;
      for (int i = k; i < N; ++i) { rowt[i] = i; hc[i] = 0.0; }
      break;
    }
    rowt[k] = pr;
    if (pr != k)
      for (int c = k; c < N; ++c) { double t = A[k + N * c]; A[k + N * c] = A[pr + N * c]; A[pr + N * c] = t; }
    if (pc != k) {
      for (int r = 0; r < N; ++r) { double t = A[r + N * k]; A[r + N * k] = A[r + N * pc]; A[r + N * pc] = t; }
      int t = colp[k]; colp[k] = colp[pc]; colp[pc] = t;
    }
    // Householder reflector of A[k:, k]  (Eigen makeHouseholderInPlace)
    double c0 = A[k + N * k];
    double tail2 = 0.0;
    for (int r = k + 1; r < N; ++r) tail2 = fma(A[r + N * k], A[r + N * k], tail2);
    double beta, tau;
    if (tail2 <= 2.2250738585072014e-308) {
      tau = 0.0;
      beta = c0;
      for (int r = k + 1; r < N; ++r) A[r + N * k] = 0.0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0.0) beta = -beta;
      double inv = 1.0 / (c0 - beta);
      for (int r = k + 1; r < N; ++r) A[r + N * k] *= inv;
      tau = (beta - c0) / beta;
    }
    A[k + N * k] = beta;
    hc[k] = tau;
    if (fabs(beta) > maxpivot) maxpivot = fabs(beta);
    if (tau != 0.0)
      for (int c = k + 1; c < N; ++c) {
        double s = A[k + N * c];
        for (int r = k + 1; r < N; ++r) s = fma(A[r + N * k], A[r + N * c], s);
        s *= tau;
        A[k + N * c] -= s;
        for (int r = k + 1; r < N; ++r) A[r + N * c] = fma(-s, A[r + N * k], A[r + N * c]);
      }
  }
  int rank = 0;
  for (int k = 0; k < nonzero; ++k)
    if (fabs(A[k + N * k]) > maxpivot * threshold) ++rank;
  if (rank < N) return rank;
  // solve A X = I : c = Q^T I (apply reflectors with the row transpositions), back-substitute, permute
  for (int j = 0; j < N; ++j) {
    double c[N];
    for (int r = 0; r < N; ++r) c[r] = (r == j) ? 1.0 : 0.0;
    for (int k = 0; k < N; ++k) {
      int r = rowt[k];
      if (r != k) { double t = c[k]; c[k] = c[r]; c[r] = t; }
      double s = c[k];
      for (int i = k + 1; i < N; ++i) s = fma(A[i + N * k], c[i], s);
      s *= hc[k];
      c[k] -= s;
      for (int i = k + 1; i < N; ++i) c[i] = fma(-s, A[i + N * k], c[i]);
    }
    for (int i = N - 1; i >= 0; --i) {
      double s = c[i];
      for (int l = i + 1; l < N; ++l) s = fma(-A[i + N * l], c[l], s);
      c[i] = s / A[i + N * i];
    }
    for (int i = 0; i < N; ++i) X[colp[i] + N * j] = c[i];
  }
  return rank;
}

}  // namespace isv
