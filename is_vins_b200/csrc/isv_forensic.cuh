// Forensic / debug kernels.  NOT on the hot path: they exist so that a parity break can be localised on the GPU itself,
// without any CPU code (SURVEY.md 7.3 item 5 "both must be available"), and so that the reference's KLD diagnostics
// (computed and discarded by the reference, Q7) can be inspected.
//
//   literal_schur_kernel   the reference's dense route (/root/reference/src/estimator.cpp:1286-1288, :1417-1419):
//                          Lamda_mm^-1 of the WHOLE marginalized block by a full-pivot elimination (the algorithm class
//                          of Eigen's FullPivLU + solve(Identity)), then Lamda_rr - Lamda_rm Lamda_mm^-1 Lamda_rm^T.
//                          One CTA per problem, matrices in global memory (L2 resident), O(m^3): fine for forensics.
//   marg_kld_kernel        the KLD of the recovered factor set against the dense marginal: forward (:1333-1345) and
//                          backward (:1522-1534, with the abs-position / yaw informations of :1518-1519 that only feed
//                          the KLD), one THREAD per window, everything in local memory.
#pragma once
#include "isv_window_kernels.cuh"

namespace isv {

constexpr int kLitThreads = 256;

// A: [np][n*n] column-major symmetric.  kept block = rows/cols [0, m0), marginalized block = [m0, n).
// work: [np][m * 2m] row-major augmented [W | B] scratch.  prior_out: [np][m0*m0] column-major.
// inv_out (may be null): [np][m*m] column-major.  rank_out: [np] pivots accepted (Eigen: |pivot| > eps * m * maxpivot).
__global__ void __launch_bounds__(kLitThreads)
literal_schur_kernel(int n, int m0, const double* __restrict__ A_all, double* __restrict__ work_all,
                     double* __restrict__ prior_out, double* __restrict__ inv_out, int32_t* __restrict__ rank_out) {
  const int m = n - m0, tid = threadIdx.x, ld = 2 * m;
  const double* A = A_all + (size_t)blockIdx.x * n * n;
  double* W = work_all + (size_t)blockIdx.x * m * ld;
  __shared__ double red_v[kLitThreads];
  __shared__ int red_i[kLitThreads];
  __shared__ int s_piv, s_rank;
  __shared__ double s_maxpivot;
  extern __shared__ int colperm[];      // m ints
  for (int idx = tid; idx < m * ld; idx += kLitThreads) {
    const int i = idx / ld, j = idx - i * ld;
    W[idx] = j < m ? A[(m0 + i) + (size_t)n * (m0 + j)] : ((j - m) == i ? 1.0 : 0.0);
  }
  for (int j = tid; j < m; j += kLitThreads) colperm[j] = j;
  if (tid == 0) { s_rank = m; s_maxpivot = 0.0; }
  __syncthreads();
  for (int k = 0; k < m; ++k) {
    // full pivoting: the largest |entry| of the remaining corner (ties: lowest flat index, like a sequential scan)
    double bv = -1.0;
    int bi = 0;
    const int side = m - k;
    for (int idx = tid; idx < side * side; idx += kLitThreads) {
      const int i = k + idx / side, j = k + idx % side;
      const double v = fabs(W[i * ld + j]);
      if (v > bv) { bv = v; bi = i * m + j; }
    }
    red_v[tid] = bv;
    red_i[tid] = bi;
    __syncthreads();
    for (int s = kLitThreads / 2; s > 0; s >>= 1) {
      if (tid < s) {
        const double ov = red_v[tid + s];
        const int oi = red_i[tid + s];
        if (ov > red_v[tid] || (ov == red_v[tid] && oi < red_i[tid])) { red_v[tid] = ov; red_i[tid] = oi; }
      }
      __syncthreads();
    }
    if (tid == 0) {
      s_piv = red_i[0];
      if (k == 0) s_maxpivot = red_v[0];
      // Eigen FullPivLU: rank = #pivots with |pivot| > maxpivot * eps * size; a zero corner ends the factorisation
      if (s_rank == m && !(red_v[0] > s_maxpivot * 2.220446049250313e-16 * m)) s_rank = k;
    }
    __syncthreads();
    if (s_rank != m) break;
    const int pi = s_piv / m, pj = s_piv % m;
    // swap rows k <-> pi of [W | B], columns k <-> pj of W
    if (pi != k)
      for (int j = tid; j < ld; j += kLitThreads) { const double t = W[k * ld + j]; W[k * ld + j] = W[pi * ld + j]; W[pi * ld + j] = t; }
    __syncthreads();
    if (pj != k) {
      for (int i = tid; i < m; i += kLitThreads) { const double t = W[i * ld + k]; W[i * ld + k] = W[i * ld + pj]; W[i * ld + pj] = t; }
      if (tid == 0) { const int t = colperm[k]; colperm[k] = colperm[pj]; colperm[pj] = t; }
    }
    __syncthreads();
    const double pinv = 1.0 / W[k * ld + k];
    for (int j = tid; j < ld; j += kLitThreads) W[k * ld + j] *= pinv;
    __syncthreads();
    // eliminate column k from every other row: thread per column of the augmented matrix, rows in the inner loop
    for (int j = tid; j < ld; j += kLitThreads) {
      if (j == k || (j < k)) continue;       // columns < k of W are already unit vectors
      const double rkj = W[k * ld + j];
      if (rkj == 0.0) continue;
      for (int i = 0; i < m; ++i)
        if (i != k) W[i * ld + j] = fma(-W[i * ld + k], rkj, W[i * ld + j]);
    }
    __syncthreads();
    for (int i = tid; i < m; i += kLitThreads)
      if (i != k) W[i * ld + k] = 0.0;
    __syncthreads();
  }
  const int rank = s_rank;
  if (tid == 0 && rank_out) rank_out[blockIdx.x] = rank;
  // A^-1 = Q B: row colperm[k] of the inverse = row k of B (dependent unknowns zero-filled, like FullPivLU::solve)
  // Lamda_prior = A_rr - A_rm A_mm^-1 A_rm^T
  double* inv = inv_out ? inv_out + (size_t)blockIdx.x * m * m : nullptr;
  if (inv) {
    for (int idx = tid; idx < m * m; idx += kLitThreads) {
      const int k = idx / m, c = idx - k * m;
      inv[colperm[k] + (size_t)m * c] = k < rank ? W[k * ld + m + c] : 0.0;
    }
  }
  __syncthreads();
  // T (m x m0) = A_mm^-1 A_mr into the (dead) left half of W: T[colperm[k]][c] = sum_j B[k][j] A[m0 + j][c]
  for (int idx = tid; idx < m * m0; idx += kLitThreads) {
    const int k = idx / m0, c = idx - k * m0;
    double acc = 0.0;
    if (k < rank)
      for (int j = 0; j < m; ++j) acc = fma(W[k * ld + m + j], A[(m0 + j) + (size_t)n * c], acc);
    W[k * ld + c] = acc;      // row k of W holds T[colperm[k]][:]
  }
  __syncthreads();
  double* P = prior_out + (size_t)blockIdx.x * m0 * m0;
  for (int idx = tid; idx < m0 * m0; idx += kLitThreads) {
    const int r = idx % m0, c = idx / m0;
    double acc = A[r + (size_t)n * c];
    for (int k = 0; k < m; ++k) acc = fma(-A[r + (size_t)n * (m0 + colperm[k])], W[k * ld + c], acc);
    P[idx] = acc;
  }
}

// ---- small dense helpers in local memory (thread-serial; N is a compile-time bound, n the live size) -------------------
template <int N>
__host__ __device__ __forceinline__ double loc_logdet_lu(const double* A, int n, double* inv, double* M) {
  // PartialPivLU (Eigen's determinant() / inverse() for dynamic matrices): returns log(det) (NaN for det <= 0, like
  // log() of a negative determinant in the reference) and, if inv != nullptr, the inverse (column-major, ld n).
  // M: n * n doubles of work space (the big matrices of this file live in global scratch, not on the thread stack)
  int perm[N];
  for (int i = 0; i < n * n; ++i) M[i] = A[i];
  for (int i = 0; i < n; ++i) perm[i] = i;
  double det = 1.0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double bv = fabs(M[k + n * k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(M[i + n * k]) > bv) { bv = fabs(M[i + n * k]); p = i; }
    if (p != k) {
      for (int j = 0; j < n; ++j) { const double t = M[k + n * j]; M[k + n * j] = M[p + n * j]; M[p + n * j] = t; }
      const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
      det = -det;
    }
    const double piv = M[k + n * k];
    det *= piv;
    for (int i = k + 1; i < n; ++i) {
      const double f = M[i + n * k] / piv;
      M[i + n * k] = f;
      for (int j = k + 1; j < n; ++j) M[i + n * j] = fma(-f, M[k + n * j], M[i + n * j]);
    }
  }
  if (inv) {
    for (int c = 0; c < n; ++c) {
      double x[N];
      for (int i = 0; i < n; ++i) x[i] = (perm[i] == c) ? 1.0 : 0.0;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) x[i] = fma(-M[i + n * j], x[j], x[i]);
      for (int i = n - 1; i >= 0; --i) {
        for (int j = i + 1; j < n; ++j) x[i] = fma(-M[i + n * j], x[j], x[i]);
        x[i] /= M[i + n * i];
      }
      for (int i = 0; i < n; ++i) inv[i + n * c] = x[i];
    }
  }
  return log(det);
}

// cyclic Jacobi on a symmetric n x n matrix (column-major, destroyed): eigenvalues on the diagonal, eigenvectors in V
template <int N>
__host__ __device__ __forceinline__ void loc_jacobi_eig(double* A, int n, double* V) {
  for (int i = 0; i < n * n; ++i) V[i] = 0.0;
  for (int i = 0; i < n; ++i) V[i + n * i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) { if (i != j) off += A[i + n * j] * A[i + n * j]; else dia += A[i + n * j] * A[i + n * j]; }
    if (off <= 1e-32 * dia) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p + n * q];
        if (apq == 0.0) continue;
        const double theta = (A[q + n * q] - A[p + n * p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k + n * p], akq = A[k + n * q];
          A[k + n * p] = c * akp - s * akq;
          A[k + n * q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p + n * k], aqk = A[q + n * k];
          A[p + n * k] = c * apk - s * aqk;
          A[q + n * k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k + n * p], vkq = V[k + n * q];
          V[k + n * p] = c * vkp - s * vkq;
          V[k + n * q] = s * vkp + c * vkq;
        }
      }
  }
}

constexpr int kKldScratch = 8 * 441 + 96;

struct isv_kld_args {
  int n;
  const double* pose_bwd;      // [n][2][7]
  const double* fj;            // factor-Jacobian scratch of the batch call ([n][kFJ])
  const double* lp_fwd;        // [n][36] structured Lamda_prior (debug store of the tail kernel)
  const double* g_bwd;         // [n][315] G (15 x 21, row k at [21 k + c]) (debug store of the backward kernel)
  const double* se3_out;       // outputs of the batch call
  const double* rel_out;
  const double* vb_out;
  const double* rp_out;
  const int32_t* rank;         // [n][2]
  double alpha;
  double* kld_fwd;             // [n]
  double* kld_bwd;             // [n]
  double* lp_bwd;              // [n][441] or null
  double* eig_bwd;             // [n][21] or null (ascending)
  double* info_abs;            // [n][9] or null
  double* info_yaw;            // [n] or null
  double* scratch;             // [n][kKldScratch] work space (global memory)
};

__global__ void __launch_bounds__(64)
marg_kld_kernel(isv_kld_args a) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.n) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  // ---- forward (:1333-1345): X = covi^-1 = s^T s ; phi = Jr^T X Jr ; cov = Lamda_prior^-1 (only when qr.rank() == 6)
  if (a.kld_fwd) {
    double k = nan;
    if (a.lp_fwd && a.rank[2 * w] == 6) {
      const double* Lp = a.lp_fwd + (size_t)w * 36;
      const double* Jr = a.fj + (size_t)w * kFJ + kFJ_JR6;          // 6 x 6 column-major
      const double* s = a.se3_out + (size_t)w * ISV_SE3_REC + 12;   // 6 x 6 column-major, upper triangular
      double X[36], T[36], phi[36], cov[36], M6[36];
      for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) { double acc = 0.0; for (int l = 0; l < 6; ++l) acc = fma(s[l + 6 * r], s[l + 6 * c], acc); X[r + 6 * c] = acc; }
      for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) { double acc = 0.0; for (int l = 0; l < 6; ++l) acc = fma(X[r + 6 * l], Jr[l + 6 * c], acc); T[r + 6 * c] = acc; }
      for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) { double acc = 0.0; for (int l = 0; l < 6; ++l) acc = fma(Jr[l + 6 * r], T[l + 6 * c], acc); phi[r + 6 * c] = acc; }
      const double ld_lp = loc_logdet_lu<6>(Lp, 6, cov, M6);           // log det cov = -log det Lamda_prior
      const double ld_phi = loc_logdet_lu<6>(phi, 6, nullptr, M6);
      double tr = 0.0;
      for (int r = 0; r < 6; ++r)
        for (int l = 0; l < 6; ++l) tr = fma(phi[r + 6 * l], cov[l + 6 * r], tr);
      const double ld_cov = loc_logdet_lu<6>(cov, 6, nullptr, M6);
      k = 0.5 * (tr - ld_phi - ld_cov - 6.0);
      (void)ld_lp;
    }
    a.kld_fwd[w] = k;
  }
  // ---- backward (:1456-1534) ----------------------------------------------------------------------------------------
  if (a.kld_bwd && a.g_bwd) {
    const double* G = a.g_bwd + (size_t)w * 315;
    double* S = a.scratch + (size_t)w * kKldScratch;
    double *Lam = S, *U = S + 441, *Jr = S + 2 * 441, *JU = S + 3 * 441, *X = S + 4 * 441, *XJ = S + 5 * 441, *Am = S + 6 * 441,
           *Mw = S + 7 * 441;
    for (int c = 0; c < 21; ++c)
      for (int r = 0; r < 21; ++r) { double acc = 0.0; for (int k = 0; k < 15; ++k) acc = fma(G[21 * k + r], G[21 * k + c], acc); Lam[r + 21 * c] = acc; }
    if (a.lp_bwd)
      for (int i = 0; i < 441; ++i) a.lp_bwd[(size_t)w * 441 + i] = Lam[i];
    loc_jacobi_eig<21>(Lam, 21, U);
    // ascending order of the eigenvalues, strict cut (:1479-1485)
    // (index / eigenvalue tables in the global scratch as well: with ~30 KB of thread-local matrices around them the
    // compiler was seen to let small local arrays share stack slots)
    double* lam = S + 8 * 441;
    int* ord = reinterpret_cast<int*>(S + 8 * 441 + 24);
    int* keep = ord + 24;
    for (int i = 0; i < 21; ++i) { ord[i] = i; lam[i] = Lam[i + 21 * i]; }
    for (int i = 1; i < 21; ++i) { const int o = ord[i]; int j = i - 1; while (j >= 0 && lam[ord[j]] > lam[o]) { ord[j + 1] = ord[j]; --j; } ord[j + 1] = o; }
    if (a.eig_bwd)
      for (int i = 0; i < 21; ++i) a.eig_bwd[(size_t)w * 21 + i] = lam[ord[i]];
    int rank = 0;
    for (int i = 0; i < 21; ++i)
      if (lam[ord[i]] > a.alpha) keep[rank++] = ord[i];
    // Jr (21 x 21, :1456-1464), column-major
    for (int i = 0; i < 441; ++i) Jr[i] = 0.0;
    const double* F = a.fj + (size_t)w * kFJ;
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 6; ++r) { Jr[r + 21 * (15 + c)] = F[kFJ_REL + r + 6 * c]; Jr[r + 21 * c] = F[kFJ_REL + 36 + r + 6 * c]; }
    for (int r = 0; r < 9; ++r) Jr[(6 + r) + 21 * (6 + r)] = 1.0;
    for (int c = 0; c < 6; ++c)
      for (int r = 0; r < 2; ++r) Jr[(15 + r) + 21 * (15 + c)] = F[kFJ_RP + r + 2 * c];
    for (int r = 0; r < 3; ++r) Jr[(17 + r) + 21 * (15 + r)] = 1.0;
    {
      const double* pose_i = a.pose_bwd + (size_t)w * 14;
      Quat Qw = quat_from_pose(pose_i);
      const double ux[3] = {1.0, 0.0, 0.0};
      double ym[3], Jy[6];
      qrot(qinv(Qw), ux, ym);                                       // yaw_meas = Qw.inverse() * UnitX (yaw_factor.h:17)
      yaw_jacobian(pose_i, ym, Jy, nullptr);
      for (int c = 0; c < 6; ++c) Jr[20 + 21 * (15 + c)] = Jy[c];
    }
    // JU = Jr U (21 x rank)
    for (int k = 0; k < rank; ++k)
      for (int r = 0; r < 21; ++r) { double acc = 0.0; for (int l = 0; l < 21; ++l) acc = fma(Jr[r + 21 * l], U[l + 21 * keep[k]], acc); JU[r + 21 * k] = acc; }
    // X = blkdiag(info_rel @0, info_vb @6, info_rp @15, info_abs @17, info_yaw @20); info = s^T s of the recovered factors,
    // abs / yaw recovered here: (J U D^-1 (J U)^T)^-1 (:1518-1519)
    for (int i = 0; i < 441; ++i) X[i] = 0.0;
    auto add_sts = [&](const double* s, int nn, int off) {
      for (int c = 0; c < nn; ++c)
        for (int r = 0; r < nn; ++r) { double acc = 0.0; for (int l = 0; l < nn; ++l) acc = fma(s[l + nn * r], s[l + nn * c], acc); X[(off + r) + 21 * (off + c)] += acc; }
    };
    add_sts(a.rel_out + (size_t)w * ISV_REL_REC + 12, 6, 0);
    add_sts(a.vb_out + (size_t)w * ISV_VB_REC + 9, 9, 6);
    add_sts(a.rp_out + (size_t)w * ISV_RP_REC + 9, 2, 15);
    {
      double C3[9], I3[9];
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) { double acc = 0.0; for (int k = 0; k < rank; ++k) acc += JU[(17 + r) + 21 * k] * JU[(17 + c) + 21 * k] / lam[keep[k]]; C3[r + 3 * c] = acc; }
      loc_logdet_lu<3>(C3, 3, I3, Mw);
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) { X[(17 + r) + 21 * (17 + c)] += I3[r + 3 * c]; if (a.info_abs) a.info_abs[(size_t)w * 9 + r + 3 * c] = I3[r + 3 * c]; }
      double cy = 0.0;
      for (int k = 0; k < rank; ++k) cy += JU[20 + 21 * k] * JU[20 + 21 * k] / lam[keep[k]];
      X[20 + 21 * 20] += 1.0 / cy;
      if (a.info_yaw) a.info_yaw[w] = 1.0 / cy;
    }
    // A = (Jr U)^T X (Jr U) (rank x rank) ; kld = 0.5 (tr(A D^-1) - log det A - log det D^-1 - 21)
    for (int k = 0; k < rank; ++k)
      for (int r = 0; r < 21; ++r) { double acc = 0.0; for (int l = 0; l < 21; ++l) acc = fma(X[r + 21 * l], JU[l + 21 * k], acc); XJ[r + 21 * k] = acc; }
    for (int c = 0; c < rank; ++c)
      for (int r = 0; r < rank; ++r) { double acc = 0.0; for (int l = 0; l < 21; ++l) acc = fma(JU[l + 21 * r], XJ[l + 21 * c], acc); Am[r + rank * c] = acc; }
    double tr = 0.0, ld_dinv = 0.0;
    for (int k = 0; k < rank; ++k) { tr += Am[k + rank * k] / lam[keep[k]]; ld_dinv -= log(lam[keep[k]]); }
    const double ld_a = loc_logdet_lu<21>(Am, rank, nullptr, Mw);
    a.kld_bwd[w] = 0.5 * (tr - ld_a - ld_dinv - 21.0);
  }
}

}  // namespace isv
