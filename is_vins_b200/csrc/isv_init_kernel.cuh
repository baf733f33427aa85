// initFactorGraph sparsification tail (/root/reference/src/estimator.cpp:745-1001), one warp (one
// CTA) per window, square-root-information form:
//   M (15(V-1) x 15V) = stacked whitened IMU Jacobians L_i^-1 J_i of the V-1 IMU factors (:774-803),
//   columns in OrderMap order (:747-758: T_0..T_{V-1}@6i, VB_{V-1}@6V, VB_0..VB_{V-2}@6V+9+9i);
//   Householder-eliminate the 9(V-1) VB_0..VB_{V-2} columns (Schur complement :808-816);
//   one-sided Jacobi on the remaining 6(V-1) rows  (SelfAdjointEigenSolver + `> ALPHA` :920-940);
//   per recovered factor  cov_i = sum_k (J_i g_k)(J_i g_k)^T / |g_k|^4 ,  sqrt_info = LLT(cov^-1).L^T
//   for the V-1 RelativePoseFactors, the SE3PriorFactor on T_0 and the Linear9Factor on VB_{V-1}
//   (:821-858, :873-972).  Runs once per session; sized for clarity, not for speed.
#pragma once
#include "isv_window_kernels.cuh"

namespace isv {

__device__ __forceinline__ int init_sb_col(int i, int V) { return (i == V - 1) ? 6 * V : 6 * V + 9 + 9 * i; }

// shared memory (doubles): M [rows*cols, ld = rows|1] | P 225 | vbuf rows | lam r | sc 64 | J 72 | JU 9*r | cov 81
__host__ __device__ inline int init_ld(int V) { return (15 * (V - 1)) | 1; }
__host__ __device__ inline size_t init_smem_doubles(int V) {
  const int rows = 15 * (V - 1), cols = 15 * V, r = 6 * (V - 1);
  return (size_t)init_ld(V) * cols + 225 + rows + r + 64 + 72 + 9 * r + 81 + 8;
}

__global__ void __launch_bounds__(32)
init_sparsify_kernel(int n_windows, int V, const double* __restrict__ poses, const double* __restrict__ sbs,
                     const double* __restrict__ preint, double* rel_out, double* se3_out, double* vb_out,
                     int32_t* rank_out, int32_t* status_out, DevCfg cfg) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int win = blockIdx.x;
  if (win >= n_windows) return;
  const int rows = 15 * (V - 1), cols = 15 * V, r = 6 * (V - 1), nk = 6 * V + 9, nm = 9 * (V - 1);
  const int ld = init_ld(V);
  double* M = smem;
  double* P = M + (size_t)ld * cols;
  double* vbuf = P + 225;
  double* lam = vbuf + rows;
  double* sc = lam + r;
  double* Jb = sc + 64;
  double* JU = Jb + 72;
  double* cov = JU + 9 * r;
  int status = 0, nonfinite = 0;
  const double* ps = poses + (size_t)win * V * 7;
  const double* sb = sbs + (size_t)win * V * 9;
  for (int i = lane; i < ld * cols; i += 32) M[i] = 0.0;
  if (lane < 3) sc[lane] = cfg.g[lane];
  __syncwarp();
  // ---- V-1 IMU factors, whitened in place (:774-803) -------------------------------------------
  for (int f = 0; f < V - 1; ++f) {
    const double* pre = preint + ((size_t)win * (V - 1) + f) * ISV_PREINT_REC;
    for (int i = lane; i < 225; i += 32) P[i] = pre[17 + 225 + i];
    const int ci_p = 6 * f, ci_s = init_sb_col(f, V), cj_p = 6 * (f + 1), cj_s = init_sb_col(f + 1, V);
    if (lane == 0) {
      if (nonunit(ps + 7 * f) || nonunit(ps + 7 * (f + 1))) status |= ISV_W_NONUNIT_QUAT;
      imu_jacobians(ps + 7 * f, sb + 9 * f, ps + 7 * (f + 1), sb + 9 * (f + 1), pre, sc, M + 15 * f, ld, ci_p, ci_s,
                    cj_p, cj_s, nullptr);
    }
    __syncwarp();
    if (w_chol_lower(P, 15, 15, lane)) status |= ISV_W_NOT_SPD;
    if (lane < 30) {
      const int c = lane < 6 ? ci_p + lane : (lane < 15 ? ci_s + lane - 6 : (lane < 21 ? cj_p + lane - 15 : cj_s + lane - 21));
      double* col = M + 15 * f + (size_t)ld * c;
      for (int i = 0; i < 15; ++i) {
        double s = col[i];
        for (int l = 0; l < i; ++l) s = fma(-P[i + 15 * l], col[l], s);
        col[i] = s / P[i + 15 * i];
      }
    }
    __syncwarp();
  }
  // ---- Schur complement over VB_0..VB_{V-2} (:808-816) ----------------------------------------------
  w_householder_marginalize(M, ld, rows, cols, nk, nm, vbuf, lane);
  // ---- eigen-decomposition of Lamda_prior = G^T G, G = rows nm.. x columns 0..nk-1 (:920-940) -------
  double* G = M + nm;  // row k, element c at G[k + ld * c]
  if (w_onesided_jacobi_rows<1, 0>(G, 1, r, nk, lam, lane, 40, ld) >= 40) status |= ISV_W_EIG_NOCONV;
  int rank = 0;
  for (int k = lane; k < r; k += 32) {
    const double l = lam[k];
    const int keep = l > cfg.alpha;
    lam[k] = keep ? 1.0 / (l * l) : 0.0;
    rank += keep;
  }
  rank = (int)warp_sum((double)rank);
  __syncwarp();
  // ---- recovered factors (:821-858) and their information (:942-972) ---------------------------
  auto recover6 = [&](const double* Ja, int ca, const double* Jb2, int cb, double* out) {
    // JU (6 x r) = Ja * G[:, ca:ca+6]^T (+ Jb2 * G[:, cb:cb+6]^T)
    for (int idx = lane; idx < 6 * r; idx += 32) {
      const int a = idx % 6, k = idx / 6;
      double acc = 0.0;
      for (int c = 0; c < 6; ++c) {
        acc = fma(Ja[a + 6 * c], G[k + (size_t)ld * (ca + c)], acc);
        if (Jb2) acc = fma(Jb2[a + 6 * c], G[k + (size_t)ld * (cb + c)], acc);
      }
      JU[idx] = acc;
    }
    __syncwarp();
    for (int idx = lane; idx < 36; idx += 32) {
      const int a = idx % 6, b = idx / 6;
      double acc = 0.0;
      for (int k = 0; k < r; ++k) acc = fma(JU[a + 6 * k] * lam[k], JU[b + 6 * k], acc);
      cov[idx] = acc;
    }
    __syncwarp();
    if (w_sqrt_info_from_cov<6>(cov, 6, out, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  };
  for (int f = 0; f < V - 1; ++f) {
    double* o = rel_out + ((size_t)win * (V - 1) + f) * ISV_REL_REC;
    if (lane == 0) {
      const double* pi = ps + 7 * f;
      const double* pj = ps + 7 * (f + 1);
      Quat Qi = quat_from_pose(pi), Qj = quat_from_pose(pj);
      double dd[3] = {pj[0] - pi[0], pj[1] - pi[1], pj[2] - pi[2]};
      double tij[3], Rij[9];
      qrot(qinv(Qi), dd, tij);
      q2R(qmul(qinv(Qi), Qj), Rij);
      relpose_jacobians(pi, pj, tij, Rij, Jb, Jb + 36, nullptr);
      for (int i = 0; i < 3; ++i) o[i] = tij[i];
      for (int rr = 0; rr < 3; ++rr)
        for (int c = 0; c < 3; ++c) o[3 + rr + 3 * c] = Rij[3 * rr + c];
    }
    __syncwarp();
    recover6(Jb, 6 * f, Jb + 36, 6 * (f + 1), o + 12);
  }
  {  // SE3PriorFactor(Ps0, Q0) on T_0 (:840-847)
    double* o = se3_out + (size_t)win * ISV_SE3_REC;
    if (lane == 0) {
      double R0[9], t0[3] = {ps[0], ps[1], ps[2]};
      q2R(quat_from_pose(ps), R0);
      se3prior_jacobian(ps, t0, R0, Jb, nullptr);
      for (int i = 0; i < 3; ++i) o[i] = t0[i];
      for (int rr = 0; rr < 3; ++rr)
        for (int c = 0; c < 3; ++c) o[3 + rr + 3 * c] = R0[3 * rr + c];
    }
    __syncwarp();
    recover6(Jb, 0, nullptr, 0, o + 12);
  }
  {  // Linear9Factor(VB_{V-1}) (:850-858): J = I9 at columns 6V..6V+8
    double* o = vb_out + (size_t)win * ISV_VB_REC;
    if (lane < 9) o[lane] = sb[9 * (V - 1) + lane];
    for (int idx = lane; idx < 81; idx += 32) {
      const int a = idx % 9, b = idx / 9;
      double acc = 0.0;
      for (int k = 0; k < r; ++k) acc = fma(G[k + (size_t)ld * (6 * V + a)] * lam[k], G[k + (size_t)ld * (6 * V + b)], acc);
      cov[idx] = acc;
    }
    __syncwarp();
    if (w_sqrt_info_from_cov<9>(cov, 9, o + 9, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  }
  if (__any_sync(kFullMask, nonfinite)) status |= ISV_W_NONFINITE;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) status |= __shfl_xor_sync(kFullMask, status, o);
  if (lane == 0) {
    rank_out[win] = rank;
    if (status_out) status_out[win] = status;
  }
}

}  // namespace isv
