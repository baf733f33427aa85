// The two window kernels of the hot path: one WARP per sliding window, every stage fused, so that
// only the window's inputs and its recovered factors touch HBM (SURVEY.md 7.2 "KF").
//
//   marg_forward_kernel   Estimator::MargForward    /root/reference/src/estimator.cpp:1149-1352
//   marg_backward_kernel  Estimator::MargBackward   /root/reference/src/estimator.cpp:1354-1539
//
// MargForward, structured form.  The reference builds the dense (12+L)^2 `Lamda` with the block
// loop :1168-1202 and inverts the (L+6)^2 block with FullPivLU (:1286).  The landmark block of
// Lamda is diagonal (every ProjectionFactor touches exactly one inverse depth), so eliminating the
// landmarks first is algebraically identical and O(L):
//     s*J_k = [a_k; c_k] (2x12, columns ordered [T1|T0] as OrderMap :1153-1158), s*jl_k (2x1)
//     u_k = s*jl_k/|s*jl_k| , v_k _|_ u_k ;  e_k = (sJ_k)^T u_k , w_k = (sJ_k)^T v_k
//     Lamda[0:12,0:12]            = sum_k e_k e_k^T + w_k w_k^T   (+ prior + rel-pose blocks)
//     Lamda[0:12,0:12] - B D^-1 B^T = sum_k w_k w_k^T             (Schur over the landmarks)
// Both sums are SYRKs  X X^T  with X 12 x L: they run on the FP64 tensor pipe
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), 32 landmarks per warp iteration, staged through
// shared memory to reach the fragment layout.  Everything after that is 12x12 / 6x6 algebra.
#pragma once
#include "isv_device_math.cuh"
#include "isv_factors.cuh"
#include "isv_small_qr.cuh"
#include "isv_warp_linalg.cuh"

#include "../../include/isv_capi.h"

namespace isv {

struct DevCfg {
  double alpha;
  double ps[4];  // ProjectionFactor::sqrt_info, column-major 2x2
  double g[3];
  double qr_threshold;
};

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = 32 * kWarpsPerCta;
#ifndef ISV_FWD_MINB
#define ISV_FWD_MINB 1
#endif
#ifndef ISV_BWD_MINB
#define ISV_BWD_MINB 4
#endif

// ---- forward: shared-memory map (doubles, per warp) -----------------------------------------
constexpr int kXld = 36;                     // staging row stride: conflict-free 64-bit fragment loads
constexpr int kFwdConst = 96;                // per-window constants of the landmark phase
constexpr int kFwdWork = 832;                // staging (16*36) during the loop, tail matrices after it
constexpr int kFwdSmemPerWarp = kFwdConst + kFwdWork;
// ---- backward ---------------------------------------------------------------------------------
constexpr int kBwdScratch = 136;
constexpr int kBwdSmemPerWarp = 750 + 225 + kBwdScratch;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int nonunit(const double* ps) {
  double n2 = ps[3] * ps[3] + ps[4] * ps[4] + ps[5] * ps[5] + ps[6] * ps[6];
  return fabs(n2 - 1.0) > 1e-9;
}

// out (n x n col-major, global) = upper-triangular Eigen `LLT(M).matrixL().transpose()`;
// M (smem, ld) is destroyed.  Returns 1 when M is not SPD.
__device__ __forceinline__ int chol_store_upper(double* M, int ld, int n, double* out, int lane, int& nonfinite) {
  int bad = w_chol_lower(M, ld, n, lane);
  for (int idx = lane; idx < n * n; idx += 32) {
    int i = idx % n, j = idx / n;
    double v = (i <= j) ? M[j + i * ld] : 0.0;
    if (!isfinite(v)) nonfinite = 1;
    out[idx] = v;
  }
  __syncwarp();
  return bad;
}

// =================================================================================================
// MargForward
// =================================================================================================
__global__ void __launch_bounds__(kThreads, ISV_FWD_MINB)
marg_forward_kernel(isv_batch_in in, isv_batch_out out, DevCfg cfg) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int win = blockIdx.x * kWarpsPerCta + warp;
  if (win >= in.n_windows) return;
  double* K = smem + warp * kFwdSmemPerWarp;  // constants
  double* X = K + kFwdConst;                  // staging / tail work
  int status = 0;
  int nonfinite = 0;

  const double* pose0 = in.pose_fwd + (size_t)win * 14;
  const double* pose1 = pose0 + 7;
  const double* ex = in.ex_pose + (in.ex_pose_shared ? 0 : (size_t)win * 7);

  // ---- per-window constants (projection_factor.cpp:127-147) -----------------------------------
  // K: [0]ric [9]tic [12]R0 [21]P0 [24]R1 [33]P1 [36]B=ric^T R1^T [45]C=B R0 [54]Ap=C ric
  if (lane == 0) {
    Quat Qi = quat_from_pose(pose0), Qj = quat_from_pose(pose1), qic = quat_from_pose(ex);
    if (nonunit(pose0) || nonunit(pose1) || nonunit(ex)) status |= ISV_W_NONUNIT_QUAT;
    double ric[9], R0[9], R1[9], B[9], C[9], Ap[9], T[9];
    q2R(qic, ric);
    q2R(Qi, R0);
    q2R(Qj, R1);
    // B = ric^T * R1^T = (R1 * ric)^T
    mat3_mul(R1, ric, T);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) B[3 * r + c] = T[3 * c + r];
    mat3_mul(B, R0, C);
    mat3_mul(C, ric, Ap);
    for (int i = 0; i < 9; ++i) {
      K[i] = ric[i]; K[12 + i] = R0[i]; K[24 + i] = R1[i]; K[36 + i] = B[i]; K[45 + i] = C[i]; K[54 + i] = Ap[i];
    }
    for (int i = 0; i < 3; ++i) { K[9 + i] = ex[i]; K[21 + i] = pose0[i]; K[33 + i] = pose1[i]; }
  }
  // zero the staging buffer once: rows 12..15 stay zero for the whole loop
  for (int i = lane; i < 16 * kXld; i += 32) X[i] = 0.0;
  __syncwarp();

  const long long lm0 = in.lm_offset[win];
  const int L = (int)(in.lm_offset[win + 1] - lm0);
  const double* ob = in.lm_obs + lm0;
  const long long st = in.lm_stride;
  const double s00 = cfg.ps[0], s10 = cfg.ps[1], s01 = cfg.ps[2], s11 = cfg.ps[3];

  // accumulators: tiles (0,0) (1,0) (1,1) of E = sum e e^T and S = sum w w^T
  double e00a = 0, e00b = 0, e10a = 0, e10b = 0, e11a = 0, e11b = 0;
  double s00a = 0, s00b = 0, s10a = 0, s10b = 0, s11a = 0, s11b = 0;
  const int fm = lane >> 2, fk = lane & 3;

  for (int base = 0; base < L; base += 32) {
    const int k = base + lane;
    double ev[12], wv[12];
    if (k < L) {
      const double xi = ob[k], yi = ob[st + k], zi = ob[2 * st + k];
      const double lam = ob[5 * st + k];
      // pts_camera_i = pts_i / inv_dep ; pts_imu_i = ric*pc + tic ; pts_w = R0*pi + P0
      const double pc0 = xi / lam, pc1 = yi / lam, pc2 = zi / lam;
      double pi_[3], pw[3], pj[3], d[3], cj[3];
      for (int r = 0; r < 3; ++r) pi_[r] = K[3 * r] * pc0 + K[3 * r + 1] * pc1 + K[3 * r + 2] * pc2 + K[9 + r];
      for (int r = 0; r < 3; ++r)
        pw[r] = K[12 + 3 * r] * pi_[0] + K[12 + 3 * r + 1] * pi_[1] + K[12 + 3 * r + 2] * pi_[2] + K[21 + r];
      for (int r = 0; r < 3; ++r) d[r] = pw[r] - K[33 + r];
      // pts_imu_j = R1^T (pw - P1) ; pts_camera_j = ric^T (pts_imu_j - tic)
      for (int r = 0; r < 3; ++r) pj[r] = K[24 + r] * d[0] + K[24 + 3 + r] * d[1] + K[24 + 6 + r] * d[2];
      for (int r = 0; r < 3; ++r) d[r] = pj[r] - K[9 + r];
      for (int r = 0; r < 3; ++r) cj[r] = K[r] * d[0] + K[3 + r] * d[1] + K[6 + r] * d[2];
      const double iz = 1.0 / cj[2];
      const double rx = -cj[0] * iz * iz, ry = -cj[1] * iz * iz;  // reduce = [iz 0 rx ; 0 iz ry]
      // rB = reduce*B, rC = reduce*C, rT = reduce*ric^T   (2x3 each)
      double rB[6], rC[6], rT[6];
      for (int c = 0; c < 3; ++c) {
        rB[c] = iz * K[36 + c] + rx * K[36 + 6 + c];
        rB[3 + c] = iz * K[36 + 3 + c] + ry * K[36 + 6 + c];
        rC[c] = iz * K[45 + c] + rx * K[45 + 6 + c];
        rC[3 + c] = iz * K[45 + 3 + c] + ry * K[45 + 6 + c];
        rT[c] = iz * K[3 * c] + rx * K[3 * c + 2];        // ric^T[0][c] = ric[c][0]
        rT[3 + c] = iz * K[3 * c + 1] + ry * K[3 * c + 2];
      }
      // unweighted J (2 x 12), column order [T1 (d/dP1, d/dtheta1) | T0 (d/dP0, d/dtheta0)]
      double J0[12], J1[12];
      for (int c = 0; c < 3; ++c) {
        J0[c] = -rB[c];          J1[c] = -rB[3 + c];        // jaco_j left  = ric^T * -Rj^T
        J0[6 + c] = rB[c];       J1[6 + c] = rB[3 + c];     // jaco_i left  = ric^T * Rj^T
      }
      // M*skew(p) row = (M1 p2 - M2 p1, M2 p0 - M0 p2, M0 p1 - M1 p0)
      J0[3] = rT[1] * pj[2] - rT[2] * pj[1];  J0[4] = rT[2] * pj[0] - rT[0] * pj[2];  J0[5] = rT[0] * pj[1] - rT[1] * pj[0];
      J1[3] = rT[4] * pj[2] - rT[5] * pj[1];  J1[4] = rT[5] * pj[0] - rT[3] * pj[2];  J1[5] = rT[3] * pj[1] - rT[4] * pj[0];
      J0[9] = -(rC[1] * pi_[2] - rC[2] * pi_[1]);  J0[10] = -(rC[2] * pi_[0] - rC[0] * pi_[2]);  J0[11] = -(rC[0] * pi_[1] - rC[1] * pi_[0]);
      J1[9] = -(rC[4] * pi_[2] - rC[5] * pi_[1]);  J1[10] = -(rC[5] * pi_[0] - rC[3] * pi_[2]);  J1[11] = -(rC[3] * pi_[1] - rC[4] * pi_[0]);
      // jacobian_feature = reduce * Ap * pts_i * -1/(lam^2)
      double f[3];
      for (int r = 0; r < 3; ++r) f[r] = K[54 + 3 * r] * xi + K[54 + 3 * r + 1] * yi + K[54 + 3 * r + 2] * zi;
      const double sc = -1.0 / (lam * lam);
      const double jl0 = (iz * f[0] + rx * f[2]) * sc, jl1 = (iz * f[1] + ry * f[2]) * sc;
      // weighted by the 2x2 sqrt_info, then rotated into the (u, v) basis of s*jl
      const double g0 = s00 * jl0 + s01 * jl1, g1 = s10 * jl0 + s11 * jl1;
      double nrm = sqrt(g0 * g0 + g1 * g1);
      double u0 = 1.0, u1 = 0.0;
      if (nrm > 0.0) { u0 = g0 / nrm; u1 = g1 / nrm; } else { status |= ISV_W_SINGULAR; }
      // e = (sJ)^T u , w = (sJ)^T v with v = (-u1, u0)
      const double eu0 = u0 * s00 + u1 * s10, eu1 = u0 * s01 + u1 * s11;      // u^T s
      const double ev0 = -u1 * s00 + u0 * s10, ev1 = -u1 * s01 + u0 * s11;    // v^T s
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        ev[c] = eu0 * J0[c] + eu1 * J1[c];
        wv[c] = ev0 * J0[c] + ev1 * J1[c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < 12; ++c) { ev[c] = 0.0; wv[c] = 0.0; }
    }
    // ---- E += e e^T on the FP64 tensor pipe --------------------------------------------------
#pragma unroll
    for (int c = 0; c < 12; ++c) X[c * kXld + lane] = ev[c];
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const double lo = X[fm * kXld + 4 * s + fk], hi = X[(8 + fm) * kXld + 4 * s + fk];
      dmma884(e00a, e00b, lo, lo);
      dmma884(e10a, e10b, hi, lo);
      dmma884(e11a, e11b, hi, hi);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 12; ++c) X[c * kXld + lane] = wv[c];
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const double lo = X[fm * kXld + 4 * s + fk], hi = X[(8 + fm) * kXld + 4 * s + fk];
      dmma884(s00a, s00b, lo, lo);
      dmma884(s10a, s10b, hi, lo);
      dmma884(s11a, s11b, hi, hi);
    }
    __syncwarp();
  }

  // ---- tail -------------------------------------------------------------------------------------
  // work map (doubles): S12[0] E12/H12[144] T16e[288] T16s[544] ; after the unpack:
  // Wst[288] G[432] Jr6[504] sp[540] sr[576] tA[612] tB[684] wk[756]
  double* S12 = X;
  double* H12 = X + 144;
  double* T16e = X + 288;
  double* T16s = X + 544;
  {
    const int r = fm, c = 2 * fk;
    T16e[r + 16 * c] = e00a;             T16e[r + 16 * (c + 1)] = e00b;
    T16e[8 + r + 16 * c] = e10a;         T16e[8 + r + 16 * (c + 1)] = e10b;
    T16e[8 + r + 16 * (8 + c)] = e11a;   T16e[8 + r + 16 * (8 + c + 1)] = e11b;
    T16s[r + 16 * c] = s00a;             T16s[r + 16 * (c + 1)] = s00b;
    T16s[8 + r + 16 * c] = s10a;         T16s[8 + r + 16 * (c + 1)] = s10b;
    T16s[8 + r + 16 * (8 + c)] = s11a;   T16s[8 + r + 16 * (8 + c + 1)] = s11b;
  }
  __syncwarp();
  for (int idx = lane; idx < 144; idx += 32) {
    int r = idx % 12, c = idx / 12;
    int rr = r, cc = c;
    if (r < 8 && c >= 8) { rr = c; cc = r; }  // tile (0,1) = tile (1,0)^T
    S12[idx] = T16s[rr + 16 * cc];
    H12[idx] = T16e[rr + 16 * cc];
  }
  __syncwarp();
  double* Wst = X + 288;
  double* G = X + 432;
  double* Jr6 = X + 504;
  double* sp = X + 540;
  double* sr = X + 576;
  double* tA = X + 612;
  double* tB = X + 684;
  double* wk = X + 756;
  const double* pse3 = in.prior_se3 + (size_t)win * ISV_SE3_REC;
  const double* prel = in.prior_rel + (size_t)win * ISV_REL_REC;
  double* o_se3 = out.se3_out + (size_t)win * ISV_SE3_REC;
  double* o_pg = out.pg_out + (size_t)win * ISV_PG_REC;
  // tA <- Jp (36) | tB <- Ji (36), tB+36.. no: use wk for Ji/Jj
  for (int i = lane; i < 36; i += 32) { sp[i] = pse3[12 + i]; sr[i] = prel[12 + i]; }
  if (lane == 0) {
    double Rp[9];
    load_mat3_colmajor(pse3 + 3, Rp);
    se3prior_jacobian(pose0, pse3, Rp, tA, nullptr);  // vioPosePriorEdge->EvaluateOnlyJacobians(para_Pose[0])
  } else if (lane == 1) {
    double dR[9];
    load_mat3_colmajor(prel + 3, dR);
    relpose_jacobians(pose0, pose1, prel, dR, wk, wk + 36, nullptr);  // vioRelativePoseEdges[1]
  } else if (lane == 2) {
    // pose-graph factor at the current estimate (:1244-1255): tij, Rij, J = [Ji | Jj]
    Quat Qi = quat_from_pose(pose0), Qj = quat_from_pose(pose1);
    double dd[3] = {pose1[0] - pose0[0], pose1[1] - pose0[1], pose1[2] - pose0[2]};
    double tij[3], Rij[9];
    qrot(qinv(Qi), dd, tij);
    q2R(qmul(qinv(Qi), Qj), Rij);
    relpose_jacobians(pose0, pose1, tij, Rij, G, G + 36, nullptr);
    for (int i = 0; i < 3; ++i) o_pg[i] = tij[i];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) o_pg[3 + r + 3 * c] = Rij[3 * r + c];
    o_pg[84] = sqrt(tij[0] * tij[0] + tij[1] * tij[1] + tij[2] * tij[2]);  // distance = delta_t.norm()
  } else if (lane == 3) {
    // new SE3PriorFactor(P1, Q1) evaluated at para_Pose[1] (:1291-1297)
    Quat Q1 = quat_from_pose(pose1);
    double R1[9];
    q2R(Q1, R1);
    double t1[3] = {pose1[0], pose1[1], pose1[2]};
    se3prior_jacobian(pose1, t1, R1, Jr6, nullptr);
    for (int i = 0; i < 3; ++i) o_se3[i] = t1[i];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) o_se3[3 + r + 3 * c] = R1[3 * r + c];
  } else if (lane == 4) {
    // covAbs = (s^T s)^-1 of vioRollPitchEdges[0] when its index is 0 (:1265-1271)
    double cA[4] = {0, 0, 0, 0};
    if (in.prior_rp) {
      const double* rp = in.prior_rp + (size_t)win * ISV_RP_IN_REC;
      if (rp[0] != 0.0) {
        double a = rp[1], b = rp[2], c = rp[3], d = rp[4];  // col-major s = [a c; b d]
        double m00 = a * a + b * b, m01 = a * c + b * d, m11 = c * c + d * d;
        double det = m00 * m11 - m01 * m01;
        cA[0] = m11 / det; cA[1] = -m01 / det; cA[2] = -m01 / det; cA[3] = m00 / det;
      }
    }
    for (int i = 0; i < 4; ++i) o_pg[85 + i] = cA[i];
  }
  __syncwarp();
  // Wst (12x12): rows 0-5 = sp * [0 | Jp], rows 6-11 = sr * [Jj | Ji]   (OrderMap: T1@0, T0@6)
  for (int idx = lane; idx < 144; idx += 32) {
    int r = idx % 12, c = idx / 12;
    double acc = 0.0;
    if (r < 6) {
      if (c >= 6)
        for (int l = 0; l < 6; ++l) acc = fma(sp[r + 6 * l], tA[l + 6 * (c - 6)], acc);
    } else {
      const double* Jb = (c < 6) ? (wk + 36 + 6 * c) : (wk + 6 * (c - 6));
      for (int l = 0; l < 6; ++l) acc = fma(sr[(r - 6) + 6 * l], Jb[l], acc);
    }
    Wst[idx] = acc;
  }
  __syncwarp();
  // S12 += Wst^T Wst ; H12 = E12 + S12  (= Lamda[0:12,0:12], :1243)
  w_gemm<true, false>(12, 12, 12, Wst, 12, Wst, 12, S12, 12, 1, lane);
  for (int idx = lane; idx < 144; idx += 32) H12[idx] += S12[idx];
  __syncwarp();
  // ---- pose-graph relative-pose factor (:1243-1259) -------------------------------------------
  // J = G (6x12, [Ji|Jj]) ; Jpinv = J^T (J J^T)^-1 (full row rank) ; rpOmega = Jpinv^T H12 Jpinv
  w_gemm<false, true>(6, 6, 12, G, 6, G, 6, tA, 6, 0, lane);          // tA = J J^T
  if (w_inverse(tA, 6, 6, wk, lane)) status |= ISV_W_SINGULAR;
  w_gemm<true, false>(12, 6, 6, G, 6, tA, 6, tB, 12, 0, lane);        // tB = Jpinv (12x6)
  w_gemm<false, false>(12, 6, 12, H12, 12, tB, 12, wk, 12, 0, lane);  // wk = H12 Jpinv
  w_gemm<true, false>(6, 6, 12, tB, 12, wk, 12, tA, 6, 0, lane);      // tA = rpOmega
  w_copy(Wst, tA, 36, lane);
  if (chol_store_upper(Wst, 6, 6, o_pg + 12, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  if (w_inverse(tA, 6, 6, wk, lane)) status |= ISV_W_SINGULAR;        // covRel = rpOmega^-1
  for (int i = lane; i < 36; i += 32) {
    if (!isfinite(tA[i])) nonfinite = 1;
    o_pg[48 + i] = tA[i];
  }
  __syncwarp();
  // ---- Schur complement over T0 (:1286-1288 with the landmarks already eliminated) -------------
  w_copy2d(tA, 6, S12 + 6 + 12 * 6, 12, 6, 6, lane);                  // tA = S[6:12,6:12]
  if (w_inverse(tA, 6, 6, wk, lane)) status |= ISV_W_SINGULAR;
  w_gemm<false, false>(6, 6, 6, S12 + 12 * 6, 12, tA, 6, tB, 6, 0, lane);        // tB = S[0:6,6:12] Smm^-1
  w_copy2d(Wst, 6, S12, 12, 6, 6, lane);                                          // Wst = S[0:6,0:6]
  w_gemm<false, true>(6, 6, 6, tB, 6, S12 + 12 * 6, 12, Wst, 6, -1, lane);       // Lamda_prior (6x6)
  // ---- rank decision + recovery of the SE3 prior on T1 (:1304-1349) ---------------------------
  int rank = 0;
  if (lane == 0) {
    for (int i = 0; i < 36; ++i) tB[i] = Wst[i];
    rank = serial_fullpiv_qr_inverse<6>(tB, tA, cfg.qr_threshold);     // tA = cov = qr.solve(I)
  }
  rank = __shfl_sync(kFullMask, rank, 0);
  __syncwarp();
  int out_rank = rank;
  if (rank == 6) {
    w_gemm<false, false>(6, 6, 6, Jr6, 6, tA, 6, tB, 6, 0, lane);     // Jr cov
    w_gemm<false, true>(6, 6, 6, tB, 6, Jr6, 6, tA, 6, 0, lane);      // covi = Jr cov Jr^T
  } else {
    status |= ISV_W_RANK_DEFICIENT;
    // truncated eigen path (:1311-1331), factored form: Lamda_prior = sum_k g_k g_k^T (rows of tB)
    for (int i = lane; i < 36; i += 32) tB[i] = 0.0;
    __syncwarp();
    const int nrow = w_pivoted_cholesky_rows(Wst, 6, 6, tB, 6, wk, lane);
    if (w_onesided_jacobi_rows<4, 2>(tB, 6, nrow, 6, wk, lane) >= 30) status |= ISV_W_EIG_NOCONV;
    int er = 0;
    for (int k = 0; k < nrow; ++k) er += (wk[k] > cfg.alpha) ? 1 : 0;
    out_rank = er;
    for (int idx = lane; idx < 36; idx += 32) {   // wk+8 = Jr G^T (6 x nrow)
      int r = idx % 6, k = idx / 6;
      double acc = 0.0;
      for (int c = 0; c < 6; ++c) acc = fma(Jr6[r + 6 * c], tB[c + 6 * k], acc);
      wk[8 + idx] = acc;
    }
    __syncwarp();
    for (int idx = lane; idx < 36; idx += 32) {
      int r = idx % 6, c = idx / 6;
      double acc = 0.0;
      for (int k = 0; k < nrow; ++k) {
        double lamk = wk[k];
        if (lamk > cfg.alpha) acc += wk[8 + r + 6 * k] * wk[8 + c + 6 * k] / (lamk * lamk);
      }
      tA[idx] = acc;
    }
    __syncwarp();
  }
  if (w_inverse(tA, 6, 6, wk, lane)) status |= ISV_W_SINGULAR;        // covi.inverse()
  if (chol_store_upper(tA, 6, 6, o_se3 + 12, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  if (__any_sync(kFullMask, nonfinite)) status |= ISV_W_NONFINITE;
  // merge per-lane status bits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) status |= __shfl_xor_sync(kFullMask, status, o);
  if (lane == 0) {
    out.rank[2 * win] = out_rank;
    if (out.status) atomicOr(out.status + win, status);
  }
}

// =================================================================================================
// MargBackward -- square-root-information form.
//   M (24 x 30) = [ s_vb on VB_{V-1} ; L^-1 J_imu ]   with  covariance = L L^T   (rows 0-8 / 9-23)
//   Householder-eliminate the VB_{V-1} columns -> rows 9-23, columns 0-20 = G (15 x 21) with
//   G^T G = Lamda_prior (:1419);  one-sided Jacobi makes the rows of G orthogonal:
//   Lamda_prior = sum_k g_k g_k^T, eigenvalues |g_k|^2 (:1479-1497), and for every recovered factor
//   cov_i = J_i U D^-1 U^T J_i^T = sum_{|g_k|^2 > ALPHA} (J_i g_k)(J_i g_k)^T / |g_k|^4  (:1500-1516).
// Algebraically identical to the reference's information-form route, without its cancellation.
// =================================================================================================
constexpr int kMld = 25;  // odd leading dimension: conflict-free strided row access
__global__ void __launch_bounds__(kThreads, ISV_BWD_MINB)
marg_backward_kernel(isv_batch_in in, isv_batch_out out, DevCfg cfg, int vo_size) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int win = blockIdx.x * kWarpsPerCta + warp;
  if (win >= in.n_windows) return;
  double* M = smem + warp * kBwdSmemPerWarp;  // 24 x 30 (ld 25)
  double* P = M + 750;                        // 15 x 15 covariance -> Cholesky factor
  double* sc = M + 975;                       // scratch: poses [0..31], G[32..34], dinv[40..60], lam/vbuf[64..]
  int status = 0, nonfinite = 0;

  const double* pose_i = in.pose_bwd + (size_t)win * 14;
  const double* pose_j = pose_i + 7;
  const double* sb_i = in.sb_bwd + (size_t)win * 18;
  const double* sb_j = sb_i + 9;
  const double* pvb = in.prior_vb + (size_t)win * ISV_VB_REC;
  const double* pre = in.preint + (size_t)win * ISV_PREINT_REC;

  for (int i = lane; i < 750; i += 32) M[i] = 0.0;
  for (int i = lane; i < 225; i += 32) P[i] = pre[17 + 225 + i];
  if (lane < 7) { sc[lane] = pose_i[lane]; sc[16 + lane] = pose_j[lane]; }
  if (lane >= 7 && lane < 16) { sc[lane] = sb_i[lane - 7]; sc[16 + lane] = sb_j[lane - 7]; }
  if (lane < 3) sc[32 + lane] = cfg.g[lane];
  __syncwarp();
  // vioVBPrior (Linear9Factor, J = I9 on VB_{V-1}): rows 0-8, columns 21-29 = sqrt_info  (:1372-1380)
  for (int idx = lane; idx < 81; idx += 32) M[(idx % 9) + kMld * (21 + idx / 9)] = pvb[9 + idx];
  // ---- IMUFactor::Evaluate, tangent twin (imu_factor.h:161-265), columns in OrderMap order -----
  // OrderMap (:1358-1366): T_V@0, VB_V@6, T_{V-1}@15, VB_{V-1}@21
  if (lane == 0) {
    if (nonunit(sc) || nonunit(sc + 16)) status |= ISV_W_NONUNIT_QUAT;
    imu_jacobians(sc, sc + 7, sc + 16, sc + 23, pre, sc + 32, M + 9, kMld, 15, 21, 0, 6, nullptr);
  }
  __syncwarp();
  // sqrt_info^T sqrt_info = covariance^-1 (imu_factor.h:181): P = L L^T, rows 9-23 <- L^-1 J
  if (w_chol_lower(P, 15, 15, lane)) status |= ISV_W_NOT_SPD;
  if (lane < 30) {
    double* col = M + 9 + kMld * lane;
    for (int i = 0; i < 15; ++i) {
      double s = col[i];
      for (int l = 0; l < i; ++l) s = fma(-P[i + 15 * l], col[l], s);
      col[i] = s / P[i + 15 * i];
    }
  }
  __syncwarp();
  // ---- Schur complement over VB_{V-1} (:1413-1419) as a QR elimination -------------------------
  w_householder_marginalize(M, kMld, 24, 30, 21, 9, sc + 64, lane);
  // ---- eigen-decomposition (:1479-1497): orthogonalise the 15 rows of G ------------------------
  double* G = M + 9;  // row k, element c at G[k + kMld * c]
  if (w_onesided_jacobi_rows<4, 6>(G, 1, 15, 21, sc + 64, lane, 30, kMld) >= 30) status |= ISV_W_EIG_NOCONV;
  int rank = 0;
  if (lane < 21) {
    double lamk = (lane < 15) ? sc[64 + lane] : 0.0;
    int keep = lamk > cfg.alpha;  // strict, Q12
    sc[40 + lane] = keep ? 1.0 / (lamk * lamk) : 0.0;
    rank = keep;
  }
  rank = __popc(__ballot_sync(kFullMask, rank));
  __syncwarp();
  // ---- recovered factors (:1424-1452) and their Jacobian rows (:1456-1477) ---------------------
  double* T = M + 525;       // dead: R factor columns (225) + P (225) = 450 contiguous doubles
  double* Jrel = T;          // Ji (36) | Jj (36)
  double* Jrp = T + 72;      // 2 x 6
  double* JU = T + 88;       // up to 9 x 15
  double* cov = T + 224;     // up to 9 x 9
  double* o_rel = out.rel_out + (size_t)win * ISV_REL_REC;
  double* o_vb = out.vb_out + (size_t)win * ISV_VB_REC;
  double* o_rp = out.rp_out + (size_t)win * ISV_RP_REC;
  const double* dinv = sc + 40;
  if (lane == 0) {
    Quat Qi = quat_from_pose(sc), Qj = quat_from_pose(sc + 16);
    double dd[3] = {sc[16] - sc[0], sc[17] - sc[1], sc[18] - sc[2]};
    double tij[3], Rij[9];
    qrot(qinv(Qi), dd, tij);
    q2R(qmul(qinv(Qi), Qj), Rij);
    relpose_jacobians(sc, sc + 16, tij, Rij, Jrel, Jrel + 36, nullptr);
    for (int i = 0; i < 3; ++i) o_rel[i] = tij[i];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) o_rel[3 + r + 3 * c] = Rij[3 * r + c];
  } else if (lane == 1) {
    // RollPitchFactor(Qw) with Qw = Q_{V-1}: member R = Qw.toRotationMatrix()
    double Rm[9];
    q2R(quat_from_pose(sc), Rm);
    rollpitch_jacobian(sc, Rm, Jrp, nullptr);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) o_rp[r + 3 * c] = Rm[3 * r + c];
  } else if (lane >= 2 && lane < 11) {
    o_vb[lane - 2] = sc[23 + (lane - 2)];  // Linear9Factor(vb): VB = para_SpeedBias[V]
  }
  __syncwarp();
  // relative pose (rows 0-5 of Jr): Jj -> cols 0:6 (T_V), Ji -> cols 15:21 (T_{V-1})
  for (int idx = lane; idx < 6 * 15; idx += 32) {
    int r = idx % 6, k = idx / 6;
    double acc = 0.0;
    for (int c = 0; c < 6; ++c) {
      acc = fma(Jrel[36 + r + 6 * c], G[k + kMld * c], acc);
      acc = fma(Jrel[r + 6 * c], G[k + kMld * (15 + c)], acc);
    }
    JU[idx] = acc;
  }
  __syncwarp();
  for (int idx = lane; idx < 36; idx += 32) {
    int r = idx % 6, c = idx / 6;
    double acc = 0.0;
    for (int k = 0; k < 15; ++k) acc = fma(JU[r + 6 * k] * dinv[k], JU[c + 6 * k], acc);
    cov[idx] = acc;
  }
  __syncwarp();
  if (w_sqrt_info_from_cov<6>(cov, 6, o_rel + 12, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  // speed-bias prior (rows 6-14): J = I9 at cols 6:15
  for (int idx = lane; idx < 81; idx += 32) {
    int r = idx % 9, c = idx / 9;
    double acc = 0.0;
    for (int k = 0; k < 15; ++k) acc = fma(G[k + kMld * (6 + r)] * dinv[k], G[k + kMld * (6 + c)], acc);
    cov[idx] = acc;
  }
  __syncwarp();
  if (w_sqrt_info_from_cov<9>(cov, 9, o_vb + 9, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  // roll/pitch (rows 15-16): 2x6 at cols 15:21
  for (int idx = lane; idx < 2 * 15; idx += 32) {
    int r = idx % 2, k = idx / 2;
    double acc = 0.0;
    for (int c = 0; c < 6; ++c) acc = fma(Jrp[r + 2 * c], G[k + kMld * (15 + c)], acc);
    JU[idx] = acc;
  }
  __syncwarp();
  for (int idx = lane; idx < 4; idx += 32) {
    int r = idx % 2, c = idx / 2;
    double acc = 0.0;
    for (int k = 0; k < 15; ++k) acc = fma(JU[r + 2 * k] * dinv[k], JU[c + 2 * k], acc);
    cov[idx] = acc;
  }
  __syncwarp();
  if (w_sqrt_info_from_cov<2>(cov, 2, o_rp + 9, lane, nonfinite)) status |= ISV_W_NOT_SPD;
  if (__any_sync(kFullMask, nonfinite)) status |= ISV_W_NONFINITE;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) status |= __shfl_xor_sync(kFullMask, status, o);
  if (lane == 0) {
    out.rank[2 * win + 1] = rank;
    if (out.status) atomicOr(out.status + win, status);
  }
  (void)vo_size;
}

}  // namespace isv
