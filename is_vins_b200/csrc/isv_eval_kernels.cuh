// ceres `CostFunction::Evaluate` contract, batched (include/isv_capi.h "isv_eval_*"): weighted
// residuals and weighted ROW-MAJOR Jacobian blocks with a zero 7th pose column, for every factor
// problemSolve() adds to the ceres problem (/root/reference/src/estimator.cpp:1004-1146).
//
//   eval_projection_kernel  ProjectionFactor::Evaluate   src/factor/projection_factor.cpp:24-122
//   eval_imu_kernel         IMUFactor::Evaluate          include/factor/imu_factor.h:23-159
//   eval_small_kernel       RelativePoseFactor / SE3PriorFactor / Linear9Factor / RollPitchFactor /
//                           YawFactor ::Evaluate         include/factor/*.h
//
// The projection kernel is the HBM-bound one: 64 B in, up to 368 B out per factor and ~450 FP64
// instructions, so it is laid out for the memory system: component-major (SoA) index / observation
// reads (one coalesced 128/256-byte request per component per warp), parameter blocks gathered
// through L1/L2 (a window has 18 poses and ~10^3 features shared by ~10^4 factors), and every
// [n][14] Jacobian block transposed through shared memory so that a warp stores 32 consecutive
// doubles per instruction instead of 32 strided ones.
#pragma once
#include "isv_device_math.cuh"
#include "isv_factors.cuh"
#include "isv_warp_linalg.cuh"

#include "../../include/isv_capi.h"

namespace isv {

constexpr int kEvalWarps = 4;
constexpr int kEvalThreads = 32 * kEvalWarps;
constexpr int kStageLd = 15;  // odd stride: conflict-free 64-bit row writes

// ceres::CauchyLoss(a) + Corrector with rho'' <= 0: scale = sqrt(rho'(s)),  rho' = 1 / (1 + s / a^2)
ISV_DI double cauchy_scale(double a, double sq_norm) {
  if (!(a > 0.0)) return 1.0;
  return sqrt(1.0 / (1.0 + sq_norm / (a * a)));
}

// Each lane holds one factor's 2x7 block (v[14], row-major).  The warp's 32 blocks are contiguous
// in `dst` (dst -> block of lane 0); transpose through `stage` so global stores are coalesced.
ISV_DI void store_rows14(double* stage, int lane, const double* v, double* dst, int nvalid) {
#pragma unroll
  for (int c = 0; c < 14; ++c) stage[lane * kStageLd + c] = v[c];
  __syncwarp();
  const int total = nvalid * 14;
#pragma unroll
  for (int it = 0; it < 14; ++it) {
    const int i = it * 32 + lane;
    if (i < total) {
      const int f = i / 14, c = i - 14 * f;
      dst[i] = stage[f * kStageLd + c];
    }
  }
  __syncwarp();
}

// v * skew(p) for a row vector v: [v1 p2 - v2 p1, v2 p0 - v0 p2, v0 p1 - v1 p0]
ISV_DI void row_skew(const double* v, const double* p, double* o) {
  o[0] = v[1] * p[2] - v[2] * p[1];
  o[1] = v[2] * p[0] - v[0] * p[2];
  o[2] = v[0] * p[1] - v[1] * p[0];
}
// o (2x3) = red (2x3) * M (3x3 row-major)  /  red * M^T
ISV_DI void red_mul(const double* red, const double* M, double* o) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) o[3 * r + c] = red[3 * r] * M[c] + red[3 * r + 1] * M[3 + c] + red[3 * r + 2] * M[6 + c];
}
ISV_DI void red_mul_t(const double* red, const double* M, double* o) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      o[3 * r + c] = red[3 * r] * M[3 * c] + red[3 * r + 1] * M[3 * c + 1] + red[3 * r + 2] * M[3 * c + 2];
}

// Register plan: after the common part only the four points (pts_camera_i, pts_imu_i, pts_imu_j,
// pts_camera_j) and five 2x3 products  red, red*A, red*A*Ri, red*ric^T, red*tmp_r  stay live
// (A = ric^T Rj^T, tmp_r = A Ri ric); every Jacobian block is a cheap combination of those and is
// stored as soon as it is formed, so the kernel fits 128 registers -> 4 CTAs (16 warps) per SM.
// HAS_TD: ProjectionTdFactor (5th block, shifted observations); compiled separately so that the plain
// ProjectionFactor path keeps its register budget.
template <bool HAS_TD>
__global__ void __launch_bounds__(kEvalThreads, 4)
eval_projection_kernel(isv_param_blocks pb, isv_proj_factors fs, isv_proj_eval out, DevCfg cfg, int32_t* status) {
  __shared__ double stage_all[kEvalWarps][32 * kStageLd];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* stage = stage_all[warp];
  const long long wbase = ((long long)blockIdx.x * kEvalWarps + warp) * 32;
  if (wbase >= fs.n) return;
  const long long k = wbase + lane;
  const int nvalid = (int)((fs.n - wbase) < 32 ? (fs.n - wbase) : 32);
  const bool live = k < fs.n;
  bool ok = false;
  double pc_i[3], pim_i[3], pim_j[3], pc_j[3], red[6], rA[6], rARi[6], rT[6], rTmp[6];
  double res0 = 0.0, res1 = 0.0, jf0 = 0.0, jf1 = 0.0, jt0 = 0.0, jt1 = 0.0;
  if (live) {
    const int ii = fs.idx[k], jj = fs.idx[fs.stride + k], ie = fs.idx[2 * fs.stride + k], iff = fs.idx[3 * fs.stride + k];
    ok = !(ii < 0 || ii >= pb.n_pose || jj < 0 || jj >= pb.n_pose || ie < 0 || ie >= pb.n_ex_pose || iff < 0 ||
           iff >= pb.n_feature);
    if (!ok) {
      if (status) atomicOr(status, ISV_W_BAD_INDEX);
    } else {
      const double* PSi = pb.pose + (size_t)ii * 7;
      const double* PSj = pb.pose + (size_t)jj * 7;
      const double* PSe = pb.ex_pose + (size_t)ie * 7;
      const Quat Qi = quat_from_pose(PSi), Qj = quat_from_pose(PSj), qic = quat_from_pose(PSe);
      const double lam = pb.feature[iff];
      double pts_i[3] = {fs.obs[k], fs.obs[fs.stride + k], fs.obs[2 * fs.stride + k]};
      double xj = fs.obs[3 * fs.stride + k], yj = fs.obs[4 * fs.stride + k];
      double vix = 0.0, viy = 0.0, vjx = 0.0, vjy = 0.0;
      if (HAS_TD) {
        // ProjectionTdFactor: pts_td = pts - (td - td_obs + TR / ROW * row) * velocity   (velocity.z = 0)
        const int it = fs.td_idx ? fs.td_idx[k] : 0;
        if (it < 0 || it >= fs.n_td) {
          ok = false;
          if (status) atomicOr(status, ISV_W_BAD_INDEX);
        } else {
          const double td = fs.td[it];
          vix = fs.td_obs[k]; viy = fs.td_obs[fs.stride + k];
          vjx = fs.td_obs[2 * fs.stride + k]; vjy = fs.td_obs[3 * fs.stride + k];
          const double di = td - fs.td_obs[4 * fs.stride + k] + fs.tr_over_row * fs.td_obs[6 * fs.stride + k];
          const double dj = td - fs.td_obs[5 * fs.stride + k] + fs.tr_over_row * fs.td_obs[7 * fs.stride + k];
          pts_i[0] -= di * vix; pts_i[1] -= di * viy;
          xj -= dj * vjx; yj -= dj * vjy;
        }
      }
      // :38-42
      double pw[3], d[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) pc_i[a] = pts_i[a] / lam;
      qrot(qic, pc_i, pim_i);
#pragma unroll
      for (int a = 0; a < 3; ++a) pim_i[a] += PSe[a];
      qrot(Qi, pim_i, pw);
#pragma unroll
      for (int a = 0; a < 3; ++a) d[a] = pw[a] + PSi[a] - PSj[a];
      qrot(qinv(Qj), d, pim_j);
#pragma unroll
      for (int a = 0; a < 3; ++a) d[a] = pim_j[a] - PSe[a];
      qrot(qinv(qic), d, pc_j);
      const double dep_j = pc_j[2];
      const double r0 = pc_j[0] / dep_j - xj, r1 = pc_j[1] / dep_j - yj;     // :48-49
      const double s00 = cfg.ps[0], s10 = cfg.ps[1], s01 = cfg.ps[2], s11 = cfg.ps[3];
      res0 = s00 * r0 + s01 * r1;                                            // :52
      res1 = s10 * r0 + s11 * r1;
      const double ls = cauchy_scale(fs.cauchy_a, res0 * res0 + res1 * res1);
      res0 *= ls;
      res1 *= ls;
      // reduce = sqrt_info * [1/z 0 -x/z^2 ; 0 1/z -y/z^2]   (:72-75), loss scale folded in
      const double iz = 1.0 / dep_j, iz2 = 1.0 / (dep_j * dep_j);
      const double u02 = -pc_j[0] * iz2, u12 = -pc_j[1] * iz2;
      red[0] = ls * s00 * iz;  red[1] = ls * s01 * iz;  red[2] = ls * (s00 * u02 + s01 * u12);
      red[3] = ls * s10 * iz;  red[4] = ls * s11 * iz;  red[5] = ls * (s10 * u02 + s11 * u12);
      {
        double Ri[9], Rj[9], ric[9], A[9], T[9];
        q2R(Qj, Rj);
        q2R(qic, ric);
        mat3_mul(Rj, ric, T);                       // A = ric^T Rj^T = (Rj ric)^T
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) A[3 * r + c] = T[3 * c + r];
        red_mul(red, A, rA);
        red_mul_t(red, ric, rT);
        q2R(Qi, Ri);
        mat3_mul(A, Ri, T);                         // A Ri
        red_mul(red, T, rARi);
        mat3_mul(T, ric, A);                        // tmp_r = A Ri ric   (:104)
        red_mul(red, A, rTmp);
      }
      // :116  reduce * tmp_r * pts_i * -1 / lam^2
      const double sc = -1.0 / (lam * lam);
      jf0 = (rTmp[0] * pts_i[0] + rTmp[1] * pts_i[1] + rTmp[2] * pts_i[2]) * sc;
      jf1 = (rTmp[3] * pts_i[0] + rTmp[4] * pts_i[1] + rTmp[5] * pts_i[2]) * sc;
      // jacobian_td = reduce * tmp_r * velocity_i / inv_dep * -1 + sqrt_info * velocity_j.head(2)
      if (HAS_TD) {
        jt0 = -(rTmp[0] * vix + rTmp[1] * viy) / lam + ls * (s00 * vjx + s01 * vjy);
        jt1 = -(rTmp[3] * vix + rTmp[4] * viy) / lam + ls * (s10 * vjx + s11 * vjy);
      }
    }
  }
  if (live && out.residuals) reinterpret_cast<double2*>(out.residuals)[k] = make_double2(res0, res1);
  if (live && out.jac_feature) reinterpret_cast<double2*>(out.jac_feature)[k] = make_double2(jf0, jf1);
  if (HAS_TD && live && out.jac_td) reinterpret_cast<double2*>(out.jac_td)[k] = make_double2(jt0, jt1);
  double v[14];
  if (out.jac_pose_i) {       // :80-85  jaco_i = [A | A Ri (-skew(pts_imu_i))]
#pragma unroll
    for (int c = 0; c < 14; ++c) v[c] = 0.0;
    if (ok) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double t[3];
        row_skew(rARi + 3 * r, pim_i, t);
#pragma unroll
        for (int c = 0; c < 3; ++c) { v[7 * r + c] = rA[3 * r + c]; v[7 * r + 3 + c] = -t[c]; }
      }
    }
    store_rows14(stage, lane, v, out.jac_pose_i + wbase * 14, nvalid);
  }
  if (out.jac_pose_j) {       // :93-98  jaco_j = [-A | ric^T skew(pts_imu_j)]
#pragma unroll
    for (int c = 0; c < 14; ++c) v[c] = 0.0;
    if (ok) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double t[3];
        row_skew(rT + 3 * r, pim_j, t);
#pragma unroll
        for (int c = 0; c < 3; ++c) { v[7 * r + c] = -rA[3 * r + c]; v[7 * r + 3 + c] = t[c]; }
      }
    }
    store_rows14(stage, lane, v, out.jac_pose_j + wbase * 14, nvalid);
  }
  if (out.jac_ex_pose) {
    // :103-111  jaco_ex = [ric^T (Rj^T Ri - I) | -tmp_r skew(pc_i) + skew(tmp_r pc_i) + skew(w2)],
    //   w2 = ric^T (Rj^T (Ri tic + Pi - Pj) - tic).  ric^T (Rj^T Ri - I) = A Ri - ric^T, and because
    //   pts_imu_j = Rj^T (Ri (ric pc_i + tic) + Pi - Pj):  tmp_r pc_i + w2 = ric^T (pts_imu_j - tic)
    //   = pts_camera_j, so the two skew terms collapse to skew(pts_camera_j).
#pragma unroll
    for (int c = 0; c < 14; ++c) v[c] = 0.0;
    if (ok) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double t1[3], t2[3];
        row_skew(rTmp + 3 * r, pc_i, t1);
        row_skew(red + 3 * r, pc_j, t2);
#pragma unroll
        for (int c = 0; c < 3; ++c) { v[7 * r + c] = rARi[3 * r + c] - rT[3 * r + c]; v[7 * r + 3 + c] = t2[c] - t1[c]; }
      }
    }
    store_rows14(stage, lane, v, out.jac_ex_pose + wbase * 14, nvalid);
  }
}

// -------------------------------------------------------------------------------------------------
// IMUFactor::Evaluate: one warp per factor.
// smem per warp (doubles): P[225] S[225] J[15*30] W[15*30] r[16] sc[48]
// -------------------------------------------------------------------------------------------------
constexpr int kImuEvalSmem = 225 + 225 + 450 + 450 + 16 + 48;

__global__ void __launch_bounds__(kEvalThreads, 4)
eval_imu_kernel(isv_param_blocks pb, isv_imu_factors fs, isv_imu_eval out, DevCfg cfg, int32_t* status) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = blockIdx.x * kEvalWarps + warp;
  if (f >= fs.n) return;
  double* P = smem + warp * kImuEvalSmem;
  double* S = P + 225;
  double* J = S + 225;   // 15 x 30, ld 15: [T_i 6 | VB_i 9 | T_j 6 | VB_j 9]
  double* W = J + 450;
  double* r = W + 450;
  double* sc = r + 16;
  const int ii = fs.idx[2 * f], jj = fs.idx[2 * f + 1];
  if (ii < 0 || jj < 0 || ii >= pb.n_pose || jj >= pb.n_pose || ii >= pb.n_speed_bias || jj >= pb.n_speed_bias) {
    if (lane == 0 && status) atomicOr(status, ISV_W_BAD_INDEX);
    return;
  }
  const double* pre = fs.preint + (size_t)f * ISV_PREINT_REC;
  for (int i = lane; i < 225; i += 32) P[i] = pre[17 + 225 + i];
  for (int i = lane; i < 450; i += 32) J[i] = 0.0;
  if (lane < 7) { sc[lane] = pb.pose[(size_t)ii * 7 + lane]; sc[16 + lane] = pb.pose[(size_t)jj * 7 + lane]; }
  if (lane >= 7 && lane < 16) {
    sc[lane] = pb.speed_bias[(size_t)ii * 9 + lane - 7];
    sc[16 + lane] = pb.speed_bias[(size_t)jj * 9 + lane - 7];
  }
  if (lane < 3) sc[32 + lane] = cfg.g[lane];
  __syncwarp();
  imu_jacobians(sc, sc + 7, sc + 16, sc + 23, pre, sc + 32, J, 15, 0, 6, 15, 21, r, lane, 32);
  __syncwarp();
  // sqrt_info = LLT(covariance.inverse()).matrixL().transpose()   (imu_factor.h:44)
  int nonfinite = 0;
  int st = 0;
  if (w_sqrt_info_from_cov<15>(P, 15, S, lane, nonfinite)) st |= ISV_W_NOT_SPD;
  // residuals = sqrt_info * r (:45), jacobians = sqrt_info * J (:76-157); S is upper triangular
  if (lane < 15) {
    double acc = 0.0;
    for (int l = lane; l < 15; ++l) acc = fma(S[lane + 15 * l], r[l], acc);
    if (!isfinite(acc)) nonfinite = 1;
    out.residuals[(size_t)f * 15 + lane] = acc;
  }
  if (out.jacobians) {
    for (int idx = lane; idx < 450; idx += 32) {
      const int i = idx % 15, c = idx / 15;
      double acc = 0.0;
      for (int l = i; l < 15; ++l) acc = fma(S[i + 15 * l], J[l + 15 * c], acc);
      W[idx] = acc;
    }
    __syncwarp();
    double* o = out.jacobians + (size_t)f * ISV_IMU_JAC_REC;
    for (int idx = lane; idx < ISV_IMU_JAC_REC; idx += 32) {
      int rem, w, c0;
      if (idx < 105) { rem = idx; w = 7; c0 = 0; }
      else if (idx < 240) { rem = idx - 105; w = 9; c0 = 6; }
      else if (idx < 345) { rem = idx - 240; w = 7; c0 = 15; }
      else { rem = idx - 345; w = 9; c0 = 21; }
      const int row = rem / w, col = rem - w * row;
      const double v = (w == 7 && col == 6) ? 0.0 : W[row + 15 * (c0 + col)];
      o[idx] = v;
    }
  }
  if (__any_sync(kFullMask, nonfinite)) st |= ISV_W_NONFINITE;
  if (lane == 0 && st && status) atomicOr(status, st);
}

// -------------------------------------------------------------------------------------------------
// The recovered / prior factors: one thread per factor, all five types in one launch.
// -------------------------------------------------------------------------------------------------
// out (rows x 7 row-major) = scale * s (rows x rows col-major) * Jt (rows x 6 col-major, ld rows); 7th col 0
template <int R>
ISV_DI void weighted_rows7(const double* __restrict__ s, const double* __restrict__ Jt, double scale,
                           double* __restrict__ o) {
  for (int i = 0; i < R; ++i) {
    for (int c = 0; c < 6; ++c) {
      double acc = 0.0;
      for (int l = 0; l < R; ++l) acc = fma(s[i + R * l], Jt[l + R * c], acc);
      o[7 * i + c] = scale * acc;
    }
    o[7 * i + 6] = 0.0;
  }
}
template <int R>
ISV_DI double weighted_res(const double* __restrict__ s, const double* __restrict__ r, double* __restrict__ o) {
  double n2 = 0.0;
  for (int i = 0; i < R; ++i) {
    double acc = 0.0;
    for (int l = 0; l < R; ++l) acc = fma(s[i + R * l], r[l], acc);
    o[i] = acc;
    n2 = fma(acc, acc, n2);
  }
  return n2;
}

__global__ void __launch_bounds__(kEvalThreads)
eval_small_kernel(isv_param_blocks pb, isv_small_factors fs, isv_small_eval out, int32_t* status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  auto bad = [&]() { if (status) atomicOr(status, ISV_W_BAD_INDEX); };
  if (t < fs.n_rel) {
    const int ii = fs.rel_idx[2 * t], jj = fs.rel_idx[2 * t + 1];
    if (ii < 0 || jj < 0 || ii >= pb.n_pose || jj >= pb.n_pose) return bad();
    const double* rec = fs.rel_rec + (size_t)t * ISV_REL_REC;
    double dR[9], Ji[36], Jj[36], r[6], wr[6];
    load_mat3_colmajor(rec + 3, dR);
    relpose_jacobians(pb.pose + (size_t)ii * 7, pb.pose + (size_t)jj * 7, rec, dR, Ji, Jj, r);
    const double ls = cauchy_scale(fs.cauchy_a, weighted_res<6>(rec + 12, r, wr));
    for (int i = 0; i < 6; ++i) out.rel_res[(size_t)t * 6 + i] = ls * wr[i];
    if (out.rel_jac) {
      weighted_rows7<6>(rec + 12, Ji, ls, out.rel_jac + (size_t)t * 84);
      weighted_rows7<6>(rec + 12, Jj, ls, out.rel_jac + (size_t)t * 84 + 42);
    }
    return;
  }
  t -= fs.n_rel;
  if (t < fs.n_se3) {
    const int ii = fs.se3_idx[t];
    if (ii < 0 || ii >= pb.n_pose) return bad();
    const double* rec = fs.se3_rec + (size_t)t * ISV_SE3_REC;
    double Rp[9], J[36], r[6], wr[6];
    load_mat3_colmajor(rec + 3, Rp);
    se3prior_jacobian(pb.pose + (size_t)ii * 7, rec, Rp, J, r);
    const double ls = cauchy_scale(fs.cauchy_a, weighted_res<6>(rec + 12, r, wr));
    for (int i = 0; i < 6; ++i) out.se3_res[(size_t)t * 6 + i] = ls * wr[i];
    if (out.se3_jac) weighted_rows7<6>(rec + 12, J, ls, out.se3_jac + (size_t)t * 42);
    return;
  }
  t -= fs.n_se3;
  if (t < fs.n_vb) {
    const int ii = fs.vb_idx[t];
    if (ii < 0 || ii >= pb.n_speed_bias) return bad();
    const double* rec = fs.vb_rec + (size_t)t * ISV_VB_REC;
    const double* vb = pb.speed_bias + (size_t)ii * 9;
    double r[9], wr[9];
    for (int i = 0; i < 9; ++i) r[i] = vb[i] - rec[i];
    const double ls = cauchy_scale(fs.cauchy_a, weighted_res<9>(rec + 9, r, wr));
    for (int i = 0; i < 9; ++i) out.vb_res[(size_t)t * 9 + i] = ls * wr[i];
    if (out.vb_jac)
      for (int i = 0; i < 9; ++i)
        for (int c = 0; c < 9; ++c) out.vb_jac[(size_t)t * 81 + 9 * i + c] = ls * rec[9 + i + 9 * c];
    return;
  }
  t -= fs.n_vb;
  if (t < fs.n_rp) {
    const int ii = fs.rp_idx[t];
    if (ii < 0 || ii >= pb.n_pose) return bad();
    const double* rec = fs.rp_rec + (size_t)t * ISV_RP_REC;
    double Rm[9], J[12], r[2], wr[2];
    load_mat3_colmajor(rec, Rm);
    rollpitch_jacobian(pb.pose + (size_t)ii * 7, Rm, J, r);
    const double ls = cauchy_scale(fs.cauchy_a, weighted_res<2>(rec + 9, r, wr));
    for (int i = 0; i < 2; ++i) out.rp_res[(size_t)t * 2 + i] = ls * wr[i];
    if (out.rp_jac) weighted_rows7<2>(rec + 9, J, ls, out.rp_jac + (size_t)t * 14);
    return;
  }
  t -= fs.n_rp;
  if (t < fs.n_yaw) {
    const int ii = fs.yaw_idx[t];
    if (ii < 0 || ii >= pb.n_pose) return bad();
    const double* rec = fs.yaw_rec + (size_t)t * ISV_YAW_REC;
    double J[6], r[1], wr[1];
    yaw_jacobian(pb.pose + (size_t)ii * 7, rec, J, r);
    const double ls = cauchy_scale(fs.cauchy_a, weighted_res<1>(rec + 3, r, wr));
    out.yaw_res[t] = ls * wr[0];
    if (out.yaw_jac) weighted_rows7<1>(rec + 3, J, ls, out.yaw_jac + (size_t)t * 7);
  }
}

// PoseLocalParameterization::Plus  src/factor/pose_local_parameterization.cpp:3-19 -- the manifold update ceres applies
// to every pose block after a step (and the convention every Jacobian above assumes: right-multiplicative
// q * deltaQ(dtheta), deltaQ = (1, dtheta / 2) NOT normalised, utility.h:11-24, then normalised).  Thread per block.
__global__ void __launch_bounds__(128)
pose_plus_kernel(long long n, const double* __restrict__ x, const double* __restrict__ delta, double* __restrict__ out) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double* xs = x + 7 * k;
  const double* dl = delta + 6 * k;
  double* o = out + 7 * k;
  const Quat q = quat_from_pose(xs);
  const Quat dq = Quat{1.0, dl[3] / 2.0, dl[4] / 2.0, dl[5] / 2.0};
  const Quat r = qnormalized(qmul(q, dq));
  o[0] = xs[0] + dl[0]; o[1] = xs[1] + dl[1]; o[2] = xs[2] + dl[2];
  o[3] = r.x; o[4] = r.y; o[5] = r.z; o[6] = r.w;
}

}  // namespace isv
