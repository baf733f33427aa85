// Device-resident estimator factor state for n independent sequences: the
// members Estimator keeps between frames -- vioRelativePoseEdges, vioPosePriorEdge, vioVBPrior,
// vioRollPitchEdges (include/estimator.h:137-154) and the pose-graph accumulator (CombinedFactors,
// include/factor/pose_graph_factors.h) -- live in HBM, so consecutive frames of a sequence never
// leave the GPU.  Kernels here restate, per sequence:
//
//   seq_update_kernel      factor->update(...) after the solve          src/estimator.cpp:1133-1144
//                          RelativePoseFactor::update  include/factor/relative_pose_factor.h:102-117
//                          SE3PriorFactor::update      include/factor/se3_prior_factor.h:73-81
//                          Linear9Factor::update       include/factor/linear9_factor.h:61-69
//                          RollPitchFactor::update     include/factor/rollpitch_factor.h:78-83
//   seq_yaw_kernel         double2vector()'s rotation of the priors     src/estimator.cpp:520-550
//   seq_rotate_kernel      slideWindow()'s factor rotation              src/estimator.cpp:1605-1638
//   seq_pg_kernel          CombinedFactors::operator+ and the keyframe cut (distance > 0.1)
//                          pose_graph_factors.h:27-51, src/pose_graph/pose_graph_builder.cpp:157-158,214
//
// Layout: edge-major [V][n][rec] so that "edge i of every sequence" is one contiguous [n][rec] array
// -- exactly what the window kernels' isv_batch_in.prior_* pointers expect.
#pragma once
#include "isv_device_math.cuh"
#include "isv_factors.cuh"
#include "isv_warp_linalg.cuh"

#include "../../include/isv_capi.h"

namespace isv {

struct SeqView {
  int n, V;
  double* rel;        // [V][n][48]   vioRelativePoseEdges[i]  (slot 0 unused)
  double* se3;        // [n][48]      vioPosePriorEdge
  double* vb;         // [n][90]      vioVBPrior
  double* rp;         // [V][n][13]   vioRollPitchEdges, slot = factor index
  int32_t* rp_valid;  // [V][n]
  double* rp_in;      // [n][5]       packed (valid, sqrt_info) of slot 0 for the forward kernel
  double* acc;        // [n][ISV_ACC_REC] accumFactor
  int32_t* pg_count;  // [n]          PoseGraphFactorCount (src/estimator.cpp:1274,1280)
  // outputs of the last MargForward / MargBackward
  double *se3_out, *pg_out, *rel_out, *vb_out, *rp_out;
  int32_t *rank, *status;
};

// Sophus::SO3d(Matrix3d).log() of  A^T * B  (row-major 3x3 in)
ISV_DI void log_At_B(const double* A, const double* B, double* lg) {
  double M[9];
  mat3_tmul(A, B, M);
  so3_log(R2q(M), lg);
}
// R (column-major record, 9) <- R * exp(w)
ISV_DI void right_mul_exp(double* Rcm, const double* w) {
  double R[9], E[9], O[9];
  load_mat3_colmajor(Rcm, R);
  q2R(so3_exp(w), E);
  mat3_mul(R, E, O);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rcm[r + 3 * c] = O[3 * r + c];
}

// RelativePoseFactor::update(ti, Ri, tj, Rj, PSi, PSj): rec = [delta_t 3 | delta_R 9 col-major | ...]
// ti/tj: old positions, Ri/Rj: old rotations (column-major 9), PSi/PSj: para_Pose after the solve.
ISV_DI void relpose_update(double* rec, const double* ti, const double* Ri_cm, const double* tj, const double* Rj_cm,
                           const double* PSi, const double* PSj) {
  double Ri[9], Rj[9], Qi_R[9], Qj_R[9];
  load_mat3_colmajor(Ri_cm, Ri);
  load_mat3_colmajor(Rj_cm, Rj);
  const Quat Qi = quat_from_pose(PSi), Qj = quat_from_pose(PSj);
  q2R(qinv(Qi), Qi_R);   // Qi.inverse() as a matrix
  q2R(qinv(Qj), Qj_R);
  double Mi[9], Mj[9], li[3], lj[3];
  mat3_mul(Qi_R, Ri, Mi);
  mat3_mul(Qj_R, Rj, Mj);
  so3_log(R2q(Mi), li);      // d_Ri.log()
  so3_log(R2q(Mj), lj);      // d_Rj.log()
  const double d_tj[3] = {PSj[0] - tj[0], PSj[1] - tj[1], PSj[2] - tj[2]};
  const double d_ti[3] = {PSi[0] - ti[0], PSi[1] - ti[1], PSi[2] - ti[2]};
  double a[3], b[3], S[9], c[3];
  mat3_tvec(Ri, d_tj, a);
  mat3_tvec(Ri, d_ti, b);
  skew3(rec, S);
  mat3_vec(S, li, c);
  double Ji[9], w[3];
  q2R(qmul(qinv(Qj), Qi), Ji);
  mat3_vec(Ji, li, w);
  for (int k = 0; k < 3; ++k) { rec[k] += a[k] - b[k] + c[k]; w[k] = -w[k]; }
  right_mul_exp(rec + 3, w);
  right_mul_exp(rec + 3, lj);
}

// SE3PriorFactor::update(Pi, Ri, PSi): rec = [t 3 | R 9 | ...]
ISV_DI void se3prior_update(double* rec, const double* Pi, const double* Ri_cm, const double* PS) {
  double R0[9], R1[9], lg[3];
  load_mat3_colmajor(Ri_cm, R0);
  // (R1.inverse() * R0).log() with R0 = SO3(Ri), R1 = SO3(Qi) (normalised)
  const Quat q0 = R2q(R0), q1 = qnormalized(quat_from_pose(PS));
  so3_log(so3_mul(qconj(q1), q0), lg);
  (void)R1;
  for (int k = 0; k < 3; ++k) rec[k] += PS[k] - Pi[k];
  right_mul_exp(rec + 3, lg);
}

// RollPitchFactor::update(Rs, Qs): rec = [R 9 | sqrt_info 4]
ISV_DI void rollpitch_update(double* rec, const double* Rs_cm, const double* PS) {
  double R0[9], lg[3];
  load_mat3_colmajor(Rs_cm, R0);
  const Quat q0 = R2q(R0), q1 = qnormalized(quat_from_pose(PS));
  so3_log(so3_mul(qconj(q1), q0), lg);
  right_mul_exp(rec, lg);
}

// one thread per (sequence, factor): f = 0 vb, 1 se3, 2..V rel[f-1], V+1.. rp slot f-V-1
__global__ void seq_update_kernel(SeqView s, const double* old_P, const double* old_R, const double* old_vb,
                                  const double* pose, const double* sb) {
  const int per = 2 + (s.V - 1) + s.V;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)s.n * per) return;
  const int q = (int)(t / per), f = (int)(t % per);
  const int V = s.V;
  const double* P = old_P + (size_t)q * V * 3;
  const double* R = old_R + (size_t)q * V * 9;
  const double* PS = pose + (size_t)q * V * 7;
  if (f == 0) {
    double* rec = s.vb + (size_t)q * ISV_VB_REC;
    for (int k = 0; k < 9; ++k) rec[k] += sb[(size_t)q * 9 + k] - old_vb[(size_t)q * 9 + k];
  } else if (f == 1) {
    se3prior_update(s.se3 + (size_t)q * ISV_SE3_REC, P, R, PS);
  } else if (f <= V) {
    const int j = f - 1, i = j - 1;   // vioRelativePoseEdges[j] links frames j-1 -> j
    relpose_update(s.rel + ((size_t)j * s.n + q) * ISV_REL_REC, P + 3 * i, R + 9 * i, P + 3 * j, R + 9 * j, PS + 7 * i,
                   PS + 7 * j);
  } else {
    const int idx = f - V - 1;
    if (s.rp_valid[(size_t)idx * s.n + q])
      rollpitch_update(s.rp + ((size_t)idx * s.n + q) * ISV_RP_REC, R + 9 * idx, PS + 7 * idx);
  }
}

// Utility::R2ypr / ypr2R (include/utility/utility.h:66-109), degrees
ISV_DI void R2ypr_deg(const double* R, double* ypr) {
  const double n0 = R[0], n1 = R[3], n2 = R[6], o0 = R[1], o1 = R[4], a0 = R[2], a1 = R[5];
  const double y = atan2(n1, n0);
  double sy, cy;
  sincos(y, &sy, &cy);
  const double p = atan2(-n2, n0 * cy + n1 * sy);
  const double r = atan2(a0 * sy - a1 * cy, -o0 * sy + o1 * cy);
  ypr[0] = y / kPi * 180.0; ypr[1] = p / kPi * 180.0; ypr[2] = r / kPi * 180.0;
}

// one thread per sequence: rot_diff from (Rs[0] before the solve, para_Pose[0] after it), applied
// to vioVBPrior->VB.tail<3>() and vioPosePriorEdge->R (:549-550); rot_out [n][9] column-major (may be null)
__global__ void seq_yaw_kernel(SeqView s, const double* old_R0, const double* pose0, double* rot_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= s.n) return;
  double Rs0[9], R00[9], y0[3], y00[3], rot[9];
  load_mat3_colmajor(old_R0 + (size_t)q * 9, Rs0);
  q2R(quat_from_pose(pose0 + (size_t)q * 7), R00);
  R2ypr_deg(Rs0, y0);
  R2ypr_deg(R00, y00);
  const double yd = (y0[0] - y00[0]) / 180.0 * kPi;
  double sn, cs;
  sincos(yd, &sn, &cs);
  rot[0] = cs; rot[1] = -sn; rot[2] = 0; rot[3] = sn; rot[4] = cs; rot[5] = 0; rot[6] = 0; rot[7] = 0; rot[8] = 1;
  if (fabs(fabs(y0[1]) - 90.0) < 1.0 || fabs(fabs(y00[1]) - 90.0) < 1.0) mat3_mult(Rs0, R00, rot);  // Rs[0] * R00^T
  double* vb = s.vb + (size_t)q * ISV_VB_REC;
  double v[3] = {vb[6], vb[7], vb[8]}, o[3];
  mat3_vec(rot, v, o);
  vb[6] = o[0]; vb[7] = o[1]; vb[8] = o[2];
  double* Rp = s.se3 + (size_t)q * ISV_SE3_REC + 3;
  double R[9], O[9];
  load_mat3_colmajor(Rp, R);
  mat3_mul(rot, R, O);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Rp[r + 3 * c] = O[3 * r + c];
      if (rot_out) rot_out[(size_t)q * 9 + r + 3 * c] = rot[3 * r + c];
    }
}

// slideWindow(): one thread per (sequence, element); slots move down in place (ascending i reads i+1)
__global__ void seq_rotate_kernel(SeqView s) {
  const int q = blockIdx.x;
  if (q >= s.n) return;
  const int V = s.V, n = s.n;
  for (int e = threadIdx.x; e < ISV_REL_REC; e += blockDim.x) {
    for (int i = 1; i < V - 1; ++i) s.rel[((size_t)i * n + q) * ISV_REL_REC + e] = s.rel[((size_t)(i + 1) * n + q) * ISV_REL_REC + e];
    s.rel[((size_t)(V - 1) * n + q) * ISV_REL_REC + e] = s.rel_out[(size_t)q * ISV_REL_REC + e];
    s.se3[(size_t)q * ISV_SE3_REC + e] = s.se3_out[(size_t)q * ISV_SE3_REC + e];
  }
  for (int e = threadIdx.x; e < ISV_VB_REC; e += blockDim.x) s.vb[(size_t)q * ISV_VB_REC + e] = s.vb_out[(size_t)q * ISV_VB_REC + e];
  // the new RollPitchFactor was pushed with index V-1 (:1516); every edge then shifts by one and
  // index < 0 is erased (:1619-1626)
  for (int e = threadIdx.x; e < ISV_RP_REC; e += blockDim.x) {
    for (int i = 0; i < V - 1; ++i) {
      const double v = (i + 1 < V - 1) ? s.rp[((size_t)(i + 1) * n + q) * ISV_RP_REC + e] : s.rp_out[(size_t)q * ISV_RP_REC + e];
      s.rp[((size_t)i * n + q) * ISV_RP_REC + e] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < V - 1; ++i) s.rp_valid[(size_t)i * n + q] = (i + 1 < V - 1) ? s.rp_valid[(size_t)(i + 1) * n + q] : 1;
    s.rp_valid[(size_t)(V - 1) * n + q] = 0;
    const int v0 = s.rp_valid[q];
    s.rp_in[(size_t)q * ISV_RP_IN_REC] = v0 ? 1.0 : 0.0;
    for (int k = 0; k < 4; ++k) s.rp_in[(size_t)q * ISV_RP_IN_REC + 1 + k] = v0 ? s.rp[(size_t)q * ISV_RP_REC + 9 + k] : 0.0;
  }
}

// install the initFactorGraph output (rel_out [n][V-1][48]) into the edge-major state, reset the rest
__global__ void seq_install_kernel(SeqView s, const double* rel_init) {
  const int q = blockIdx.x;
  if (q >= s.n) return;
  const int V = s.V, n = s.n;
  for (int e = threadIdx.x; e < ISV_REL_REC * (V - 1); e += blockDim.x) {
    const int i = e / ISV_REL_REC, k = e % ISV_REL_REC;
    s.rel[((size_t)(i + 1) * n + q) * ISV_REL_REC + k] = rel_init[((size_t)q * (V - 1) + i) * ISV_REL_REC + k];
  }
  for (int e = threadIdx.x; e < ISV_REL_REC; e += blockDim.x) s.rel[(size_t)q * ISV_REL_REC + e] = 0.0;
  for (int e = threadIdx.x; e < V; e += blockDim.x) s.rp_valid[(size_t)e * n + q] = 0;
  for (int e = threadIdx.x; e < V * ISV_RP_REC; e += blockDim.x) s.rp[((size_t)(e / ISV_RP_REC) * n + q) * ISV_RP_REC + e % ISV_RP_REC] = 0.0;
  for (int e = threadIdx.x; e < ISV_RP_IN_REC; e += blockDim.x) s.rp_in[(size_t)q * ISV_RP_IN_REC + e] = 0.0;
  for (int e = threadIdx.x; e < ISV_ACC_REC; e += blockDim.x) {
    // CombinedFactors(index = 0): delta_R = I, covRel = 0, vio_index = -1, length = 0 (pose_graph_factors.h:19-25)
    double v = 0.0;
    if (e == 3 || e == 7 || e == 11) v = 1.0;
    if (e == ISV_ACC_VIO_INDEX) v = -1.0;
    s.acc[(size_t)q * ISV_ACC_REC + e] = v;
  }
  if (threadIdx.x == 0) s.pg_count[q] = 0;
}

// -------------------------------------------------------------------------------------------------
// accumFactor = accumFactor + currentFactor ; emit + reset when accumFactor->distance > 0.1.
// One warp per sequence.  smem per warp: 5 * 36 + 72 doubles.
// -------------------------------------------------------------------------------------------------
constexpr int kPgSmemPerWarp = 5 * 36 + 72;

__global__ void seq_pg_kernel(SeqView s, const double* ts, const double* Ri, const double* ti, double cut_distance,
                              double* kf_out, int32_t* kf_flag) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= s.n) return;
  double* A = smem + warp * kPgSmemPerWarp;  // Adj (6x6)
  double* C1 = A + 36;                       // covRel1
  double* T = C1 + 36;                       // Adj * covRel1
  double* CR = T + 36;                       // covRel
  double* W = CR + 36;                       // scratch 6x6
  double* wk = W + 36;                       // 72
  double* acc = s.acc + (size_t)q * ISV_ACC_REC;
  const double* cur = s.pg_out + (size_t)q * ISV_PG_REC;
  int nonfinite = 0;
  // covRel1 = (s1^T s1)^-1   (:30)
  for (int idx = lane; idx < 36; idx += 32) {
    const int i = idx % 6, j = idx / 6;
    double v = 0.0;
    for (int l = 0; l < 6; ++l) v = fma(cur[12 + l + 6 * i], cur[12 + l + 6 * j], v);
    C1[idx] = v;
    CR[idx] = acc[ISV_ACC_COVREL + idx];
    A[idx] = 0.0;
  }
  __syncwarp();
  if (w_inverse(C1, 6, 6, wk, lane)) { /* singular: NaNs propagate as in the reference */ }
  // T0.Adj() = [R0, skew(t0) R0; 0, R0]   (Sophus SE3::Adj, tangent order (upsilon, omega))
  if (lane == 0) {
    double R0[9], S[9], SR[9];
    load_mat3_colmajor(acc + 3, R0);
    skew3(acc, S);
    mat3_mul(S, R0, SR);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        A[r + 6 * c] = R0[3 * r + c];
        A[(3 + r) + 6 * (3 + c)] = R0[3 * r + c];
        A[r + 6 * (3 + c)] = SR[3 * r + c];
      }
  }
  __syncwarp();
  w_gemm<false, false>(6, 6, 6, A, 6, C1, 6, T, 6, 0, lane);
  w_gemm<false, true>(6, 6, 6, T, 6, A, 6, CR, 6, 1, lane);       // covRel += Adj covRel1 Adj^T
  // T0 * T1 (Sophus: unit-quaternion product + first-order renormalisation; translation R0 t1 + t0)
  if (lane == 0) {
    double R0[9], R1[9], t01[3], R01[9];
    load_mat3_colmajor(acc + 3, R0);
    load_mat3_colmajor(cur + 3, R1);
    const Quat q0 = R2q(R0), q1 = R2q(R1);
    q2R(so3_mul(q0, q1), R01);
    qrot(q0, cur, t01);
    for (int k = 0; k < 3; ++k) { t01[k] += acc[k]; acc[k] = t01[k]; }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) acc[3 + r + 3 * c] = R01[3 * r + c];
    acc[ISV_ACC_DISTANCE] = sqrt(t01[0] * t01[0] + t01[1] * t01[1] + t01[2] * t01[2]);
    acc[ISV_ACC_LENGTH] += 1.0;
    // rollPitchFactor = other.rollPitchFactor (:32) -- the record travels with its covAbs
    const int rv = s.rp_in[(size_t)q * ISV_RP_IN_REC] != 0.0;   // vioRollPitchEdges[0]->index == 0 at MargForward time
    acc[ISV_ACC_RP_VALID] = rv ? 1.0 : 0.0;
    for (int k = 0; k < ISV_RP_REC; ++k) acc[ISV_ACC_RP + k] = rv ? s.rp[(size_t)q * ISV_RP_REC + k] : 0.0;
    for (int k = 0; k < 4; ++k) acc[ISV_ACC_COVABS + k] = cur[85 + k];
    if (acc[ISV_ACC_VIO_INDEX] == -1.0) {   // (:44-49)
      for (int k = 0; k < 3; ++k) acc[ISV_ACC_TI + k] = ti[(size_t)q * 3 + k];
      for (int k = 0; k < 9; ++k) acc[ISV_ACC_RI + k] = Ri[(size_t)q * 9 + k];
      acc[ISV_ACC_VIO_INDEX] = (double)s.pg_count[q];
      acc[ISV_ACC_TS] = ts[q];
    }
    s.pg_count[q] += 1;   // PoseGraphFactorCount++ (:1280)
  }
  for (int idx = lane; idx < 36; idx += 32) { acc[ISV_ACC_COVREL + idx] = CR[idx]; W[idx] = CR[idx]; }
  __syncwarp();
  // sqrt_info = LLT(covRel.inverse()).matrixL().transpose()   (:41)
  if (w_sqrt_info_from_cov_regs<6>(W, 6, acc + 12, lane, nonfinite)) { /* not SPD: flagged below */ nonfinite = 1; }
  __syncwarp();
  // keyframe cut (pose_graph_builder.cpp:158,214): emit a copy, start a fresh CombinedFactors(++pg_index)
  const bool emit = acc[ISV_ACC_DISTANCE] > cut_distance;
  if (kf_flag && lane == 0) kf_flag[q] = emit ? 1 : 0;
  if (emit) {
    const double next_pg = acc[ISV_ACC_PG_INDEX] + 1.0;
    __syncwarp();
    for (int e = lane; e < ISV_ACC_REC; e += 32) {
      if (kf_out) kf_out[(size_t)q * ISV_ACC_REC + e] = acc[e];
    }
    __syncwarp();
    for (int e = lane; e < ISV_ACC_REC; e += 32) {
      double v = 0.0;
      if (e == 3 || e == 7 || e == 11) v = 1.0;
      if (e == ISV_ACC_VIO_INDEX) v = -1.0;
      if (e == ISV_ACC_PG_INDEX) v = next_pg;
      acc[e] = v;
    }
  }
  if (__any_sync(kFullMask, nonfinite) && lane == 0 && s.status) atomicOr(s.status + q, ISV_W_NOT_SPD);
}

}  // namespace isv
