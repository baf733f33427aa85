// Tangent-space ("EvaluateOnlyJacobians") twins of the reference's factors, as scalar device
// functions writing COLUMN-MAJOR blocks to (shared) memory.  Each is executed by one lane; the
// dense algebra that consumes the blocks is warp-parallel (isv_warp_linalg.cuh).
//
//   relpose_jacobians   include/factor/relative_pose_factor.h:72-101
//   se3prior_jacobian   include/factor/se3_prior_factor.h:53-71
//   rollpitch_jacobian  include/factor/rollpitch_factor.h:59-76
//   yaw_jacobian        include/factor/yaw_factor.h:51-65
//   imu_jacobians       include/factor/imu_factor.h:161-265 (+ integration_base.h:160-186)
#pragma once
#include "isv_device_math.cuh"

namespace isv {

// Ji, Jj: 6x6 column-major (ld 6).  dR: row-major delta_R.  res: 6 (may be nullptr).
__device__ inline void relpose_jacobians(const double* PSi, const double* PSj, const double* dt, const double* dR,
                                         double* Ji, double* Jj, double* res) {
  Quat Qi = quat_from_pose(PSi), Qj = quat_from_pose(PSj);
  double Ri[9], Rj[9];
  q2R(Qi, Ri);
  q2R(Qj, Rj);
  double d[3] = {PSj[0] - PSi[0], PSj[1] - PSi[1], PSj[2] - PSi[2]};
  double tij[3];
  qrot(qinv(Qi), d, tij);
  // res_R = SO3(delta_R * Rj^T * Ri)
  double A[9], B[9];
  mat3_mult(dR, Rj, A);  // dR * Rj^T
  mat3_mul(A, Ri, B);
  Quat qr = R2q(B);
  double lg[3];
  so3_log(qr, lg);
  if (res) {
    res[0] = dt[0] - tij[0]; res[1] = dt[1] - tij[1]; res[2] = dt[2] - tij[2];
    res[3] = lg[0]; res[4] = lg[1]; res[5] = lg[2];
  }
  double J[9];
  so3_right_jacobian_inv(lg, J);
  for (int i = 0; i < 36; ++i) { Ji[i] = 0.0; Jj[i] = 0.0; }
  double S[9];
  skew3(tij, S);
  double RitRj[9], JR[9];
  mat3_tmul(Ri, Rj, RitRj);   // Ri^T Rj
  mat3_mul(J, RitRj, JR);     // J Ri^T Rj
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      Ji[r + 6 * c] = Ri[3 * c + r];              // Ri^T
      Ji[r + 6 * (3 + c)] = -S[3 * r + c];        // -skew(tij)
      Ji[(3 + r) + 6 * (3 + c)] = J[3 * r + c];
      Jj[r + 6 * c] = -Ri[3 * c + r];             // -Ri^T
      Jj[(3 + r) + 6 * (3 + c)] = -JR[3 * r + c];
    }
}

// J: 6x6 column-major = blkdiag(I3, Jr^-1(log(R_prior^T R)))
__device__ inline void se3prior_jacobian(const double* PS, const double* t, const double* Rprior_rowmajor, double* J,
                                         double* res) {
  Quat ri = qnormalized(quat_from_pose(PS));
  Quat rp = R2q(Rprior_rowmajor);
  Quat rr = so3_mul(qconj(rp), ri);
  double lg[3];
  so3_log(rr, lg);
  if (res) {
    res[0] = PS[0] - t[0]; res[1] = PS[1] - t[1]; res[2] = PS[2] - t[2];
    res[3] = lg[0]; res[4] = lg[1]; res[5] = lg[2];
  }
  double Jr[9];
  so3_right_jacobian_inv(lg, Jr);
  for (int i = 0; i < 36; ++i) J[i] = 0.0;
  J[0] = 1.0; J[7] = 1.0; J[14] = 1.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) J[(3 + r) + 6 * (3 + c)] = Jr[3 * r + c];
}

// J: 2x6 column-major (ld 2); Rmeas row-major
__device__ inline void rollpitch_jacobian(const double* PS, const double* Rmeas_rowmajor, double* J, double* res) {
  Quat ri = qnormalized(quat_from_pose(PS));
  Quat rm = R2q(Rmeas_rowmajor);
  double nz[3] = {0.0, 0.0, -1.0};
  double a[3], r3[3];
  qrot(qconj(ri), nz, a);
  qrot(rm, a, r3);
  if (res) { res[0] = r3[0]; res[1] = r3[1]; }
  double S[9], Rm[9], SR[9];
  skew3(r3, S);
  q2R(rm, Rm);
  mat3_mul(S, Rm, SR);
  for (int i = 0; i < 12; ++i) J[i] = 0.0;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c) J[r + 2 * (3 + c)] = SR[3 * r + c];
}

// J: 1x6 ; yaw_meas = Qw.inverse() * UnitX
__device__ inline void yaw_jacobian(const double* PS, const double* yaw_meas, double* J, double* res = nullptr) {
  Quat ri = qnormalized(quat_from_pose(PS));
  double R[9], S[9], RS[9];
  q2R(ri, R);
  if (res) {  // residual = (Ri * yaw_meas).y   (yaw_factor.h:31-33)
    double v[3];
    qrot(ri, yaw_meas, v);
    res[0] = v[1];
  }
  skew3(yaw_meas, S);
  mat3_mul(R, S, RS);
  for (int i = 0; i < 6; ++i) J[i] = 0.0;
  for (int c = 0; c < 3; ++c) J[3 + c] = -RS[3 + c];
}

// The (3x3, lower-right) block of Utility::Qleft(q) / Qright(q), row-major.
__device__ inline void qleft_br(const Quat& q, double* M) {
  M[0] = q.w;  M[1] = -q.z; M[2] = q.y;
  M[3] = q.z;  M[4] = q.w;  M[5] = -q.x;
  M[6] = -q.y; M[7] = q.x;  M[8] = q.w;
}
// lower-right 3x3 of Qleft(a) * Qright(b), row-major
__device__ inline void qleft_qright_br(const Quat& a, const Quat& b, double* M) {
  // full 4x4 product restricted to rows/cols 1..3: sum over k=0..3 of L[r][k] * Rm[k][c]
  double L[16], Rm[16];
  L[0] = a.w; L[1] = -a.x; L[2] = -a.y; L[3] = -a.z;
  L[4] = a.x; L[5] = a.w;  L[6] = -a.z; L[7] = a.y;
  L[8] = a.y; L[9] = a.z;  L[10] = a.w; L[11] = -a.x;
  L[12] = a.z; L[13] = -a.y; L[14] = a.x; L[15] = a.w;
  Rm[0] = b.w; Rm[1] = -b.x; Rm[2] = -b.y; Rm[3] = -b.z;
  Rm[4] = b.x; Rm[5] = b.w;  Rm[6] = b.z;  Rm[7] = -b.y;
  Rm[8] = b.y; Rm[9] = -b.z; Rm[10] = b.w; Rm[11] = b.x;
  Rm[12] = b.z; Rm[13] = b.y; Rm[14] = -b.x; Rm[15] = b.w;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += L[4 * (r + 1) + k] * Rm[4 * k + (c + 1)];
      M[3 * r + c] = s;
    }
}

// Unweighted IMU Jacobians in WINDOW column order.  Jfull: 15 x 30, column-major (ld ldj) by default, columns
// [col_j_pose .. +6) <- d r / d T_j, [col_j_sb .. +9) <- d r / d VB_j, likewise for i.
// pre: the 467-double pre-integration record (include/isv_capi.h ISV_PREINT_REC).  res: 15 or null.
// Cooperative form: called by `nparts` lanes with part = 0..nparts-1; every lane evaluates the short
// scalar prologue (broadcast loads), the nine 3x3 sub-block positions (r, c) are dealt round-robin
// and part 0 writes the residual.  part = 0, nparts = 1 is the single-lane form.
__device__ inline void imu_jacobians(const double* PSi, const double* VBi, const double* PSj, const double* VBj,
                                     const double* __restrict__ pre, const double* G, double* Jfull, int ldj,
                                     int col_i_pose, int col_i_sb, int col_j_pose, int col_j_sb, double* res,
                                     int part = 0, int nparts = 1, int row_stride = 1) {
  const double* delta_p = pre;
  Quat dq{pre[6], pre[3], pre[4], pre[5]};
  const double* delta_v = pre + 7;
  const double* lin_ba = pre + 10;
  const double* lin_bg = pre + 13;
  const double sum_dt = pre[16];
  const double* Jp = pre + 17;  // 15x15 column-major
  auto JB = [&](int r0, int c0, int r, int c) { return Jp[(r0 + r) + 15 * (c0 + c)]; };
  Quat Qi = quat_from_pose(PSi), Qj = quat_from_pose(PSj);
  Quat Qi_inv = qinv(Qi);
  double Ri_inv[9];
  q2R(Qi_inv, Ri_inv);
  double dbg[3] = {VBi[6] - lin_bg[0], VBi[7] - lin_bg[1], VBi[8] - lin_bg[2]};
  double dba[3] = {VBi[3] - lin_ba[0], VBi[4] - lin_ba[1], VBi[5] - lin_ba[2]};
  double th[3];
  for (int r = 0; r < 3; ++r) th[r] = JB(3, 12, r, 0) * dbg[0] + JB(3, 12, r, 1) * dbg[1] + JB(3, 12, r, 2) * dbg[2];
  Quat cq = qmul(dq, Quat{1.0, th[0] / 2.0, th[1] / 2.0, th[2] / 2.0});
  double s = sum_dt;
  double a1[3], a2[3], v1[3], v2[3];
  for (int k = 0; k < 3; ++k) {
    a1[k] = 0.5 * G[k] * s * s + PSj[k] - PSi[k] - VBi[k] * s;
    a2[k] = G[k] * s + VBj[k] - VBi[k];
  }
  qrot(Qi_inv, a1, v1);
  qrot(Qi_inv, a2, v2);
  if (res && part == 0) {
    Quat e = qmul(qinv(cq), qmul(Qi_inv, Qj));
    double ev[3] = {e.x, e.y, e.z};
    for (int r = 0; r < 3; ++r) {
      double cp = delta_p[r], cv = delta_v[r];
      for (int c = 0; c < 3; ++c) {
        cp += JB(0, 9, r, c) * dba[c] + JB(0, 12, r, c) * dbg[c];
        cv += JB(6, 9, r, c) * dba[c] + JB(6, 12, r, c) * dbg[c];
      }
      res[r] = v1[r] - cp;
      res[3 + r] = 2.0 * ev[r];
      res[6 + r] = v2[r] - cv;
      res[9 + r] = VBj[3 + r] - VBi[3 + r];
      res[12 + r] = VBj[6 + r] - VBi[6 + r];
    }
  }
  // Jfull must be zero-filled by the caller (the warp does it cooperatively)
  double S1[9], S2[9], QLR[9], QL1[9], QL2[9];
  skew3(v1, S1);
  skew3(v2, S2);
  Quat QjinvQi = qmul(qinv(Qj), Qi);
  qleft_qright_br(QjinvQi, cq, QLR);
  qleft_br(qmul(QjinvQi, dq), QL1);                       // Q9: uncorrected delta_q
  qleft_br(qmul(qmul(qinv(cq), Qi_inv), Qj), QL2);
  // element (row, col) at Jfull[row * row_stride + col * ldj]: (1, ld) = column-major, (ld, 1) = row-major
  auto put = [&](int row, int col, double v) { Jfull[row * row_stride + ldj * col] = v; };
  for (int rc = part; rc < 9; rc += nparts) {
      const int r = rc / 3, c = rc - 3 * r;
      // d / d T_i   (15x6)
      put(0 + r, col_i_pose + c, -Ri_inv[3 * r + c]);
      put(0 + r, col_i_pose + 3 + c, S1[3 * r + c]);
      put(3 + r, col_i_pose + 3 + c, -QLR[3 * r + c]);
      put(6 + r, col_i_pose + 3 + c, S2[3 * r + c]);
      // d / d VB_i  (15x9)
      put(0 + r, col_i_sb + c, -Ri_inv[3 * r + c] * s);
      put(0 + r, col_i_sb + 3 + c, -JB(0, 9, r, c));
      put(0 + r, col_i_sb + 6 + c, -JB(0, 12, r, c));
      double qd = QL1[3 * r] * JB(3, 12, 0, c) + QL1[3 * r + 1] * JB(3, 12, 1, c) + QL1[3 * r + 2] * JB(3, 12, 2, c);
      put(3 + r, col_i_sb + 6 + c, -qd);
      put(6 + r, col_i_sb + c, -Ri_inv[3 * r + c]);
      put(6 + r, col_i_sb + 3 + c, -JB(6, 9, r, c));
      put(6 + r, col_i_sb + 6 + c, -JB(6, 12, r, c));
      put(9 + r, col_i_sb + 3 + c, r == c ? -1.0 : 0.0);
      put(12 + r, col_i_sb + 6 + c, r == c ? -1.0 : 0.0);
      // d / d T_j   (15x6)
      put(0 + r, col_j_pose + c, Ri_inv[3 * r + c]);
      put(3 + r, col_j_pose + 3 + c, QL2[3 * r + c]);
      // d / d VB_j  (15x9)
      put(6 + r, col_j_sb + c, Ri_inv[3 * r + c]);
      put(9 + r, col_j_sb + 3 + c, r == c ? 1.0 : 0.0);
      put(12 + r, col_j_sb + 6 + c, r == c ? 1.0 : 0.0);
    }
}

}  // namespace isv
