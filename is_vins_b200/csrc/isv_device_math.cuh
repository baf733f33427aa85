// Small fixed-size FP64 geometry used by every kernel: Eigen-semantics quaternions, Sophus-semantics
// SO(3) log/exp, the right-Jacobian inverse, skew products.  3x3 matrices are ROW-MAJOR double[9]
// inside kernels (m[3*r+c]); conversion to the column-major C-ABI records happens at load/store.
//
// Reference semantics followed (paths relative to /root/reference):
//   Eigen quaternion ops (un-vendored Eigen 3.3.4): q*v formula, inverse = conj/|q|^2,
//   toRotationMatrix without normalisation, matrix->quaternion (Shepperd)      SURVEY.md Q13,Q14
//   include/utility/utility.h:11-64            deltaQ, skewSymmetric, Qleft, Qright
//   include/utility/sophus_utils.hpp:194-236   rightJacobianInvSO3
//   Sophus SO3::log / exp / ctor (un-vendored)                                 SURVEY.md section 9
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace isv {

#define ISV_DI __device__ __forceinline__

// the globals the reference reads inside the hot path (isv_config), passed by value to kernels
struct DevCfg {
  double alpha;
  double ps[4];  // ProjectionFactor::sqrt_info, column-major 2x2
  double g[3];
  double qr_threshold;
};

constexpr double kSophusEps = 1e-10;
constexpr double kSophusEpsSqrt = 1e-5;
constexpr double kPi = 3.14159265358979323846;

struct Quat {
  double w, x, y, z;
};

// Lean reciprocal square root / reciprocal for the hot chains of the window kernels.  CUDA's rsqrt() and 1.0 / x expand to
// ~20-25 instructions each (seed + Newton + exponent / special-case handling with a slow-path branch); ncu's source view put
// 18 % of the backward kernel's stall samples and 17 % of its instructions on them (profiles/r02z_backward_hot_lines.txt).
// Here: the hardware seed (MUFU.RSQ64H / MUFU.RCP64H through rsqrt.approx / rcp.approx, relative error < 2^-20) and two
// Newton steps (error 1.5 e^2 resp. e^2 per step: below 2^-80 after two), nothing else.  Faithful to ~1 ulp for finite,
// normal, positive arguments -- which is what the call sites guarantee (pivots checked > 0, norms of non-zero vectors);
// zero / negative / non-finite arguments give NaN instead of the IEEE special value, and every such case is already
// flagged (ISV_W_NOT_SPD / ISV_W_SINGULAR) or replaced by a select at the call site.  -DISV_FAST_SPECIAL=0 builds the
// library versions back in for A/B measurements.
#ifndef ISV_FAST_SPECIAL
#define ISV_FAST_SPECIAL 1
#endif
ISV_DI double fast_rsqrt(double x) {
#if ISV_FAST_SPECIAL
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-x * y, y, 1.0);
  return fma(0.5 * y, e, y);
#else
  return rsqrt(x);
#endif
}
ISV_DI double fast_rcp(double x) {
#if ISV_FAST_SPECIAL
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}

// Quaterniond(PS[6], PS[3], PS[4], PS[5])
ISV_DI Quat quat_from_pose(const double* ps) { return Quat{ps[6], ps[3], ps[4], ps[5]}; }

ISV_DI Quat qmul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
ISV_DI double qnorm2(const Quat& q) { return q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; }
ISV_DI Quat qconj(const Quat& q) { return Quat{q.w, -q.x, -q.y, -q.z}; }
// Eigen Quaternion::inverse(): conjugate / squaredNorm
ISV_DI Quat qinv(const Quat& q) {
  double n2 = qnorm2(q);
  double s = 1.0 / n2;
  return Quat{q.w * s, -q.x * s, -q.y * s, -q.z * s};
}
ISV_DI Quat qnormalized(const Quat& q) {
  double s = 1.0 / sqrt(qnorm2(q));
  return Quat{q.w * s, q.x * s, q.y * s, q.z * s};
}
// Eigen _transformVector: v + w*(2 u x v) + u x (2 u x v)
ISV_DI void qrot(const Quat& q, const double* v, double* o) {
  double uvx = 2.0 * (q.y * v[2] - q.z * v[1]);
  double uvy = 2.0 * (q.z * v[0] - q.x * v[2]);
  double uvz = 2.0 * (q.x * v[1] - q.y * v[0]);
  o[0] = v[0] + q.w * uvx + (q.y * uvz - q.z * uvy);
  o[1] = v[1] + q.w * uvy + (q.z * uvx - q.x * uvz);
  o[2] = v[2] + q.w * uvz + (q.x * uvy - q.y * uvx);
}
// Eigen toRotationMatrix (row-major out)
ISV_DI void q2R(const Quat& q, double* R) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// Eigen quaternion-from-matrix (row-major in)
ISV_DI Quat R2q(const double* m) {
  double t = m[0] + m[4] + m[8];
  double q[4];  // w,x,y,z
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t;
    q[2] = (m[2] - m[6]) * t;
    q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
  return Quat{q[0], q[1], q[2], q[3]};
}

// ---- 3x3 helpers (row-major) ----
ISV_DI void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A^T * B
ISV_DI void mat3_tmul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// C = A * B^T
ISV_DI void mat3_mult(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
ISV_DI void mat3_vec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
ISV_DI void mat3_tvec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
// Utility::skewSymmetric
ISV_DI void skew3(const double* v, double* S) {
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

// ---- Sophus SO3 ----
// SO3::log of a unit quaternion
ISV_DI void so3_log(const Quat& q, double* out) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
  double f;
  if (n2 < kSophusEps * kSophusEps) {
    f = 2.0 / q.w - (2.0 / 3.0) * n2 / (q.w * q.w * q.w);
  } else {
    double n = sqrt(n2);
    double at = (q.w < 0) ? atan2(-n, -q.w) : atan2(n, q.w);
    f = 2.0 * at / n;
  }
  out[0] = f * q.x;
  out[1] = f * q.y;
  out[2] = f * q.z;
}
ISV_DI Quat so3_exp(const double* om) {
  double t2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double im, re;
  if (t2 < kSophusEps * kSophusEps) {
    double t4 = t2 * t2;
    im = 0.5 - t2 / 48.0 + t4 / 3840.0;
    re = 1.0 - t2 / 8.0 + t4 / 384.0;
  } else {
    double t = sqrt(t2);
    double sh, ch;
    sincos(0.5 * t, &sh, &ch);
    im = sh / t;
    re = ch;
  }
  return Quat{re, im * om[0], im * om[1], im * om[2]};
}
// SO3 * SO3 with Sophus' first-order renormalisation
ISV_DI Quat so3_mul(const Quat& a, const Quat& b) {
  Quat q = qmul(a, b);
  double n2 = qnorm2(q);
  if (n2 != 1.0) {
    double s = 2.0 / (1.0 + n2);
    q.w *= s; q.x *= s; q.y *= s; q.z *= s;
  }
  return q;
}
// Sophus::rightJacobianInvSO3 (row-major out)
ISV_DI void so3_right_jacobian_inv(const double* phi, double* J) {
  double n2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double H[9], H2[9];
  skew3(phi, H);
  mat3_mul(H, H, H2);
  double coef;
  if (n2 > kSophusEps) {
    double n = sqrt(n2);
    if (n < kPi - kSophusEpsSqrt) {
      double s, c;
      sincos(n, &s, &c);
      coef = 1.0 / n2 - (1.0 + c) / (2.0 * n * s);
    } else {
      coef = 1.0 / (kPi * kPi);
    }
  } else {
    coef = 1.0 / 12.0;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) J[i] = 0.5 * H[i] + coef * H2[i];
  J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

// column-major 3x3 record <-> row-major register matrix
ISV_DI void load_mat3_colmajor(const double* src, double* R) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[3 * r + c] = src[3 * c + r];
}

}  // namespace isv
