// Batched IMU pre-integration: IntegrationBase::push_back / propagate / midPointIntegration
// (/root/reference/include/factor/integration_base.h:30-36, :130-158, :54-128).
// One warp per frame interval; the K samples are sequential (midpoint rule), the 15x15 products
//   jacobian <- F jacobian ,  covariance <- F covariance F^T + V noise V^T   (:124-125)
// run one column per lane in registers, exploiting the block structure of F and V.  Output = the 467-double pre-integration record the
// MargBackward kernel consumes (include/isv_capi.h ISV_PREINT_REC).
#pragma once
#include "isv_window_kernels.cuh"

namespace isv {

struct NoiseCfg { double acc_n, gyr_n, acc_w, gyr_w; };

constexpr int kPreLd = 17;                       // transposition tile of the covariance: 15 x 15, leading dimension 17
constexpr int kPreSmemPerWarp = 15 * kPreLd + 1; // (even: 16-byte aligned per warp)

// Column-per-lane form.  Lane c < 15 owns column c of `jacobian`, lane 15 + c column c of `covariance`, in registers for
// the whole interval.  F (15 x 15, :73-100) has five distinct dense 3 x 3 blocks -- built from Rd = R(delta_q),
// Rr = R(result_delta_q), their products with the skew matrices and I - [w]x dt -- the rest is 0, I or I dt, so  y = F x
// is four 3 x 3 matrix-vector products per column (56 flops instead of 450), evaluated by every lane on its own column
// from blocks that every lane keeps in registers (the scalar prologue is evaluated redundantly by all lanes: no shared
// memory, no barrier, no idle lanes).  covariance <- F P F^T + V N V^T uses the symmetry of P:  u_c = F p_c  is column c of
// F P; its row c, fetched through a shared-memory tile, is column c of (F P)^T = P F^T, and F times it is column c of
// F P F^T.  V N V^T (V 15 x 18 with the same blocks, N diagonal) is added column by column the same way.
// (The first version built F and V densely in shared memory on lane 0 and ran three generic 15 x 15 x 15 warp GEMMs per
// sample: ~2.5 k warp instructions per sample, 0.2 ms per 2368 windows -- more than the whole marginalization step.)
__global__ void __launch_bounds__(kThreads)
preintegrate_kernel(int n, int k_max, const int32_t* __restrict__ k_count, const double* __restrict__ imu_raw,
                    const double* __restrict__ imu_init, double* __restrict__ preint_out, NoiseCfg nz) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x * kWarpsPerCta + warp;
  if (w >= n) return;
  double* U = smem + warp * kPreSmemPerWarp;
  const double* init = imu_init + (size_t)w * 12;
  // a count outside [0, k_max] would read the next window's samples (or past the buffer): clamp it; the host entry
  // point rejects such counts up front, device callers get the clamped record
  int K = k_count ? k_count[w] : k_max;
  K = K < 0 ? 0 : (K > k_max ? k_max : K);
  const bool isP = lane >= 15 && lane < 30;
  const int c = lane < 15 ? lane : (isP ? lane - 15 : 0);   // my column
  const int cb = c / 3, ci = c - 3 * cb;                    // its 3-block and the row inside the block
  double x[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) x[i] = (!isP && lane < 15 && i == c) ? 1.0 : 0.0;   // jacobian = I, covariance = 0
  // the navigation state (:188-203), carried redundantly by every lane
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, a0[3], g0[3], ba[3], bg[3], sum_dt = 0.0;
  Quat dq{1.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i) { a0[i] = init[i]; g0[i] = init[3 + i]; ba[i] = init[6 + i]; bg[i] = init[9 + i]; }
  // noise (:21-27): diag(acc_n^2 x3, gyr_n^2 x3, acc_n^2 x3, gyr_n^2 x3, acc_w^2 x3, gyr_w^2 x3)
  const double an2 = nz.acc_n * nz.acc_n, gn2 = nz.gyr_n * nz.gyr_n, aw2 = nz.acc_w * nz.acc_w, gw2 = nz.gyr_w * nz.gyr_w;
  const double* smp = imu_raw + (size_t)w * k_max * 7;
#pragma unroll 1
  for (int k = 0; k < K; ++k, smp += 7) {
    const double dt = smp[0];
    const double a1[3] = {smp[1], smp[2], smp[3]}, g1[3] = {smp[4], smp[5], smp[6]};
    double a0x[3], a1x[3], wx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a0x[i] = a0[i] - ba[i]; a1x[i] = a1[i] - ba[i]; wx[i] = 0.5 * (g0[i] + g1[i]) - bg[i]; }
    double un_acc_0[3], un_acc_1[3];
    qrot(dq, a0x, un_acc_0);
    const Quat rq = qmul(dq, Quat{1.0, wx[0] * dt / 2, wx[1] * dt / 2, wx[2] * dt / 2});
    qrot(rq, a1x, un_acc_1);
    double Rd[9], Rr[9], Sw[9], Sa0[9], Sa1[9], W[9], S1[9], S3[9], RA[9];
    q2R(dq, Rd);
    q2R(rq, Rr);  // result_delta_q is not yet normalised here (:95)
    skew3(wx, Sw);
    skew3(a0x, Sa0);
    skew3(a1x, Sa1);
#pragma unroll
    for (int i = 0; i < 9; ++i) W[i] = ((i % 4 == 0) ? 1.0 : 0.0) - Sw[i] * dt;   // I - R_w_x dt
    {
      double RdA0[9], RrA1W[9];
      mat3_mul(Rd, Sa0, RdA0);
      mat3_mul(Rr, Sa1, RA);
      mat3_mul(RA, W, RrA1W);
#pragma unroll
      for (int i = 0; i < 9; ++i) { S1[i] = RdA0[i] + RrA1W[i]; S3[i] = Rd[i] + Rr[i]; }
    }
    const double dt2 = dt * dt, dt3 = dt2 * dt;
    // y = F x with  F = [ I  -dt2/4 S1   I dt  -dt2/4 S3   dt3/4 RA ;  0  W  0  0  -I dt ;  0  -dt/2 S1  I  -dt/2 S3  dt2/2 RA ;
    //                     0 0 0 I 0 ;  0 0 0 0 I ]                                                    (:73-100)
    auto apply_F = [&](const double* __restrict__ v, double* __restrict__ y) {
      double t1[3], t3[3], t4[3], t11[3];
      mat3_vec(S1, v + 3, t1);
      mat3_vec(S3, v + 9, t3);
      mat3_vec(RA, v + 12, t4);
      mat3_vec(W, v + 3, t11);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        y[i] = v[i] + (-0.25 * dt2) * t1[i] + dt * v[6 + i] + (-0.25 * dt2) * t3[i] + (0.25 * dt3) * t4[i];
        y[3 + i] = t11[i] - dt * v[12 + i];
        y[6 + i] = (-0.5 * dt) * t1[i] + v[6 + i] + (-0.5 * dt) * t3[i] + (0.5 * dt2) * t4[i];
        y[9 + i] = v[9 + i];
        y[12 + i] = v[12 + i];
      }
    };
    double y[15];
    apply_F(x, y);
    if (isP) {
#pragma unroll
      for (int i = 0; i < 15; ++i) U[c * kPreLd + i] = y[i];
    }
    __syncwarp();
    if (isP) {
      double r[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) r[i] = U[i * kPreLd + c];     // row c of F P = column c of P F^T
      apply_F(r, y);
      // + column c of V N V^T (:101-125).  Row c of V: blocks [s0 Rd_i | s1 RA_i | s0 Rr_i | s1 RA_i | 0 | 0] for the
      // position (s0 = dt2/4, s1 = -dt3/8) and velocity (s0 = dt/2, s1 = -dt2/4) rows, [0 | dt/2 e_i | 0 | dt/2 e_i | 0 | 0]
      // for the rotation rows, dt e_i in block 4 / 5 for the bias rows.
      const double s0 = cb == 0 ? 0.25 * dt2 : (cb == 2 ? 0.5 * dt : 0.0);
      const double s1 = cb == 0 ? -0.125 * dt3 : (cb == 2 ? -0.25 * dt2 : 0.0);
      double z0[3], z2[3], q13[3], z4[3], z5[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double rd = ci == 0 ? Rd[j] : (ci == 1 ? Rd[3 + j] : Rd[6 + j]);
        const double rr = ci == 0 ? Rr[j] : (ci == 1 ? Rr[3 + j] : Rr[6 + j]);
        const double ra = ci == 0 ? RA[j] : (ci == 1 ? RA[3 + j] : RA[6 + j]);
        const double e = (j == ci) ? 1.0 : 0.0;
        z0[j] = an2 * (s0 * rd);
        z2[j] = an2 * (s0 * rr);
        q13[j] = 2.0 * gn2 * (s1 * ra) + (cb == 1 ? gn2 * dt * e : 0.0);   // N (v_1 + v_3): the two gyro-noise blocks coincide
        z4[j] = cb == 3 ? aw2 * dt * e : 0.0;
        z5[j] = cb == 4 ? gw2 * dt * e : 0.0;
      }
      double ta[3], tb[3], tc[3];
      mat3_vec(Rd, z0, ta);
      mat3_vec(RA, q13, tb);
      mat3_vec(Rr, z2, tc);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        y[i] += (0.25 * dt2) * (ta[i] + tc[i]) + (-0.125 * dt3) * tb[i];
        y[3 + i] += (0.5 * dt) * q13[i];
        y[6 + i] += (0.5 * dt) * (ta[i] + tc[i]) + (-0.25 * dt2) * tb[i];
        y[9 + i] += dt * z4[i];
        y[12 + i] += dt * z5[i];
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 15; ++i) x[i] = y[i];
    // state update (:66-71, :148-157)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double ua = 0.5 * (un_acc_0[i] + un_acc_1[i]);
      dp[i] = dp[i] + dv[i] * dt + 0.5 * ua * dt * dt;
      dv[i] = dv[i] + ua * dt;
      a0[i] = a1[i];
      g0[i] = g1[i];
    }
    dq = qnormalized(rq);
    sum_dt += dt;
  }
  double* o = preint_out + (size_t)w * ISV_PREINT_REC;
  if (lane == 0) {
    for (int i = 0; i < 3; ++i) { o[i] = dp[i]; o[7 + i] = dv[i]; o[10 + i] = ba[i]; o[13 + i] = bg[i]; }
    o[3] = dq.x; o[4] = dq.y; o[5] = dq.z; o[6] = dq.w;
    o[16] = sum_dt;
  }
  if (lane < 30) {
    double* oc = o + (isP ? 242 : 17) + 15 * c;   // column-major 15 x 15
#pragma unroll
    for (int i = 0; i < 15; ++i) oc[i] = x[i];
  }
}

}  // namespace isv
