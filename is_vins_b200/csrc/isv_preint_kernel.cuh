// Batched IMU pre-integration: IntegrationBase::push_back / propagate / midPointIntegration
// (/root/reference/include/factor/integration_base.h:30-36, :130-158, :54-128).
// One warp per frame interval; the K samples are sequential (midpoint rule), the 15x15 products
//   jacobian <- F jacobian ,  covariance <- F covariance F^T + V noise V^T   (:124-125)
// are warp-parallel in shared memory.  Output = the 467-double pre-integration record the
// MargBackward kernel consumes (include/isv_capi.h ISV_PREINT_REC).
#pragma once
#include "isv_window_kernels.cuh"

namespace isv {

struct NoiseCfg { double acc_n, gyr_n, acc_w, gyr_w; };

constexpr int kPreSmemPerWarp = 225 * 4 + 270 + 32;  // J, P, F, T, V(15x18), scratch

__global__ void __launch_bounds__(kThreads)
preintegrate_kernel(int n, int k_max, const int32_t* __restrict__ k_count, const double* __restrict__ imu_raw,
                    const double* __restrict__ imu_init, double* __restrict__ preint_out, NoiseCfg nz) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x * kWarpsPerCta + warp;
  if (w >= n) return;
  double* J = smem + warp * kPreSmemPerWarp;
  double* P = J + 225;
  double* F = P + 225;
  double* T = F + 225;
  double* V = T + 225;
  double* sc = V + 270;
  const double* init = imu_init + (size_t)w * 12;
  // a count outside [0, k_max] would read the next window's samples (or past the buffer): clamp it; the host entry
  // point rejects such counts up front, device callers get the clamped record
  int K = k_count ? k_count[w] : k_max;
  K = K < 0 ? 0 : (K > k_max ? k_max : K);
  for (int i = lane; i < 225; i += 32) { J[i] = (i % 16 == 0) ? 1.0 : 0.0; P[i] = 0.0; }
  // lane 0 carries the navigation state (:188-203)
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, a0[3], g0[3], ba[3], bg[3], sum_dt = 0.0;
  Quat dq{1.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 3; ++i) { a0[i] = init[i]; g0[i] = init[3 + i]; ba[i] = init[6 + i]; bg[i] = init[9 + i]; }
  // noise (:21-27): diag(acc_n^2 x3, gyr_n^2 x3, acc_n^2 x3, gyr_n^2 x3, acc_w^2 x3, gyr_w^2 x3)
  if (lane < 18) {
    const int b = lane / 3;
    const double s = (b == 0 || b == 2) ? nz.acc_n : ((b == 1 || b == 3) ? nz.gyr_n : (b == 4 ? nz.acc_w : nz.gyr_w));
    sc[lane] = s * s;
  }
  __syncwarp();
  for (int k = 0; k < K; ++k) {
    const double* smp = imu_raw + ((size_t)w * k_max + k) * 7;
    for (int i = lane; i < 225; i += 32) F[i] = 0.0;
    for (int i = lane; i < 270; i += 32) V[i] = 0.0;
    __syncwarp();
    if (lane == 0) {
      const double dt = smp[0];
      const double a1[3] = {smp[1], smp[2], smp[3]}, g1[3] = {smp[4], smp[5], smp[6]};
      double a0x[3], a1x[3], wx[3];
      for (int i = 0; i < 3; ++i) { a0x[i] = a0[i] - ba[i]; a1x[i] = a1[i] - ba[i]; wx[i] = 0.5 * (g0[i] + g1[i]) - bg[i]; }
      double un_acc_0[3], un_acc_1[3];
      qrot(dq, a0x, un_acc_0);
      Quat rq = qmul(dq, Quat{1.0, wx[0] * dt / 2, wx[1] * dt / 2, wx[2] * dt / 2});
      qrot(rq, a1x, un_acc_1);
      double Rd[9], Rr[9], Sw[9], Sa0[9], Sa1[9];
      q2R(dq, Rd);
      q2R(rq, Rr);  // result_delta_q is not yet normalised here (:95)
      skew3(wx, Sw);
      skew3(a0x, Sa0);
      skew3(a1x, Sa1);
      double ImW[9];  // I - R_w_x dt
      for (int i = 0; i < 9; ++i) ImW[i] = ((i % 4 == 0) ? 1.0 : 0.0) - Sw[i] * dt;
      double RdA0[9], RrA1[9], RrA1W[9];
      mat3_mul(Rd, Sa0, RdA0);
      mat3_mul(Rr, Sa1, RrA1);
      mat3_mul(RrA1, ImW, RrA1W);
      auto f = [&](int r, int c) -> double& { return F[r + 15 * c]; };
      auto v = [&](int r, int c) -> double& { return V[r + 15 * c]; };
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          const int i = 3 * r + c;
          const double id = (r == c) ? 1.0 : 0.0;
          f(r, c) = id;
          f(r, 3 + c) = -0.25 * RdA0[i] * dt * dt + -0.25 * RrA1W[i] * dt * dt;
          f(r, 6 + c) = id * dt;
          f(r, 9 + c) = -0.25 * (Rd[i] + Rr[i]) * dt * dt;
          f(r, 12 + c) = -0.25 * RrA1[i] * dt * dt * -dt;
          f(3 + r, 3 + c) = ImW[i];
          f(3 + r, 12 + c) = -1.0 * id * dt;
          f(6 + r, 3 + c) = -0.5 * RdA0[i] * dt + -0.5 * RrA1W[i] * dt;
          f(6 + r, 6 + c) = id;
          f(6 + r, 9 + c) = -0.5 * (Rd[i] + Rr[i]) * dt;
          f(6 + r, 12 + c) = -0.5 * RrA1[i] * dt * -dt;
          f(9 + r, 9 + c) = id;
          f(12 + r, 12 + c) = id;
          v(r, c) = 0.25 * Rd[i] * dt * dt;
          v(r, 3 + c) = 0.25 * -RrA1[i] * dt * dt * 0.5 * dt;
          v(r, 6 + c) = 0.25 * Rr[i] * dt * dt;
          v(r, 9 + c) = v(r, 3 + c);
          v(3 + r, 3 + c) = 0.5 * id * dt;
          v(3 + r, 9 + c) = 0.5 * id * dt;
          v(6 + r, c) = 0.5 * Rd[i] * dt;
          v(6 + r, 3 + c) = 0.5 * -RrA1[i] * dt * 0.5 * dt;
          v(6 + r, 6 + c) = 0.5 * Rr[i] * dt;
          v(6 + r, 9 + c) = v(6 + r, 3 + c);
          v(9 + r, 12 + c) = id * dt;
          v(12 + r, 15 + c) = id * dt;
        }
      // state update (:66-71, :148-157)
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
    __syncwarp();
    // jacobian = F * jacobian ; covariance = F * covariance * F^T + V * noise * V^T
    w_gemm<false, false>(15, 15, 15, F, 15, J, 15, T, 15, 0, lane);
    w_copy(J, T, 225, lane);
    w_gemm<false, false>(15, 15, 15, F, 15, P, 15, T, 15, 0, lane);
    w_gemm<false, true>(15, 15, 15, T, 15, F, 15, P, 15, 0, lane);
    for (int idx = lane; idx < 225; idx += 32) {
      const int i = idx % 15, j = idx / 15;
      double acc = 0.0;
      for (int l = 0; l < 18; ++l) acc = fma(V[i + 15 * l] * sc[l], V[j + 15 * l], acc);
      P[idx] += acc;
    }
    __syncwarp();
  }
  double* o = preint_out + (size_t)w * ISV_PREINT_REC;
  if (lane == 0) {
    for (int i = 0; i < 3; ++i) { o[i] = dp[i]; o[7 + i] = dv[i]; o[10 + i] = ba[i]; o[13 + i] = bg[i]; }
    o[3] = dq.x; o[4] = dq.y; o[5] = dq.z; o[6] = dq.w;
    o[16] = sum_dt;
  }
  for (int i = lane; i < 225; i += 32) { o[17 + i] = J[i]; o[242 + i] = P[i]; }
}

}  // namespace isv
