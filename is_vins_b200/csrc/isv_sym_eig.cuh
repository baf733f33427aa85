// Symmetric eigen-decomposition of the reduced system of the generic marginalization engine -- what VINS-Mono's
// `Eigen::SelfAdjointEigenSolver<MatrixXd> saes2(A)` does in MarginalizationInfo::marginalize (published
// marginalization_factor.cpp; IS-VINS keeps the same solver for its own truncations, src/estimator.cpp:920,1311,
// 1479) -- with the same algorithm class as Eigen: Householder tridiagonalization, implicit QL sweeps on the
// tridiagonal with the rotations accumulated into the orthogonal factor.  One CTA per problem, n <= 1024.
//
//   phase 1  tridiagonalize  W = Q T Q^T   (full symmetric storage in global memory / L2; per step a symv with one
//            warp per column -- lanes along the contiguous rows -- and a rank-2 update with the same mapping)
//   phase 2  form Q                          (column j of Q = H_0 ... H_{n-2} e_j is independent of every other
//            column: one warp walks all reflectors for its column held in shared memory, no CTA barriers)
//   phase 3  implicit QL                     (lane 0 of warp 0 walks the scalar recurrence and records (c, s) per
//            rotation; then every thread owns rows of Z and streams the rotation sequence through them with one
//            load + one store per rotation, coalesced across threads)
//   phase 4  outputs in ascending eigenvalue order: linearized_jacobians = S^1/2 V^T, linearized_residuals =
//            S^-1/2 V^T b with S thresholded at eps (rows of dropped eigenvalues are exactly zero).
//
// The earlier engine (pivoted Cholesky + CTA-level one-sided Jacobi on the factor rows) needed ~30 n^3 flop in
// n-1 barrier-separated rounds per sweep: 315 ms per problem at n = 307 (BASELINE configs[3] variant b).
#pragma once
#include "isv_device_math.cuh"

#include "../../include/isv_capi.h"

namespace isv {

constexpr int kSeMaxN = 1024;

__device__ __forceinline__ double se_block_sum(double v, double* red, int nw) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

__host__ __device__ inline int sym_eig_threads(int n) { return n > 128 ? 512 : 256; }
// smem doubles: d[n] e[n] tau[n] vs[n] ps[n] cs[2n] red[32] + nw * n (phase 2 column buffers)
__host__ __device__ inline size_t sym_eig_smem_doubles(int n, int nw) { return (size_t)7 * n + 32 + (size_t)nw * n; }

// W: n x n column-major symmetric input (destroyed: holds the reflectors afterwards); Z: n x n column-major scratch that
// ends up holding the eigenvectors (columns, unsorted); returns the eigenvalues in smem d[] and sets *noconv.
__device__ void sym_eig_cta(double* __restrict__ W, double* __restrict__ Z, int n, double* smem, int* noconv) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  double* d = smem;
  double* e = d + n;
  double* tau = e + n;
  double* vs = tau + n;
  double* ps = vs + n;
  double* cs = ps + n;        // 2n: c then s
  double* red = cs + 2 * n;   // 32
  double* colbuf = red + 32;  // nw * n
  __shared__ int s_m, s_cnt, s_fail;
  if (tid == 0) s_fail = 0;
  // ---------------- phase 1: Householder tridiagonalization (LAPACK dsytd2, lower variant, full storage) -------------
  for (int k = 0; k + 1 < n; ++k) {
    const int len = n - k - 1;
    double* x = W + (size_t)(k + 1) + (size_t)n * k;   // column k below the diagonal
    double part = 0.0;
    for (int i = 1 + tid; i < len; i += nt) part = fma(x[i], x[i], part);
    const double xn2 = se_block_sum(part, red, nw);
    const double alpha = x[0];
    double beta = alpha, tk = 0.0, scale = 0.0;
    if (xn2 > 0.0) {
      beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
      tk = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    for (int i = tid; i < len; i += nt) vs[i] = (i == 0) ? 1.0 : x[i] * scale;
    if (tid == 0) { d[k] = W[(size_t)k + (size_t)n * k]; e[k] = beta; tau[k] = tk; }
    __syncthreads();
    for (int i = tid; i < len; i += nt) x[i] = vs[i];        // keep v_k for phase 2 (x[0] = 1 stored explicitly)
    if (tk != 0.0) {
      double* A22 = W + (size_t)(k + 1) + (size_t)n * (k + 1);
      // p = tau * A22 v : warp per column, lanes along the rows
      for (int j = warp; j < len; j += nw) {
        const double* col = A22 + (size_t)n * j;
        double a0 = 0.0, a1 = 0.0;
        int i = lane;
        for (; i + 32 < len; i += 64) { a0 = fma(col[i], vs[i], a0); a1 = fma(col[i + 32], vs[i + 32], a1); }
        if (i < len) a0 = fma(col[i], vs[i], a0);
        a0 = warp_sum(a0 + a1);
        if (lane == 0) ps[j] = tk * a0;
      }
      __syncthreads();
      double pv = 0.0;
      for (int i = tid; i < len; i += nt) pv = fma(ps[i], vs[i], pv);
      pv = se_block_sum(pv, red, nw);
      const double a2 = -0.5 * tk * pv;
      __syncthreads();
      for (int i = tid; i < len; i += nt) ps[i] = fma(a2, vs[i], ps[i]);     // w
      __syncthreads();
      // A22 -= v w^T + w v^T
      for (int j = warp; j < len; j += nw) {
        double* col = A22 + (size_t)n * j;
        const double wj = ps[j], vj = vs[j];
        for (int i = lane; i < len; i += 32) col[i] -= fma(vs[i], wj, ps[i] * vj);
      }
    }
    __syncthreads();
  }
  if (tid == 0) { d[n - 1] = W[(size_t)(n - 1) + (size_t)n * (n - 1)]; e[n - 1] = 0.0; tau[n - 1] = 0.0; }
  __syncthreads();
  // ---------------- phase 2: Z = H_0 H_1 ... H_{n-2}, one warp per column, the column in shared memory -----------------
  {
    double* q = colbuf + (size_t)warp * n;
    for (int j = warp; j < n; j += nw) {
      for (int i = lane; i < n; i += 32) q[i] = (i == j) ? 1.0 : 0.0;
      __syncwarp();
      // H_k touches rows k+1.. ; column j of the identity is zero there while k + 1 > j
      for (int k = min(j - 1, n - 2); k >= 0; --k) {
        const double tk = tau[k];
        if (tk == 0.0) continue;
        const int len = n - k - 1;
        const double* v = W + (size_t)(k + 1) + (size_t)n * k;
        double* qq = q + k + 1;
        double a0 = 0.0, a1 = 0.0;
        int i = lane;
        for (; i + 32 < len; i += 64) { a0 = fma(v[i], qq[i], a0); a1 = fma(v[i + 32], qq[i + 32], a1); }
        if (i < len) a0 = fma(v[i], qq[i], a0);
        const double t = tk * warp_sum(a0 + a1);
        for (i = lane; i < len; i += 32) qq[i] = fma(-t, v[i], qq[i]);
        __syncwarp();
      }
      double* zc = Z + (size_t)n * j;
      for (int i = lane; i < n; i += 32) zc[i] = q[i];
      __syncwarp();
    }
  }
  __syncthreads();
  // ---------------- phase 3: implicit QL with Wilkinson shift (the tqli recurrence), rotations accumulated into Z ----
  const double epsm = 2.220446049250313e-16;
  double* cr = cs;
  double* sr = cs + n;
  for (int l = 0; l < n; ++l) {
    for (int iter = 0;; ++iter) {
      int m_scan = n - 1;
      if (warp == 0) {   // first negligible sub-diagonal at or after l (LAPACK dsteqr's neighbour-relative test)
        for (int base = l; base < n - 1; base += 32) {
          const int mm = base + lane;
          const bool small = mm < n - 1 && fabs(e[mm]) <= epsm * (fabs(d[mm]) + fabs(d[mm + 1]));
          const unsigned hit = __ballot_sync(kFullMask, small);
          if (hit) { m_scan = base + __ffs(hit) - 1; break; }
        }
      }
      if (tid == 0) {
        int m = m_scan;
        int cnt = 0;
        if (m != l) {
          if (iter >= 60) { s_fail = 1; m = l; }
        }
        if (m != l) {
          double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
          double r = sqrt(fma(g, g, 1.0));
          g = d[m] - d[l] + e[l] / (g + copysign(r, g));
          double s = 1.0, c = 1.0, p = 0.0;
          int i = m - 1;
          bool broke = false;
          for (; i >= l; --i) {
            const double f = s * e[i], b = c * e[i];
            const double r2 = fma(f, f, g * g);
            if (r2 == 0.0) { d[i + 1] -= p; e[m] = 0.0; broke = true; break; }
            double y = rsqrt(r2);
            y = y * fma(-0.5 * r2 * y, y, 1.5);          // one Newton step: 1/sqrt to ~1 ulp
            r = r2 * y;
            e[i + 1] = r;
            s = f * y;
            c = g * y;
            g = d[i + 1] - p;
            r = fma(d[i] - g, s, 2.0 * c * b);
            p = s * r;
            d[i + 1] = g + p;
            g = fma(c, r, -b);
            cr[i] = c;
            sr[i] = s;
            ++cnt;
          }
          if (!broke) { d[l] -= p; e[l] = g; e[m] = 0.0; }
        }
        s_m = m;
        s_cnt = cnt;
      }
      __syncthreads();
      const int m = s_m, cnt = s_cnt;
      __syncthreads();         // everyone has read s_m / s_cnt before thread 0 overwrites them
      if (m == l) break;       // uniform
      // rotations i = m-1, m-2, ..., m-cnt on the column pairs (i, i+1) of Z; thread per row
      for (int k = tid; k < n; k += nt) {
        double* zr = Z + k;
        double carry = zr[(size_t)n * m];
        int i = m - 1;
        const int stop = m - cnt;
        for (; i - 3 >= stop; i -= 4) {
          const double z0 = zr[(size_t)n * i], z1 = zr[(size_t)n * (i - 1)], z2 = zr[(size_t)n * (i - 2)],
                       z3 = zr[(size_t)n * (i - 3)];
          double c = cr[i], s = sr[i];
          zr[(size_t)n * (i + 1)] = fma(s, z0, c * carry);
          carry = fma(c, z0, -s * carry);
          c = cr[i - 1]; s = sr[i - 1];
          zr[(size_t)n * i] = fma(s, z1, c * carry);
          carry = fma(c, z1, -s * carry);
          c = cr[i - 2]; s = sr[i - 2];
          zr[(size_t)n * (i - 1)] = fma(s, z2, c * carry);
          carry = fma(c, z2, -s * carry);
          c = cr[i - 3]; s = sr[i - 3];
          zr[(size_t)n * (i - 2)] = fma(s, z3, c * carry);
          carry = fma(c, z3, -s * carry);
        }
        for (; i >= stop; --i) {
          const double z0 = zr[(size_t)n * i];
          const double c = cr[i], s = sr[i];
          zr[(size_t)n * (i + 1)] = fma(s, z0, c * carry);
          carry = fma(c, z0, -s * carry);
        }
        zr[(size_t)n * (i + 1)] = carry;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) *noconv = s_fail;
  __syncthreads();
}

// linearized_jacobians / residuals from the reduced system: one CTA per problem.  A_red is read (not modified);
// Wbuf / Zbuf: n x n scratch per problem.
__global__ void sym_eig_prior_kernel(int n, const double* __restrict__ A_red, const double* __restrict__ b_red,
                                     double* __restrict__ Wbuf, double* __restrict__ Zbuf, double* __restrict__ LJ_all,
                                     double* __restrict__ LR_all, int32_t* __restrict__ rank_out, int32_t* __restrict__ status,
                                     double eps, size_t w_stride) {
  extern __shared__ double smem[];
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const double* Ar = A_red + (size_t)prob * n * n;
  const double* br = b_red + (size_t)prob * n;
  double* W = Wbuf + (size_t)prob * w_stride;
  double* Z = Zbuf + (size_t)prob * n * n;
  double* LJ = LJ_all + (size_t)prob * n * n;
  double* LR = LR_all + (size_t)prob * n;
  __shared__ int s_noconv, s_kept;
  // W = (A_red + A_red^T) / 2   (Eigen reads the lower triangle only; VINS-Mono symmetrises Amm the same way)
  for (size_t idx = tid; idx < (size_t)n * n; idx += nt) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    W[idx] = 0.5 * (Ar[idx] + Ar[j + (size_t)n * i]);
  }
  __syncthreads();
  sym_eig_cta(W, Z, n, smem, &s_noconv);
  double* d = smem;
  double* gb = smem + n;          // v_k . b_red     (e[] is dead)
  int* perm = reinterpret_cast<int*>(smem + 2 * n);   // tau[] is dead
  // rank of every eigenvalue in ascending order (ties by index) and v_k . b
  for (int k = warp; k < n; k += nw) {
    const double* z = Z + (size_t)n * k;
    double a = 0.0;
    for (int i = lane; i < n; i += 32) a = fma(z[i], br[i], a);
    a = warp_sum(a);
    if (lane == 0) gb[k] = a;
  }
  if (tid == 0) s_kept = 0;
  for (int k = tid; k < n; k += nt) perm[k] = k;      // stays a valid index table even if an eigenvalue is NaN
  __syncthreads();
  int kept_local = 0;
  for (int k = tid; k < n; k += nt) {
    const double lam = d[k];
    int below = 0;
    for (int t = 0; t < n; ++t) {
      const double lt = d[t];
      below += (lt < lam || (lt == lam && t < k)) ? 1 : 0;
    }
    perm[below] = k;
    kept_local += (lam > eps) ? 1 : 0;
  }
  if (kept_local) atomicAdd(&s_kept, kept_local);
  __syncthreads();
  // row r of linearized_jacobians (column-major n x n) = sqrt(lam) v^T for the r-th smallest eigenvalue, 0 if <= eps
  for (int r = warp; r < n; r += nw) {
    const int k = perm[r];
    const double lam = d[k];
    const bool keep = lam > eps;
    const double sq = keep ? sqrt(lam) : 0.0;
    const double* z = Z + (size_t)n * k;
    for (int j = lane; j < n; j += 32) LJ[r + (size_t)n * j] = keep ? sq * z[j] : 0.0;
    if (lane == 0) LR[r] = keep ? gb[k] / sq : 0.0;
  }
  if (tid == 0) {
    rank_out[prob] = s_kept;
    if (s_noconv && status) atomicOr(status + prob, ISV_W_EIG_NOCONV);
  }
}

// unit-test hook: eigenvalues (ascending) + eigenvectors (columns, column-major) of nb symmetric matrices
__global__ void sym_eig_test_kernel(int n, double* __restrict__ A, double* __restrict__ Zbuf, double* __restrict__ lam_out,
                                    double* __restrict__ V_out, int32_t* __restrict__ info) {
  extern __shared__ double smem[];
  const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  __shared__ int s_noconv;
  double* W = A + (size_t)prob * n * n;
  double* Z = Zbuf + (size_t)prob * n * n;
  sym_eig_cta(W, Z, n, smem, &s_noconv);
  double* d = smem;
  int* perm = reinterpret_cast<int*>(smem + 2 * n);
  for (int k = tid; k < n; k += nt) perm[k] = k;
  __syncthreads();
  for (int k = tid; k < n; k += nt) {
    const double lam = d[k];
    int below = 0;
    for (int t = 0; t < n; ++t) below += (d[t] < lam || (d[t] == lam && t < k)) ? 1 : 0;
    perm[below] = k;
  }
  __syncthreads();
  for (int r = warp; r < n; r += nw) {
    const int k = perm[r];
    const double* z = Z + (size_t)n * k;
    double* o = V_out + (size_t)prob * n * n + (size_t)n * r;
    for (int j = lane; j < n; j += 32) o[j] = z[j];
    if (lane == 0) lam_out[(size_t)prob * n + r] = d[k];
  }
  if (tid == 0) info[prob] = s_noconv;
}

}  // namespace isv
